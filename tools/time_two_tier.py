"""Times every MLP launch of one 400x400 render (64 + 128 samples) by itself: coarse / fine tier 1 (every point, single-pass fp16),
coarse / fine tier 2 (active points, fp16x3), the dense fp16x3 and single-pass launches, and the whole forward call.  GPU box only.
A/B two builds with NSR_LIB_PATH=build/libnsr_X.so (tools/build_variant.sh)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O  # noqa: E402
import neural_sim_nerf_b200 as nsr  # noqa: E402

z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
nets = []
for p in ('coarse/', 'fine/'):
    m = nsr.NeRF()
    m.load_state_dict({k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)})
    nets.append(m.cuda())
pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
L = nsr.lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
H = W = 400
S, Ni = 64, 128
T = S + Ni
n = H * W
rays = nsr.make_rays(H, W, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
new = lambda *s: torch.empty(*s, device='cuda')
rgb, z0, zf, raw0, rawf = new(n, 3), new(n, S), new(n, T), new(n, S, 4), new(n, T, 4)
ws = torch.empty(L.nsr_render_workspace_bytes(n, S, Ni), dtype=torch.uint8, device='cuda')
as0 = torch.zeros(L.nsr_active_set_bytes(n, S), dtype=torch.uint8, device='cuda')
as1 = torch.zeros(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ok(rc):
    assert rc == 0, L.nsr_last_error()


def forward():
    ok(L.nsr_render_rays_forward(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, None, P(zf), None,
                                 P(ws), ws.numel(), None))


print(f'lib: {os.environ.get("NSR_LIB_PATH", "default")}')
t_all = timed(forward)
# coarse depths = what the forward call used: linspace near..far
t = torch.linspace(0, 1, S, device='cuda')
z0.copy_((O.YCBV_NEAR * (1 - t) + O.YCBV_FAR * t).expand(n, S))
rows = []
for name, zz, TT, blob, raw, aset in (('coarse', z0, S, pc, raw0, as0), ('fine', zf, T, pf, rawf, as1)):
    t1 = timed(lambda: ok(L.nsr_mlp_two_tier(P(rays), P(zz), n, TT, P(blob), P(raw), P(aset), None, 1, None)))
    cnt = int(aset[:4].view(torch.int32).item())
    t2 = timed(lambda: ok(L.nsr_mlp_two_tier(P(rays), P(zz), n, TT, P(blob), P(raw), P(aset), None, 2, None)))
    td = timed(lambda: ok(L.nsr_mlp_forward(P(rays), P(zz), n, TT, P(blob), 0, P(raw), None)), reps=3)
    tf = timed(lambda: ok(L.nsr_mlp_forward(P(rays), P(zz), n, TT, P(blob), 8, P(raw), None)), reps=3)
    pts = n * TT
    f1, f2 = 2 * 491264, 2 * 593408
    print(f'{name:6s} tier 1: {t1:7.3f} ms ({pts * f1 / t1 / 1e9:6.1f} TFLOP/s)   tier 2: {t2:7.3f} ms on {cnt / pts:.4f} of the points '
          f'({cnt * f2 / t2 / 1e9:6.1f} algorithmic TFLOP/s)   dense fp16x3: {td:7.3f} ms ({pts * f2 / td / 1e9:6.1f})   '
          f'single-pass fp16: {tf:7.3f} ms ({pts * f2 / tf / 1e9:6.1f})')
print(f'whole forward call: {t_all:7.3f} ms  = {n / t_all / 1e3:.3f} M rays/s')
