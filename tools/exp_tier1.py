"""Debug-hook builds only: time the fine-pass tier-1 launch (single-pass fp16, 240 000 tiles) under NSR_EXPERIMENT (mlp_forward.cu)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O
import neural_sim_nerf_b200 as nsr
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
m = nsr.NeRF(); m.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}); m.cuda()
pf = nsr.packed_weights(m)
L = nsr.lib(); P = lambda t: ctypes.c_void_p(t.data_ptr())
n, T = 160000, 192
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
zf = torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, T, device='cuda').expand(n, T).contiguous()
raw = torch.empty(n, T, 4, device='cuda')
aset = torch.zeros(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device='cuda')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
exp = int(os.environ.get('NSR_EXPERIMENT', 0))
grid = exp - 100 if exp >= 100 else 148
for name, fn in (('tier 1 (8 steps)', lambda: L.nsr_mlp_two_tier(P(rays), P(zf), n, T, P(pf), P(raw), P(aset), None, 1, None)),
                 ('fp16x3 dense', lambda: L.nsr_mlp_forward(P(rays), P(zf), n, T, P(pf), 0, P(raw), None))):
    reps = 1 if grid < 148 else 4
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        assert fn() == 0
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'NSR_EXPERIMENT={exp:3d} {name:18s} {ms:8.3f} ms   per tile and SM: {ms * 1e3 / (240000 / grid):7.3f} us', flush=True)
