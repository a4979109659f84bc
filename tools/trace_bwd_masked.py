"""One forward (saving the ReLU sign bits) + two saved-bits backward launches over 160000 rays x 192 samples (for ncu):
  ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_bwd -s 1 -c 1 -f -o gpurun_out/prof_bwd_masked python tools/trace_bwd_masked.py
"""
import ctypes, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import neural_sim_nerf_b200 as nsr, nerf_oracle as O
z = np.load('tests/golden/wfit.npz')
nets = []
for pre in ('coarse/', 'fine/'):
    m = nsr.NeRF(); m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}); nets.append(m.cuda())
pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
L = nsr.lib(); n, S, Ni = 160000, 64, 128; T = S + Ni
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
new = lambda *s: torch.empty(*s, device='cuda')
rgb, raw, zv = new(n, 3), new(n, T, 4), new(n, T)
ws = torch.empty(L.nsr_render_workspace_bytes(n, S, Ni), dtype=torch.uint8, device='cuda')
mask = torch.empty(L.nsr_relu_mask_bytes(n, T), dtype=torch.uint8, device='cuda')
assert L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw), P(zv), None,
                                    P(mask), None, None, P(ws), ws.numel(), None) == 0
g = torch.randn(n, 3, device='cuda'); d_rays = new(n, 11)
bws = torch.empty(L.nsr_render_backward_workspace_bytes(n, T), dtype=torch.uint8, device='cuda')
for _ in range(2):
    assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(pf), 0, P(g), P(d_rays), None, None, None, P(mask), None, P(bws), bws.numel(), None) == 0
torch.cuda.synchronize()
