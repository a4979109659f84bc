"""One launch of every kernel of a 400x400 render (64 + 128 samples) and of its pose-gradient backward pass, for ncu:
  ncu --set full --clock-control none --import-source on -k regex:'nerf_mlp|raw2outputs|resample_merge|ray_grad' -f -o gpurun_out/r02_kernels \\
      python tools/ncu_kernels.py
Launch order (the keys tools/ncu_traffic.py gives them):
  forward render (two-tier): coarse_tier1, coarse_tier2, coarse_redo (exits), composite_coarse, resample_merge,
      fine_tier1, fine_tier2, fine_redo (exits), composite_fine
  the same forward keeping raw, the sign bits of the active points and the active list for the backward pass: the same nine with the
      prefix "pg_" (pg_fine_tier2 writes 272 B of sign bits per active point)
  backward over the active set: composite_bwd, bwd_masked_active, ray_grad_reduce
  fine_dense: the fine pass evaluated densely (fp16x3, every point)
"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O
import neural_sim_nerf_b200 as nsr
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
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
aset = torch.empty(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device='cuda')
g = torch.randn(n, 3, device='cuda'); d_rays = new(n, 11)
bws = torch.empty(L.nsr_render_backward_workspace_bytes(n, T), dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()
assert L.nsr_render_rays_forward(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, None, None, None,
                                 P(ws), ws.numel(), None) == 0, L.nsr_last_error()
assert L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw), P(zv), None,
                                    P(mask), None, P(aset), P(ws), ws.numel(), None) == 0, L.nsr_last_error()
assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(pf), 0, P(g), P(d_rays), None, None, None, P(mask), P(aset), P(bws), bws.numel(), None) == 0
assert L.nsr_mlp_forward(P(rays), P(zv), n, T, P(pf), 0, P(raw), None) == 0
torch.cuda.synchronize()
print('active points of the fine pass:', int(aset[:4].view(torch.int32).item()))
