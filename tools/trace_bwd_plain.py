"""Two backward launches over 160000 rays x 192 samples (for ncu)."""
import ctypes, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import neural_sim_nerf_b200 as nsr, nerf_oracle as O
z = np.load('tests/golden/wfit.npz')
sd = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
net = nsr.NeRF(); net.load_state_dict(sd); net.cuda()
pf = nsr.packed_weights(net)
L = nsr.lib(); n, T = 160000, 192
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
zf = torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, T, device='cuda').expand(n, T).contiguous()
raw = torch.randn(n, T, 4, device='cuda')
g = torch.randn(n, 3, device='cuda'); d_rays = torch.empty(n, 11, device='cuda')
wsb = L.nsr_render_backward_workspace_bytes(n, T); ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    assert L.nsr_render_rays_backward(P(rays), P(zf), P(raw), n, T, P(pf), 0, P(g), P(d_rays), None, None, None, P(ws), wsb, None) == 0
torch.cuda.synchronize()
