"""Debug: dump CTA 0's per-step timeline of the MLP kernel (NSR_TRACE_FILE) for both precisions."""
import ctypes, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
os.environ['NSR_TRACE_FILE'] = 'gpurun_out/trace.txt'
import numpy as np, torch
import neural_sim_nerf_b200 as nsr, nerf_oracle as O
z = np.load('tests/golden/wfit.npz')
sd = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
net = nsr.NeRF(); net.load_state_dict(sd); net.cuda()
pf = nsr.packed_weights(net)
L = nsr.lib(); n, T = 160000, 192
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
zf = torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, T, device='cuda').expand(n, T).contiguous()
raw = torch.empty(n, T, 4, device='cuda')
P = lambda t: ctypes.c_void_p(t.data_ptr())
for flags in (0, 8, 0, 8):
    L.nsr_mlp_forward(P(rays), P(zf), n, T, P(pf), flags, P(raw), None)
torch.cuda.synchronize()
print(open('gpurun_out/trace.txt').read()[:200])
