import sys, numpy as np, torch
sys.path[:0]=['/root/repo','/root/repo/oracle']
import nerf_oracle as O, neural_sim_nerf_b200 as nsr
z=np.load('/root/repo/tests/golden/wfit.npz'); nets=[]
for pre in ('coarse/','fine/'):
    m=nsr.NeRF(); m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}); nets.append(m.cuda().requires_grad_(False))
H=W=400
ro,rd=O.get_rays(H,W,O.YCBV_K_400,O.pose_spherical(90.,22.5-180.,1.01)[:3,:4])
sel=torch.arange(0,H*W,97)[:1500]
rays=O.pack_rays(ro.reshape(-1,3)[sel],rd.reshape(-1,3)[sel],O.YCBV_NEAR,O.YCBV_FAR).cuda().requires_grad_(True)
out=nsr.render_rays(rays,nets[0],None,64,N_importance=128,network_fine=nets[1])
g,=torch.autograd.grad(out['rgb_map'],rays,grad_outputs=torch.ones_like(out['rgb_map']))
nsr.lib().nsr_set_tier1_pair(1)
with torch.no_grad(): out2=nsr.render_rays(rays.detach(),nets[0],None,64,N_importance=128,network_fine=nets[1])
torch.cuda.synchronize()
print('ok', float(out['rgb_map'].sum()), float(g.abs().sum()), bool(torch.equal(out['rgb_map'].detach(), out2['rgb_map'])))
