import os, sys, numpy as np, torch
ROOT='/root/repo'; sys.path[:0]=[ROOT, ROOT+'/oracle']
import nerf_oracle as O, neural_sim_nerf_b200 as nsr
z=np.load(ROOT+'/tests/golden/wfit.npz'); nets=[]
for pre in ('coarse/','fine/'):
    m=nsr.NeRF(); m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}); nets.append(m.cuda().requires_grad_(False))
H=W=400; K=O.YCBV_K_400
kw=dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
psi=torch.full((8,),0.02); psi[4]=0.86
prob=torch.softmax(psi.cuda()/0.25,0)
_,log=nsr.sample_pose_nograd(prob,2,0.1,seed=0,device='cuda')
poses=nsr.sample_pose(prob,2,0.1,log)
pose=poses[1][:3,:4].detach()
with torch.no_grad():
    ro,rd=nsr.get_rays(H,W,K,pose)
    br=torch.stack([ro.reshape(-1,3),rd.reshape(-1,3)],0)
    a=nsr.render(H,W,K,chunk=512,rays=br,retraw=True,**kw)[0]      # dense, torch-made rays, one launch
    b=nsr.render(H,W,K,chunk=512,rays=br,**kw)[0]                  # two-tier, same rays
    c=nsr.render(H,W,K,chunk=512,c2w=pose,**kw)[0].reshape(-1,3)   # kernel-made rays
    d=torch.cat([nsr.render(H,W,K,chunk=512,rays=br[:,i:i+512],retraw=True,**kw)[0] for i in range(0,H*W,512)],0)  # per-chunk calls
    rk=nsr.make_rays(H,W,K,pose,O.YCBV_NEAR,O.YCBV_FAR)
print('dense vs two-tier (same rays):', float((a-b).abs().max()))
print('one launch vs 313 chunk calls (same rays):', float((a-d).abs().max()))
diff=(a-c).abs().max(-1).values
print('torch rays vs kernel rays: max', float(diff.max()), ' pixels > 1e-3:', int((diff>1e-3).sum()), ' > 1e-4:', int((diff>1e-4).sum()))
print('ray direction max diff', float((rk[:,3:6]-rd.reshape(-1,3)).abs().max()), 'origin', float((rk[:,0:3]-ro.reshape(-1,3)).abs().max()))
idx=torch.argsort(diff,descending=True)[:5]
print('worst pixels', [(int(i)//W, int(i)%W, float(diff[i])) for i in idx])
