import ctypes, os, sys
import numpy as np, torch
ROOT='/root/repo'
sys.path[:0]=[ROOT, ROOT+'/oracle']
import nerf_oracle as O
import neural_sim_nerf_b200 as nsr
z=np.load(ROOT+'/tests/golden/wfit.npz')
nets=[]
for pre in ('coarse/','fine/'):
    m=nsr.NeRF(); m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}); nets.append(m.cuda())
pc,pf=nsr.packed_weights(nets[0]),nsr.packed_weights(nets[1])
L=nsr.lib(); n,S,Ni=160000,64,128; T=S+Ni
P=lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
for phi in (22.5, 112.5):
    rays=nsr.make_rays(400,400,O.YCBV_K_400,O.pose_spherical(90.,phi-180.,1.01)[:3,:4],O.YCBV_NEAR,O.YCBV_FAR)
    rgb=torch.empty(n,3,device='cuda'); zv=torch.empty(n,T,device='cuda'); raw=torch.empty(n,T,4,device='cuda')
    ws=torch.empty(L.nsr_render_workspace_bytes(n,S,Ni),dtype=torch.uint8,device='cuda')
    assert L.nsr_render_rays_forward(P(rays),n,P(pc),P(pf),S,Ni,0,None,None,P(rgb),None,None,None,None,None,None,None,P(zv),None,P(ws),ws.numel(),None)==0
    aset=torch.zeros(L.nsr_active_set_bytes(n,T),dtype=torch.uint8,device='cuda')
    assert L.nsr_mlp_two_tier(P(rays),P(zv),n,T,P(pf),P(raw),P(aset),None,1,None)==0
    s1=raw[...,3].clone()
    assert L.nsr_mlp_forward(P(rays),P(zv),n,T,P(pf),0,P(raw),None)==0
    s=raw[...,3]
    torch.cuda.synchronize()
    tot=s1.numel()
    for lo,hi in ((-1e9,-8),(-8,-6),(-6,-4),(-4,-3),(-3,-2),(-2,-1),(-1,0),(0,1e9)):
        print(f'phi {phi}: sigma~ in ({lo},{hi}]: {int(((s1>lo)&(s1<=hi)).sum())/tot:.4f}   true sigma: {int(((s>lo)&(s<=hi)).sum())/tot:.4f}')
    print('max |s1-s| overall', float((s1-s).abs().max()), ' on sigma<=0 points:', float((s1-s).abs()[s<=0].max()))
    # how many active points sit behind an optical depth that makes them invisible?  tau_lb(i) = sum_{j<i} max(sigma~_j - 1.5, 0) * dist_j
    d = rays[:, 3:6].norm(dim=-1, keepdim=True)
    dist = torch.cat([zv[:, 1:] - zv[:, :-1], torch.full_like(zv[:, :1], 1e10)], -1) * d
    od = torch.clamp(s1 - 1.5, min=0) * dist
    tau = torch.cumsum(od, -1) - od                      # exclusive prefix
    act = s1 > -4
    for eps in (1e-3, 1e-4, 1e-5, 1e-6, 1e-8):
        hid = act & (tau >= -np.log(eps))
        print(f'phi {phi}: T <= {eps:g}: {int(hid.sum()) / int(act.sum()):.3f} of the active points are hidden ({int(hid.sum())} of {int(act.sum())})')
