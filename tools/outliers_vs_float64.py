"""For the rays of a view on which this renderer and fp32 eager PyTorch differ by more than 1e-3: who is closer to a float64 evaluation of the
same algorithm (oracle in double on the GPU)?  Supports (or refutes) the statement that the residual outliers are rays whose pixel
the reference's own fp32 rounding decides."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O
import neural_sim_nerf_b200 as nsr
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
sds = [{k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)} for p in ('coarse/', 'fine/')]
nets = []
for sd in sds:
    m = nsr.NeRF(); m.load_state_dict(sd); nets.append(m.cuda())
H = W = 400; n = H * W
for phi in (22.5, 67.5, 112.5, 202.5):
    pose = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, pose)
    packed = O.pack_rays(ro.reshape(-1, 3), rd.reshape(-1, 3), O.YCBV_NEAR, O.YCBV_FAR).cuda()
    with torch.device('cuda'), torch.no_grad():
        sdc = {k: v.cuda() for k, v in sds[0].items()}; sdf = {k: v.cuda() for k, v in sds[1].items()}
        ref = torch.cat([O.render_rays(packed[i:i + 16384], sdc, sdf, 64, 128)['rgb_map'] for i in range(0, n, 16384)], 0)
        got = nsr.render_rays(packed, nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map']
        d = (got - ref).abs().max(-1).values
        bad = torch.nonzero(d > 1e-3).reshape(-1)
        if bad.numel() == 0:
            print(f'phi {phi}: no ray beyond 1e-3 (max {float(d.max()):.2e})')
            continue
        sdc64 = {k: v.cuda().double() for k, v in sds[0].items()}; sdf64 = {k: v.cuda().double() for k, v in sds[1].items()}
        r64 = O.render_rays(packed[bad].double(), sdc64, sdf64, 64, 128)['rgb_map']
        for i, b in enumerate(bad.tolist()):
            print(f'phi {phi} ray {b}: |ours - fp32 eager| {float(d[b]):.2e};  |ours - float64| {float((got[b].double() - r64[i]).abs().max()):.2e};  '
                  f'|fp32 eager - float64| {float((ref[b].double() - r64[i]).abs().max()):.2e}')
