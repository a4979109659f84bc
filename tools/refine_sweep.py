"""Coarse-pass refinement: points picked, time of the stage, and rays beyond 1e-3 against fp32 eager PyTorch, per acc0 limit and view."""
import ctypes, os, sys
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
sdc = {k: v.cuda() for k, v in sds[0].items()}; sdf = {k: v.cuda() for k, v in sds[1].items()}
L = nsr.lib(); P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
H = W = 400; n = H * W; S = 64
pc = nsr.packed_weights(nets[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for phi in (22.5, 67.5, 112.5):
    pose = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, pose)
    packed = O.pack_rays(ro.reshape(-1, 3), rd.reshape(-1, 3), O.YCBV_NEAR, O.YCBV_FAR).cuda()
    with torch.device('cuda'), torch.no_grad():
        ref = torch.cat([O.render_rays(packed[i:i + 16384], sdc, sdf, 64, 128)['rgb_map'] for i in range(0, n, 16384)], 0)
    t = torch.linspace(0, 1, S, device='cuda')
    z0 = (O.YCBV_NEAR * (1 - t) + O.YCBV_FAR * t).expand(n, S).contiguous()
    raw0 = torch.empty(n, S, 4, device='cuda')
    assert L.nsr_mlp_forward(P(packed), P(z0), n, S, P(pc), 0, P(raw0), None) == 0
    ws = torch.zeros(L.nsr_coarse_refine_workspace_bytes(n), dtype=torch.uint8, device='cuda')
    for lim, shi in ((None, 0.), (0.9, 0.), (0.9, 2.), (0.9, 10.), (0.9, 30.)):
        L.nsr_set_coarse_refine_sigma(shi)
        if lim is None:
            L.nsr_set_coarse_refine(0)
            cnt, ms = 0, 0.0
        else:
            L.nsr_set_coarse_refine(1)
            L.nsr_set_coarse_refine_limit(lim)
            r = raw0.clone()
            L.nsr_coarse_refine(P(packed), P(z0), n, S, P(pc), P(r), P(ws), ws.numel(), None)
            torch.cuda.synchronize()
            cnt = int(ws[:4].view(torch.int32).item())
            e0.record()
            for _ in range(3):
                r.copy_(raw0)
                L.nsr_coarse_refine(P(packed), P(z0), n, S, P(pc), P(r), P(ws), ws.numel(), None)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
        with torch.no_grad():
            got = nsr.render_rays(packed, nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map']
        d = (got - ref).abs().max(-1).values
        print(f'phi {phi:6.1f}  acc0 limit {str(lim):5s} sigma_hi {shi:4.0f}: {cnt:7d} points, stage {ms:6.3f} ms (incl. a 41 MB copy); rays beyond 1e-3: {int((d > 1e-3).sum())}, max {float(d.max()):.2e}', flush=True)
L.nsr_set_coarse_refine(1); L.nsr_set_coarse_refine_limit(0.9); L.nsr_set_coarse_refine_sigma(10.)
