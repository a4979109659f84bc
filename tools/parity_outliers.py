"""Why do a handful of rays per image differ from the reference by more than 1e-3?  Finds them (against the reference's eager CUDA path, same
rays), then compares the coarse pass (sigma, weights) and the merged fine depths of both renderers on those rays."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O, ref_import
import neural_sim_nerf_b200 as nsr
RN, RH = ref_import.load()
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
sds = [{k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)} for p in ('coarse/', 'fine/')]
H = W = 400; K = O.YCBV_K_400
phi = float(os.environ.get('PHI', 22.5))
pose = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
ro, rd = O.get_rays(H, W, K, pose)
rays2 = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0).cuda()
kw = ref_import.render_kwargs(sds[0], sds[1], O.YCBV_NEAR, O.YCBV_FAR)
kw['network_fn'].cuda(); kw['network_fine'].cuda()
nets = []
for sd in sds:
    m = nsr.NeRF(); m.load_state_dict(sd); nets.append(m.cuda())
nkw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
           near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
with torch.device('cuda'), torch.no_grad():
    ref = RN.render(H, W, K, chunk=32768, rays=rays2, **kw)[0]
    ours = nsr.render(H, W, K, chunk=1 << 20, rays=rays2, **nkw)[0]
d = (ref - ours).abs().max(-1).values
bad = torch.nonzero(d > 1e-3).reshape(-1)
print(f'phi {phi}: {bad.numel()} rays differ by more than 1e-3: {[(int(i), round(float(d[i]), 4)) for i in bad]}')
sel = bad[:8]
packed = O.pack_rays(ro.reshape(-1, 3)[sel.cpu()], rd.reshape(-1, 3)[sel.cpu()], O.YCBV_NEAR, O.YCBV_FAR)
# reference internals on these rays (oracle restatement, fp32 on the GPU = same kernels as the reference's eager path)
sdc = {k: v.cuda() for k, v in sds[0].items()}; sdf = {k: v.cuda() for k, v in sds[1].items()}
with torch.device('cuda'), torch.no_grad():
    r = O.render_rays(packed.cuda(), sdc, sdf, 64, 128, return_internals=True)['_internals']
# ours: stage by stage through the C ABI
L = nsr.lib(); P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
n = sel.numel(); pk = packed.cuda().contiguous()
pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
z0 = r['z0'].contiguous(); raw0 = torch.empty(n, 64, 4, device='cuda')
assert L.nsr_mlp_forward(P(pk), P(z0), n, 64, P(pc), 0, P(raw0), None) == 0
w0 = torch.empty(n, 64, device='cuda'); o3 = torch.empty(n, 3, device='cuda'); o1 = torch.empty(n, device='cuda'); o2 = torch.empty(n, device='cuda')
assert L.nsr_raw2outputs(P(raw0), P(z0), P(pk[:, 3:6].contiguous()), 3, n, 64, 0, P(o3), P(o1), P(o2), P(w0), None, None) == 0
zf = torch.empty(n, 192, device='cuda')
assert L.nsr_resample_merge(P(z0), P(w0), n, 64, 128, None, P(zf), None, None, None) == 0
torch.cuda.synchronize()
for i in range(n):
    moved = torch.nonzero((zf[i] - r['z1'][i]).abs() > 1e-6).reshape(-1)
    print(f'--- ray {int(sel[i])}: pixel diff {float(d[sel[i]]):.4f}; acc0 ref {float(r["weights0"][i].sum()):.6f}; {moved.numel()} of 192 merged depths differ, '
          f'largest move {float((zf[i] - r["z1"][i]).abs().max()):.5f} (coarse spacing {float(z0[i, 1] - z0[i, 0]):.5f})')
    ds = (raw0[i, :, 3] - r['raw0'][i, :, 3]).abs()
    near0 = torch.nonzero(r['raw0'][i, :, 3].abs() < 0.05).reshape(-1)
    print('    coarse sigma: max |ours - ref|', float(ds.max()), '; samples with |sigma_ref| < 0.05:', [(int(k), float(r['raw0'][i, k, 3]), float(raw0[i, k, 3])) for k in near0])
    dw = (w0[i] - r['weights0'][i]).abs()
    k = int(torch.argmax(dw))
    print(f'    coarse weights: max |ours - ref| {float(dw.max()):.3e} at sample {k} (ref {float(r["weights0"][i, k]):.3e}, ours {float(w0[i, k]):.3e})')
    small = torch.nonzero((r['weights0'][i] > 0) & (r['weights0'][i] < 1e-7)).reshape(-1)
    print('    ref weights in (0, 1e-7):', [(int(k), float(r['weights0'][i, k]), float(w0[i, k])) for k in small][:6])
