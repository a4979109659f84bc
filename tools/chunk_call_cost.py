"""Cost of ONE 512-ray render() + the two autograd.grad calls of RN:168-181 through this package (the unit the unmodified reference loop
repeats 313 times per image), with and without the reference loop's own per-chunk torch.cuda.empty_cache() + .cpu() (RN:189-194)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O
import neural_sim_nerf_b200 as nsr
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
nets = []
for pre in ('coarse/', 'fine/'):
    m = nsr.NeRF(); m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}); nets.append(m.cuda().requires_grad_(False))
H = W = 400; K = O.YCBV_K_400
kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
          near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
pose = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4].cuda().requires_grad_(True)
ro, rd = nsr.get_rays(H, W, K, pose)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
g = torch.randn(H * W, 3, device='cuda') * 1e-3
for mode in ('plain', 'with empty_cache + .cpu() per chunk (as RN:189-194)'):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(0, 64 * 512, 512):
            batch = torch.stack([ro[i:i + 512], rd[i:i + 512]], 0)
            rgb, _, _, _ = nsr.render(H, W, K, chunk=512, rays=batch, retraw=True, **kw)
            dray = torch.autograd.grad(rgb, batch, grad_outputs=g[i:i + 512])
            dpose = torch.autograd.grad(batch, pose, grad_outputs=dray, retain_graph=True)
            if mode != 'plain':
                _ = dpose[0].cpu().detach(); _ = rgb.cpu().detach()
                torch.cuda.empty_cache()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{mode}: {dt / 64 * 1e3:.2f} ms per 512-ray chunk ({dt / 64 * 313:.2f} s per 400x400 image)')
