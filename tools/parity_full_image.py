"""Whole-image parity, every one of the 160 000 rays: this renderer against the UNMODIFIED reference (oracle/_ref or /root/reference) run on the
host CPU and as eager PyTorch on this GPU, and the reference against itself (CPU vs CUDA).  Hierarchical sampling is discontinuous in
the last bits of the coarse weights (RH:239, `denom < 1e-5`): a sample can move by a coarse bin, and on a ray grazing the surface that
moves the pixel.  The table shows how many pixels each pair of renderers disagrees on beyond 1e-3 -- including the reference with itself."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O  # noqa: E402
import ref_import  # noqa: E402
import neural_sim_nerf_b200 as nsr  # noqa: E402

RN, RH = ref_import.load()
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
sds = [{k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)} for p in ('coarse/', 'fine/')]
H = W = 400
K = O.YCBV_K_400
phi = float(os.environ.get('PHI', 22.5))
pose = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
ro, rd = O.get_rays(H, W, K, pose)                                   # CPU torch: the reference's own get_rays arithmetic (pinned in tests)
rays_cpu = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)
out = {}
# reference on the CPU
torch.set_num_threads(16)
kw = ref_import.render_kwargs(sds[0], sds[1], O.YCBV_NEAR, O.YCBV_FAR)
t0 = time.perf_counter()
with ref_import.cpu_shim(), torch.no_grad():
    out['reference CPU'] = RN.render(H, W, K, chunk=512, rays=rays_cpu, **kw)[0]
print(f'reference CPU: {time.perf_counter() - t0:.1f} s', flush=True)
# reference eager on this GPU, same rays
kw['network_fn'].cuda()
kw['network_fine'].cuda()
with torch.device('cuda'), torch.no_grad():
    out['reference CUDA'] = RN.render(H, W, K, chunk=32768, rays=rays_cpu.cuda(), **kw)[0].cpu()
# this renderer, same rays and kernel-made rays
nets = []
for sd in sds:
    m = nsr.NeRF()
    m.load_state_dict(sd)
    nets.append(m.cuda())
nkw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
           near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
with torch.no_grad():
    out['this (same rays)'] = nsr.render(H, W, K, chunk=1 << 20, rays=rays_cpu.cuda(), **nkw)[0].cpu()
    out['this (rays from nsr_make_rays)'] = nsr.render(H, W, K, chunk=1 << 20, c2w=pose, **nkw)[0].reshape(-1, 3).cpu()
names = list(out)
print(f'phi = {phi}: pixels (of {H * W}) whose rgb differs by more than 1e-3 / 1e-2, and the largest difference')
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        d = (out[names[i]] - out[names[j]]).abs().max(-1).values
        print(f'  {names[i]:32s} vs {names[j]:32s}: {int((d > 1e-3).sum()):4d} / {int((d > 1e-2).sum()):3d}   max {float(d.max()):.3e}   median {float(d.median()):.1e}')
