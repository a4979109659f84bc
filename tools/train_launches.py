"""Three fused training iterations (nsr_train_step, N_rand = 1024, 64+128 samples, perturb=1) for a per-kernel launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_launches.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np
import torch

import nerf_oracle as O
import neural_sim_nerf_b200 as nsr

z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
nets = []
for pre in ('coarse/', 'fine/'):
    m = nsr.NeRF()
    m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)})
    nets.append(m.cuda())
opt = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=5e-4)
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
n_rand = int(os.environ.get('N_RAND', 1024))
gen = torch.Generator(device='cuda').manual_seed(0)
target = torch.rand(n_rand, 3, device='cuda', generator=gen)
kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
          near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=1.0)
for it in range(3):
    sel = torch.randint(0, 160000, (n_rand,), device='cuda', generator=gen)
    r = rays[sel]
    out = nsr.train_step(torch.stack([r[:, 0:3], r[:, 3:6]], 0), target, opt, **kw)
torch.cuda.synchronize()
print('loss', float(out['loss']))
