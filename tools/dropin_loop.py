"""The two-line binding at work: the UNMODIFIED reference image loop render_path_grad (RN:126-210; 313 chunks of 512 rays per 400x400 image,
each a render() + two autograd.grad calls, a .cpu() and an empty_cache()) with nsr.install(RN) rebinding render / get_rays, against this
package's own render_path_grad (one launch sequence per image).  GPU box: the reference comes from oracle/_ref.  Prints seconds per image."""
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
nets = []
for pre in ('coarse/', 'fine/'):
    m = nsr.NeRF()
    m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)})
    nets.append(m.cuda().requires_grad_(False))
H = W = 400
K = O.YCBV_K_400
hwf = [H, W, K[0][0]]
kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
          near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
n_img = int(os.environ.get('N_IMG', 2))
psi = torch.full((8,), 0.02)
psi[4] = 0.86
grad_E = [{'grad_E': torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(i)) * 1e-3} for i in range(n_img)]


def poses_for():
    prob = torch.softmax(psi.cuda() / 0.25, 0).requires_grad_()
    _, log = nsr.sample_pose_nograd(prob.detach(), n_img, 0.1, seed=0, device='cuda')
    return prob, nsr.sample_pose(prob, n_img, 0.1, log)


res = {}
nsr.install(RN)                                  # MAIN:35 star-imports these names; RN:168 / RN:148 look them up in RN's globals
for name, fn in (('reference loop + install()', RN.render_path_grad), ('nsr.render_path_grad', nsr.render_path_grad)):
    for rep in range(2):                         # second pass is the timed one
        prob, poses = poses_for()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rgbs, dl = fn(prob, poses, hwf, K, 512, grad_E, kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    g = torch.stack([d.reshape(-1) for d in dl], 0).mean(0)
    res[name] = (dt / n_img, g, rgbs)
    print(f'{name:28s}: {dt / n_img * 1e3:8.1f} ms per image ({len(dl)} gradient entries), mean dL/dpsi norm {float(g.norm()):.4e}', flush=True)
a, b = res['reference loop + install()'], res['nsr.render_path_grad']
print(f'same mean gradient: max diff {float((a[1] - b[1]).abs().max()):.3e} of {float(b[1].abs().max()):.3e};  same pixels: max diff {float(np.abs(a[2] - b[2]).max()):.3e}')
