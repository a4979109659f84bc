"""Fit the two NeRF MLPs to an analytic scene -> tests/golden/wfit.npz.  TEST INFRASTRUCTURE.

The pretrained YCB-V checkpoints the reference needs (logs/nerf_models/ycbvid*.tar,
MAIN:66) are not in the repo and there is no network, and
default-initialised weights give sigma <= 0 almost everywhere (an all-zero render).
This script regresses the reference MLP architecture (via the oracle's
mlp_forward, RH:99-122) onto a closed-form density/colour field so that parity is
exercised on weights with realistic magnitudes: sparse sigma (0 outside the object,
~60 inside), empty background rays (acc == 0 -> NaN disparity, RN:381), peaky
sample_pdf inputs, view-dependent colour.

Scene: a box (half-extent 0.07,0.10,0.05) united with a sphere (r=0.08, centred at
(0.05,0,0.06)) at the origin; cameras sit on the r~1.01 shell (LL:292-293).

Run:  python oracle/make_weights.py   (about 10 minutes on 8 cores)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nerf_oracle as O  # noqa: E402


def field(x, d):
    """x [P,3], d [P,3] unit -> (sigma_target_raw [P], rgb [P,3])."""
    box = torch.tensor([0.07, 0.10, 0.05])
    inside_box = (x.abs() < box).all(-1)
    inside_sph = ((x - torch.tensor([0.05, 0.0, 0.06])).norm(dim=-1) < 0.08)
    inside = inside_box | inside_sph
    sigma_raw = torch.where(inside, torch.tensor(60.0), torch.tensor(-6.0))
    base = 0.5 + 0.45 * torch.sin(x * torch.tensor([40.0, 55.0, 70.0]) + torch.tensor([0.0, 1.0, 2.0]))
    view = 0.25 * (d * torch.tensor([1.0, -1.0, 0.5])).sum(-1, keepdim=True)
    rgb = (base + view).clamp(0.02, 0.98)
    return sigma_raw, rgb


def sample_points(n, gen):
    """A third tight around the object, a third in a wider box, a third over the camera shell's ray span."""
    t = n // 3
    a = (torch.rand(t, 3, generator=gen) - 0.5) * 0.26          # tight around the object
    m = (torch.rand(t, 3, generator=gen) - 0.5) * 0.6
    b = (torch.rand(n - 2 * t, 3, generator=gen) - 0.5) * 2.2   # the whole span rays cover
    x = torch.cat([a, m, b], 0)
    d = torch.randn(n, 3, generator=gen)
    d = d / d.norm(dim=-1, keepdim=True)
    return x, d


def fit(seed, steps, batch=16384):
    gen = torch.Generator().manual_seed(seed)
    sd = {k: v.clone().requires_grad_(True) for k, v in O.random_state_dict(seed).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=5e-4)
    t0 = time.time()
    for it in range(steps):
        x, d = sample_points(batch, gen)
        sig_t, rgb_t = field(x, d)
        inp = torch.cat([O.embed(x, O.N_FREQ_XYZ), O.embed(d, O.N_FREQ_DIR)], -1)
        out = O.mlp_forward(inp, sd)
        loss = ((torch.sigmoid(out[:, :3]) - rgb_t) ** 2).mean() \
            + (((out[:, 3] - sig_t) / 30.0) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        for g in opt.param_groups:
            g['lr'] = 5e-4 * (0.1 ** (it / steps))
        if it % 100 == 0:
            print(f'seed {seed} it {it} loss {loss.item():.5f} ({time.time() - t0:.0f}s)', flush=True)
    return {k: v.detach() for k, v in sd.items()}


def main():
    torch.set_num_threads(os.cpu_count())
    out = {}
    for tag, seed, steps in (('coarse', 11, 1200), ('fine', 12, 1500)):
        sd = fit(seed, steps)
        for k, v in sd.items():
            out[f'{tag}/{k}'] = v.numpy()
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'wfit.npz')
    np.savez(dst, **out)
    print('wrote', dst)


if __name__ == '__main__':
    main()
