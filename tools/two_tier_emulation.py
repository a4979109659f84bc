"""CPU emulation of the two-tier precision scheme (research aid behind DESIGN.md "precision"; not a test).

Tier 1: every sample point is evaluated with ONE fp16 MMA per product (operands rounded to fp16, fp32 accumulate).
Tier 2: rays whose first-order error bound exceeds a budget are re-evaluated exactly (stands for fp16x3, error ~2^-22).
The bound is computed from tier-1 outputs only: per sample |d alpha| from sigma +- tau_sigma (exact, handles the ReLU kink),
pushed through the compositing derivative T_i (c_i - C_back_{i+1}), plus w_i * tau_c for the colours.

  python tools/two_tier_emulation.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np
import torch

import nerf_oracle as O
from precision_emulation import make_mlp, camera_rays

single_mlp = make_mlp(list(range(10)))
exact_mlp = O.mlp_forward


def run_net(pts, viewdirs, sd, single):
    O.mlp_forward = single_mlp if single else exact_mlp
    try:
        return O.run_network(pts, viewdirs, sd)
    finally:
        O.mlp_forward = exact_mlp


def alpha_of(sig, dists):
    return 1. - torch.exp(-torch.relu(sig) * dists)


def sensitivity(raw, z, rays_d, tau_s, tau_c):
    """first-order bounds on |d rgb_map|, |d acc_map|, |d (depth/acc)| / (depth/acc) per ray from tier-1 outputs"""
    dists = z[..., 1:] - z[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1) * torch.norm(rays_d[..., None, :], dim=-1)
    sig = raw[..., 3]
    c = torch.sigmoid(raw[..., :3])
    a = alpha_of(sig, dists)
    da = torch.maximum((alpha_of(sig + tau_s, dists) - a).abs(), (alpha_of(sig - tau_s, dists) - a).abs())
    one_m = 1. - a + 1e-10
    T = torch.cumprod(torch.cat([torch.ones_like(a[:, :1]), one_m], -1), -1)[:, :-1]
    w = a * T
    n, S = a.shape
    # back-to-front composites behind each sample
    Cb = torch.zeros(n, S + 1, 3)
    Ab = torch.zeros(n, S + 1)
    Db = torch.zeros(n, S + 1)
    for i in range(S - 1, -1, -1):
        Cb[:, i] = a[:, i, None] * c[:, i] + (1 - a[:, i, None]) * Cb[:, i + 1]
        Ab[:, i] = a[:, i] + (1 - a[:, i]) * Ab[:, i + 1]
        Db[:, i] = a[:, i] * z[:, i] + (1 - a[:, i]) * Db[:, i + 1]
    d_rgb = (T[..., None] * (c - Cb[:, 1:]).abs() * da[..., None]).sum(1).amax(-1)
    # colour error: d sigmoid <= 0.25 tau_c
    d_rgb = d_rgb + (w * 0.25 * tau_c).sum(1)
    d_acc = (T * (1 - Ab[:, 1:]) * da).sum(1)
    acc = w.sum(1)
    depth = (w * z).sum(1)
    zbar = depth / acc.clamp(min=1e-30)
    d_zbar = (T * ((z - Db[:, 1:]) - zbar[:, None] * (1 - Ab[:, 1:])).abs() * da).sum(1) / acc.clamp(min=1e-30)
    rel_disp = d_zbar / zbar.clamp(min=1e-30)
    maybe_hit = (da > 0).any(1)           # some alpha could be non-zero
    return d_rgb, d_acc, rel_disp, acc, maybe_hit


def rel_err(a, b):
    e = (a - b).abs() / b.abs().clamp(min=1.0)
    return e.reshape(e.shape[0], -1).amax(1)


def main():
    torch.set_num_threads(min(16, os.cpu_count() or 8))
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sdc = {k[len('coarse/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('coarse/')}
    sdf = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
    nets = {'wfit': (sdc, sdf)}
    if os.environ.get('RAND3', '1') == '1':
        r3 = (O.random_state_dict(21, scale=3.0), O.random_state_dict(22, scale=3.0))
        for sd in r3:
            sd['alpha_linear.bias'] += 2.0
        nets['rand3'] = r3
    n_side = int(os.environ.get('N_SIDE', 64))
    views = [float(v) for v in os.environ.get('VIEWS', '22.5,202.5').split(',')]
    with torch.no_grad():
        for name, (a, b) in nets.items():
            for phi in views:
                rays = camera_rays(n_side, phi)
                n = rays.shape[0]
                ro, rd, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 8:11]
                ref = O.render_rays(rays, a, b, 64, 128, return_internals=True)
                I = ref['_internals']
                # ---- coarse pass, both tiers on the same points
                z0 = I['z0']
                pts0 = ro[:, None] + rd[:, None] * z0[..., None]
                raw0_s = run_net(pts0, vd, a, True)
                raw0_e = I['raw0']
                dsig = (raw0_s[..., 3] - raw0_e[..., 3]).abs()
                dcol = (raw0_s[..., :3] - raw0_e[..., :3]).abs().amax(-1)
                print(f'[{name} phi={phi}] coarse raw: |dsigma| max {dsig.max():.3e} p99.9 {dsig.flatten().quantile(0.999):.3e} mean {dsig.mean():.3e}; '
                      f'|dcol| max {dcol.max():.3e} p99.9 {dcol.flatten().quantile(0.999):.3e}; sigma range [{raw0_e[..., 3].min():.1f}, {raw0_e[..., 3].max():.1f}]')
                # relative to local gain?
                out0_s = O.raw2outputs(raw0_s, z0, rd)
                out0_e = O.raw2outputs(raw0_e, z0, rd)
                e_rgb0 = rel_err(out0_s[0], out0_e[0])
                e_acc0 = rel_err(out0_s[2], out0_e[2])
                print(f'   coarse single-pass: rgb0 err max {e_rgb0.max():.3e}, acc0 err max {e_acc0.max():.3e}, rays > 1e-3: {(torch.maximum(e_rgb0, e_acc0) > 1e-3).float().mean():.4f}')
                # ---- fine pass on the EXACT depths, both tiers
                z1 = I['z1']
                pts1 = ro[:, None] + rd[:, None] * z1[..., None]
                raw1_s = run_net(pts1, vd, b, True)
                raw1_e = I['raw1']
                dsig1 = (raw1_s[..., 3] - raw1_e[..., 3]).abs()
                dcol1 = (raw1_s[..., :3] - raw1_e[..., :3]).abs().amax(-1)
                print(f'   fine raw: |dsigma| max {dsig1.max():.3e} p99.9 {dsig1.flatten().quantile(0.999):.3e} mean {dsig1.mean():.3e}; |dcol| max {dcol1.max():.3e} p99.9 {dcol1.flatten().quantile(0.999):.3e}')
                out1_s = O.raw2outputs(raw1_s, z1, rd)
                out1_e = O.raw2outputs(raw1_e, z1, rd)
                e_rgb = rel_err(out1_s[0], out1_e[0])
                e_acc = rel_err(out1_s[2], out1_e[2])
                dn = torch.isnan(out1_s[1]) != torch.isnan(out1_e[1])
                e_disp = rel_err(torch.nan_to_num(out1_s[1]), torch.nan_to_num(out1_e[1]))
                e_all = torch.maximum(torch.maximum(e_rgb, e_acc), e_disp)
                print(f'   fine single-pass (exact depths): rgb err max {e_rgb.max():.3e} acc {e_acc.max():.3e} disp {e_disp.max():.3e} nan-mismatch {int(dn.sum())}; rays > 1e-3: {(e_all > 1e-3).float().mean():.4f}  > 3e-4: {(e_all > 3e-4).float().mean():.4f}')
                # ---- criterion sweep
                for tau_s, tau_c in ((0.05, 0.01), (0.2, 0.02), (0.5, 0.05), (1.0, 0.1), (2.0, 0.2)):
                    d_rgb, d_acc, rel_disp, acc, maybe = sensitivity(raw1_s, z1, rd, tau_s, tau_c)
                    bound = torch.maximum(torch.maximum(d_rgb, d_acc), torch.where(acc > 0, rel_disp, torch.zeros_like(acc)))
                    for budget in (1e-3, 3e-4):
                        esc = bound > budget
                        missed = (~esc) & (e_all > 1e-3)
                        worst_unesc = e_all[~esc].max() if (~esc).any() else 0.
                        print(f'      tau_s={tau_s} tau_c={tau_c} budget={budget:.0e}: escalate {esc.float().mean():.4f}  missed(>1e-3) {int(missed.sum())}  worst unescalated err {worst_unesc:.3e}  nan-mismatch among unescalated {int((dn & ~esc).sum())}')


if __name__ == '__main__':
    main()
