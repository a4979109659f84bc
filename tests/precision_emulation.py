"""CPU emulation of reduced-precision GEMM steps inside the NeRF MLP (research aid behind DESIGN.md "precision"; not a test).

Each of the ten GEMM steps (pts_linears.0-7, feature_linear, views_linears.0) is evaluated either exactly in fp32 (stands for
the fp16 hi/lo split, whose dropped term is ~2^-22) or as ONE fp16 x fp16 product with fp32 accumulation (operands rounded to
fp16), and the rendered maps are compared with the exact oracle on rays of the fitted scene and of scaled random networks.

  python tests/precision_emulation.py            per-step sensitivities and a few combinations
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np
import torch

import nerf_oracle as O

lin = torch.nn.functional.linear


def rounded_linear(x, w, b, single):
    if single:
        return lin(x.half().float(), w.half().float(), b)
    return lin(x, w, b)


def make_mlp(single_steps):
    S = set(single_steps)

    def mlp(x, sd):
        enc_xyz, enc_dir = x[..., :63], x[..., 63:]
        h = enc_xyz
        for i in range(8):
            h = torch.relu(rounded_linear(h, sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias'], i in S))
            if i == 4:
                h = torch.cat([enc_xyz, h], -1)
        sigma = lin(h, sd['alpha_linear.weight'], sd['alpha_linear.bias'])
        feat = rounded_linear(h, sd['feature_linear.weight'], sd['feature_linear.bias'], 8 in S)
        h = torch.relu(rounded_linear(torch.cat([feat, enc_dir], -1), sd['views_linears.0.weight'], sd['views_linears.0.bias'], 9 in S))
        rgb = lin(h, sd['rgb_linear.weight'], sd['rgb_linear.bias'])
        return torch.cat([rgb, sigma], -1)
    return mlp


def camera_rays(n_side, phi):
    H = W = 400
    c2w = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    ii = torch.linspace(0, 399, n_side).long()
    sel = (ii[:, None] * W + ii[None, :]).reshape(-1)
    return O.pack_rays(ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel], O.YCBV_NEAR, O.YCBV_FAR)


def max_err(a, b):
    return float(((a - b).abs() / b.abs().clamp(min=1.0)).max())


def main():
    torch.set_num_threads(min(16, os.cpu_count() or 8))
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sdc = {k[len('coarse/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('coarse/')}
    sdf = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
    nets = {'wfit': (sdc, sdf), 'rand3': (O.random_state_dict(21, scale=3.0), O.random_state_dict(22, scale=3.0))}
    for sd in nets['rand3']:
        sd['alpha_linear.bias'] += 2.0
    n_side = int(os.environ.get('N_SIDE', 40))
    views = [float(v) for v in os.environ.get('VIEWS', '22.5,202.5').split(',')]
    configs = [('all single', list(range(10)))] + [(f'step {s} single', [s]) for s in range(10)] + \
              [('8+9 (views branch)', [8, 9]), ('0+8+9', [0, 8, 9])]
    if len(sys.argv) > 1:
        configs = [(a, [int(t) for t in a.split('+')]) for a in sys.argv[1:]]
    exact = O.mlp_forward
    refs = {}
    with torch.no_grad():
        for name, (a, b) in nets.items():
            for phi in views:
                refs[(name, phi)] = O.render_rays(camera_rays(n_side, phi), a, b, 64, 128)
        for label, steps in configs:
            t0 = time.time()
            O.mlp_forward = make_mlp(steps)
            out = {}
            for name, (a, b) in nets.items():
                worst = 0.0
                for phi in views:
                    got = O.render_rays(camera_rays(n_side, phi), a, b, 64, 128)
                    ref = refs[(name, phi)]
                    worst = max(worst, max(max_err(got[k], ref[k]) for k in ('rgb_map', 'acc_map', 'rgb0', 'acc0')))
                out[name] = worst
            O.mlp_forward = exact
            print(f'{label:22s} ' + '  '.join(f'{k}: {v:.2e}' for k, v in out.items()) + f'   ({time.time() - t0:.0f} s)', flush=True)


if __name__ == '__main__':
    main()
