"""CPU emulation of reduced-precision GEMMs in the BACKWARD (data-gradient) pass of the NeRF MLP (research aid, not a test).

The backward kernel runs every product as the fp16 hi/lo split (3 MMAs), like the forward.  Gradients feed a stochastic outer
optimisation (the psi update), so their precision need may be lower than the 1e-3 bar on rendered maps.  This script measures
what ONE fp16 x fp16 MMA per product (operands rounded to fp16, fp32 accumulate, rows pre-scaled by a power of two as in the
kernel) or a 2-MMA variant would do to dL/d(rays), against exact fp32 autograd, on the fitted scene and on scaled random networks.

  python tests/precision_emulation_backward.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np
import torch

import nerf_oracle as O
from precision_emulation import camera_rays


def h16(t):
    return t.half().float()


def prod(g, w, mode):
    """g [P, out] @ w [out, in] with the operand roundings of `mode`."""
    if mode == 'exact':
        return g @ w
    if mode == 'fp16':
        return h16(g) @ h16(w)
    if mode == 'g_split':      # g_hi W_hi + g_lo W_hi          (2 MMAs: weights rounded, gradients exact to 2^-22)
        return g @ h16(w)
    if mode == 'w_split':      # g_hi W_hi + g_hi W_lo          (2 MMAs: gradients rounded)
        return h16(g) @ w
    raise ValueError(mode)


def mlp_backward(x, sd, d_raw, mode):
    """Manual backward of O.mlp_forward w.r.t. its input x [P,90]; forward in exact fp32; every backward product through prod()."""
    W = lambda n: sd[n + '.weight']
    B = lambda n: sd[n + '.bias']
    lin = torch.nn.functional.linear
    enc_xyz, enc_dir = x[..., :63], x[..., 63:]
    hs, h = [], enc_xyz
    for i in range(8):
        z = lin(h, W(f'pts_linears.{i}'), B(f'pts_linears.{i}'))
        h = torch.relu(z)
        hs.append(h)
        if i == 4:
            h = torch.cat([enc_xyz, h], -1)
    feat = lin(h, W('feature_linear'), B('feature_linear'))
    hv = torch.relu(lin(torch.cat([feat, enc_dir], -1), W('views_linears.0'), B('views_linears.0')))
    # per-row power-of-two scale (the kernel's): largest |d_raw| component -> [1, 2)
    gmax = d_raw.abs().amax(-1, keepdim=True).clamp(min=1e-30)
    scale = torch.exp2(torch.floor(torch.log2(gmax)))
    g_raw = d_raw / scale
    g_rgb, g_sig = g_raw[:, :3], g_raw[:, 3:4]
    g_hv = (g_rgb @ W('rgb_linear')) * (hv > 0)                         # CUDA cores in the kernel: exact
    g_cat = prod(g_hv, W('views_linears.0'), mode)                       # [P, 256 + 27]
    g_feat, g_dir = g_cat[:, :256], g_cat[:, 256:]
    g_h = prod(g_feat, W('feature_linear'), mode) + g_sig * W('alpha_linear')   # alpha head added on CUDA cores
    g_xyz = torch.zeros_like(enc_xyz)
    for i in range(7, -1, -1):
        g_z = g_h * (hs[i] > 0)
        g_in = prod(g_z, W(f'pts_linears.{i}'), mode)
        if i == 5:                                                       # input was cat[enc_xyz, h4]
            g_xyz = g_xyz + g_in[:, :63]
            g_h = g_in[:, 63:]
        elif i == 0:
            g_xyz = g_xyz + g_in
        else:
            g_h = g_in
    return torch.cat([g_xyz, g_dir], -1) * scale


def ray_grads(rays, sdc, sdf, g, mode):
    """dL/d(ray_batch) of the fine pass given dL/drgb_map = g, MLP data gradient through mlp_backward(mode)."""
    with torch.no_grad():
        ref = O.render_rays(rays, sdc, sdf, 64, 128, return_internals=True)
    z = ref['_internals']['z1']
    r = rays.clone().requires_grad_(True)
    pts = r[:, None, 0:3] + r[:, None, 3:6] * z[:, :, None]
    flat = pts.reshape(-1, 3)
    dirs = r[:, None, 8:11].expand(pts.shape).reshape(-1, 3)
    x = torch.cat([O.embed(flat, 10), O.embed(dirs, 4)], -1)
    x_d = x.detach().requires_grad_(True)
    raw = O.mlp_forward(x_d, sdf).reshape(rays.shape[0], -1, 4)
    raw_d = raw.detach().requires_grad_(True)
    rgb = O.raw2outputs(raw_d, z, r[:, 3:6])[0]
    d_raw, d_r_comp = torch.autograd.grad(rgb, [raw_d, r], grad_outputs=g, allow_unused=True)
    with torch.no_grad():
        d_x = mlp_backward(x_d.detach(), sdf, d_raw.reshape(-1, 4), mode)
    d_r_mlp, = torch.autograd.grad(x, r, grad_outputs=d_x)
    return d_r_mlp + (d_r_comp if d_r_comp is not None else 0)


def main():
    torch.set_num_threads(min(16, os.cpu_count() or 8))
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sdc = {k[len('coarse/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('coarse/')}
    sdf = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
    nets = {'wfit': (sdc, sdf), 'rand3': (O.random_state_dict(21, scale=3.0), O.random_state_dict(22, scale=3.0))}
    for sd in nets['rand3']:
        sd['alpha_linear.bias'] += 2.0
    n_side = int(os.environ.get('N_SIDE', 24))
    for name, (a, b) in nets.items():
        for phi in (22.5, 202.5):
            rays = camera_rays(n_side, phi)
            g = torch.randn(rays.shape[0], 3, generator=torch.Generator().manual_seed(1))
            cols = [0, 1, 2, 3, 4, 5, 8, 9, 10]     # near / far (6:8) are constants of the reference's graph (batch_rays holds o, d only)
            exact = ray_grads(rays, a, b, g, 'exact')[:, cols]
            # sanity: the manual backward in exact mode is autograd
            r = rays.clone().requires_grad_(True)
            ref = O.render_rays(r, a, b, 64, 128)
            auto, = torch.autograd.grad(ref['rgb_map'], r, grad_outputs=g)
            auto = auto[:, cols]
            sc = float(auto.abs().max())
            line = f'{name} phi={phi}: max|grad| {sc:.3e}  manual-vs-autograd {float((exact - auto).abs().max()) / sc:.1e}'
            for mode in ('fp16', 'g_split', 'w_split'):
                got = ray_grads(rays, a, b, g, mode)[:, cols]
                err = float((got - exact).abs().max()) / sc
                # what the outer loop sees: the image-summed gradient (dL/dc2w is a weighted sum over rays)
                agg = float((got.sum(0) - exact.sum(0)).abs().max() / exact.sum(0).abs().max().clamp(min=1e-30))
                line += f' | {mode}: per-ray {err:.1e}, summed {agg:.1e}'
            print(line, flush=True)


if __name__ == '__main__':
    main()
