"""Generate tests/golden/render_golden.npz by running the UNMODIFIED reference.  TEST INFRASTRUCTURE.

Runs only in the build container (needs /root/reference; see oracle/ref_import.py for the
import recipe).  Every array stored here is an output of the reference's own functions
(RN.render, RN.run_network, RN.raw2outputs, RH.sample_pdf, RH.get_rays) on CPU, fp32:

  inputs   rays [n,11]  (pixels on a 20x20 grid of the 400x400 YCB-V camera, pose theta=90, phi=22.5-180, r=1.01;
                         K / near / far from logs/nerfdata/nerf_traindata_info.json, LL:185-198)
           weights: tests/golden/wfit.npz (oracle/make_weights.py)
  stages   z0, raw0, weights0, rgb0/disp0/acc0, z_samples, z1, raw1, weights1, rgb_map/disp_map/acc_map, z_std
  e2e      e2e_* = RN.render(rays=...) outputs for the same rays (chunk=512)
  rays     getrays_o / getrays_d = RH.get_rays on a 12x10 image
  perturb  a second, stratified (perturb=1) run with explicit t_rand / u: p_* arrays

Usage:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nerf_oracle as O  # noqa: E402  (only for the camera constants / pose helper)
import ref_import  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def load_wfit():
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sdc = {k[len('coarse/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('coarse/')}
    sdf = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
    return sdc, sdf


def main():
    torch.manual_seed(0)
    RN, RH = ref_import.load()
    torch.autograd.set_detect_anomaly(False)
    sdc, sdf = load_wfit()
    coarse, fine, query = ref_import.build_models(sdc, sdf)
    H = W = 400
    K = torch.tensor(O.YCBV_K_400)
    c2w = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    out = {}
    with torch.no_grad():
        rays_o, rays_d = RH.get_rays(H, W, K, c2w)
        ii = torch.arange(10, 400, 20)
        sel = (ii[:, None] * W + ii[None, :]).reshape(-1)
        ro, rd = rays_o.reshape(-1, 3)[sel], rays_d.reshape(-1, 3)[sel]
        n = ro.shape[0]
        near, far = O.YCBV_NEAR, O.YCBV_FAR
        # --- end to end through the reference's public entry point
        kw = ref_import.render_kwargs(sdc, sdf, near, far)
        e2e = RN.render(H, W, K, chunk=512, rays=torch.stack([ro, rd], 0), retraw=True, **kw)
        out['e2e_rgb_map'], out['e2e_disp_map'], out['e2e_acc_map'] = (t.numpy() for t in e2e[:3])
        for k, v in e2e[3].items():
            out['e2e_' + k] = v.numpy()
        # --- stage by stage with the reference's functions (RN:433-495)
        vd = rd / torch.norm(rd, dim=-1, keepdim=True)
        rays = torch.cat([ro, rd, near * torch.ones(n, 1), far * torch.ones(n, 1), vd], -1).float()
        out['rays'] = rays.numpy()
        t_vals = torch.linspace(0., 1., steps=64)
        z0 = (near * (1. - t_vals) + far * t_vals).expand([n, 64]).contiguous()
        pts = ro[:, None, :] + rd[:, None, :] * z0[:, :, None]
        raw0 = query(pts, vd, coarse)
        rgb0, disp0, acc0, w0, depth0 = RN.raw2outputs(raw0, z0, rd, 0, False)
        z_mid = .5 * (z0[..., 1:] + z0[..., :-1])
        z_samples = RH.sample_pdf(z_mid, w0[..., 1:-1], 128, det=True)
        z1, _ = torch.sort(torch.cat([z0, z_samples], -1), -1)
        pts = ro[:, None, :] + rd[:, None, :] * z1[:, :, None]
        raw1 = query(pts, vd, fine)
        rgb1, disp1, acc1, w1, depth1 = RN.raw2outputs(raw1, z1, rd, 0, False)
        for k, v in dict(z0=z0, raw0=raw0, weights0=w0, rgb0=rgb0, disp0=disp0, acc0=acc0, depth0=depth0,
                         z_samples=z_samples, z1=z1, raw1=raw1, weights1=w1, rgb_map=rgb1, disp_map=disp1,
                         acc_map=acc1, depth_map=depth1, z_std=torch.std(z_samples, dim=-1, unbiased=False)).items():
            out[k] = v.numpy()
        # --- white background + lindisp variants of the compositor / depths
        out['wb_rgb_map'] = RN.raw2outputs(raw1, z1, rd, 0, True)[0].numpy()
        out['lindisp_z0'] = (1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)).expand([n, 64]).numpy().copy()
        # --- stratified run with explicit randoms (RN:447-461, RH:211): the reference draws them
        #     internally, so re-state those two lines around the reference's own functions
        g = torch.Generator().manual_seed(7)
        t_rand = torch.rand(n, 64, generator=g)
        u = torch.rand(n, 128, generator=g)
        mids = .5 * (z0[..., 1:] + z0[..., :-1])
        upper = torch.cat([mids, z0[..., -1:]], -1)
        lower = torch.cat([z0[..., :1], mids], -1)
        pz0 = lower + (upper - lower) * t_rand
        praw0 = query(ro[:, None, :] + rd[:, None, :] * pz0[:, :, None], vd, coarse)
        pw0 = RN.raw2outputs(praw0, pz0, rd, 0, False)[3]
        # sample_pdf(det=False) with our u: reproduce via the pytest hook-free path -> monkeypatch torch.rand
        _rand = torch.rand
        torch.rand = lambda *a, **k: u
        try:
            pzs = RH.sample_pdf(.5 * (pz0[..., 1:] + pz0[..., :-1]), pw0[..., 1:-1], 128, det=False)
        finally:
            torch.rand = _rand
        pz1, _ = torch.sort(torch.cat([pz0, pzs], -1), -1)
        for k, v in dict(p_t_rand=t_rand, p_u=u, p_z0=pz0, p_weights0=pw0, p_z_samples=pzs, p_z1=pz1).items():
            out[k] = v.numpy()
        # --- get_rays on a small image
        go, gd = RH.get_rays(10, 12, K / 33.0 + torch.eye(3) * 0, c2w)
        out['getrays_K'] = (K / 33.0).numpy()
        out['getrays_c2w'] = c2w.numpy()
        out['getrays_o'], out['getrays_d'] = go.numpy(), gd.numpy()
    dst = os.path.join(ROOT, 'tests', 'golden', 'render_golden.npz')
    np.savez_compressed(dst, **{k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()})
    print('wrote', dst, {k: v.shape for k, v in out.items()})
    hit = (out['acc_map'] > 0.5).mean()
    print(f'rays hitting the object: {hit:.2%}; NaN disparity rays: {np.isnan(out["disp_map"]).mean():.2%}')


if __name__ == '__main__':
    main()
