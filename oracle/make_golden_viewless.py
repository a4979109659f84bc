"""Generate tests/golden/viewless_golden.npz by running the UNMODIFIED reference with use_viewdirs=False.  TEST INFRASTRUCTURE.

Build container only (needs /root/reference; oracle/ref_import.py).  The networks are the reference's own
NeRF(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=5, skips=[4], use_viewdirs=False) (RH:70-97, RN:263-278) loaded with
nerf_oracle.viewless_state_dict(tests/golden/wfit.npz): the fitted trunk, `output_linear` = [a linear colour head; alpha_linear; 0.01].
Every stored output comes from the reference's functions on CPU, fp32:

  inputs   ro, rd [n,3] (20x20 grid of the 400x400 YCB-V camera, pose theta=90, phi=22.5-180, r=1.01), near, far;
           c_output_w / c_output_b / f_output_w / f_output_b: the `output_linear` tensors used (so that the tests rebuild the very same networks)
  e2e      RN.render(rays=..., use_viewdirs=False, retraw=True): rgb_map, disp_map, acc_map, rgb0, disp0, acc0, z_std, raw [n/8,192,5] (every 8th ray)
  points   pts [m,16,3] and RN.run_network(pts, None, network_fine, embed_fn, None) -> raw_pts [m,16,5]

Usage:  python oracle/make_golden_viewless.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nerf_oracle as O  # noqa: E402  (camera constants, pose helper, the state-dict derivation)
import ref_import  # noqa: E402
from make_golden import load_wfit  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def main():
    RN, RH = ref_import.load()
    torch.autograd.set_detect_anomaly(False)
    sdc, sdf = (O.viewless_state_dict(sd, output_ch=5) for sd in load_wfit())
    coarse, fine, query = ref_import.build_models(sdc, sdf)
    assert not coarse.use_viewdirs and coarse.input_ch_views == 0
    H = W = 400
    K = torch.tensor(O.YCBV_K_400)
    c2w = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    near, far = O.YCBV_NEAR, O.YCBV_FAR
    out = {'near': np.float32(near), 'far': np.float32(far),
           'c_output_w': sdc['output_linear.weight'].numpy(), 'c_output_b': sdc['output_linear.bias'].numpy(),
           'f_output_w': sdf['output_linear.weight'].numpy(), 'f_output_b': sdf['output_linear.bias'].numpy()}
    with torch.no_grad():
        rays_o, rays_d = RH.get_rays(H, W, K, c2w)
        ii = torch.arange(10, 400, 20)
        sel = (ii[:, None] * W + ii[None, :]).reshape(-1)
        ro, rd = rays_o.reshape(-1, 3)[sel], rays_d.reshape(-1, 3)[sel]
        out['ro'], out['rd'] = ro.numpy(), rd.numpy()
        kw = ref_import.render_kwargs(sdc, sdf, near, far)
        assert kw['use_viewdirs'] is False
        e2e = RN.render(H, W, K, chunk=512, rays=torch.stack([ro, rd], 0), retraw=True, **kw)
        out['rgb_map'], out['disp_map'], out['acc_map'] = (t.numpy() for t in e2e[:3])
        for k, v in e2e[3].items():
            out[k] = v.numpy()
        out['raw'] = out['raw'][::8]                      # every 8th ray keeps the fixture small
        z = torch.linspace(near, far, 16)
        pts = ro[::9, None, :] + rd[::9, None, :] * z[None, :, None]
        out['pts'] = pts.numpy()
        out['raw_pts'] = query(pts, None, fine).numpy()
    assert out['raw'].shape[-1] == 5 and out['raw_pts'].shape[-1] == 5
    assert (out['acc_map'] > 0.9).any() and (out['acc_map'] < 0.1).any()
    path = os.path.join(ROOT, 'tests', 'golden', 'viewless_golden.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: getattr(v, 'shape', ()) for k, v in out.items()})


if __name__ == '__main__':
    main()
