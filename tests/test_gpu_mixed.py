"""GPU parity of the mixed-precision forward mode (NSR_FLAG_MIXED_F8 / NSR_PRECISION='mixed'): fp16 hi/lo split for the
first three layers, fp16 main term + e4m3 residual products for the later ones (DESIGN.md "precision").

On the fitted scene the bar is the same as for the default mode: every rendered map within 1e-3 * max(1, |ref|) of the
fp32 reference (CPU emulation of the exact arithmetic on 12 800 rays predicted <= 2.4e-4; measured 3.7e-4).  On
adversarial random networks it misses that bar (2.6e-3), so the mode is opt-in; the tests print what the GPU delivers.
"""
import numpy as np
import pytest
import torch

import nerf_oracle as O
from test_gpu_parity import C, TOL_MAP, TOL_RAW, assert_close, assert_mostly_close, camera_rays, module_from_sd, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nsr():
    import neural_sim_nerf_b200 as m
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return m


@pytest.fixture(scope='module')
def nets(nsr, wfit):
    return module_from_sd(nsr, wfit[0]), module_from_sd(nsr, wfit[1])


@pytest.fixture()
def mixed(nsr):
    prev = nsr.run_nerf.PRECISION
    nsr.set_precision('mixed')
    yield
    nsr.set_precision(prev)


def test_mixed_mlp_raw_vs_golden(nsr, golden, nets, mixed):
    rays = C(golden['rays'])
    for zkey, net, key in (('z0', nets[0], 'raw0'), ('z1', nets[1], 'raw1')):
        z = C(golden[zkey])
        pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
        raw = nsr.run_network(pts, rays[:, 8:11].contiguous(), net)
        mx = assert_close(raw, golden[key], 4 * TOL_RAW, key)     # sigma_raw reaches 60: 4e-3 of it is 7e-5 of alpha at most
        print(f'mixed {key}: max err {mx:.3e}')


def test_mixed_render_vs_golden(nsr, golden, nets, mixed):
    rays = C(golden['rays'])
    with torch.no_grad():
        r = nsr.render_rays(rays, nets[0], None, 64, N_importance=128, network_fine=nets[1])
    for k in ('rgb_map', 'acc_map', 'rgb0', 'acc0'):
        mx = assert_close(r[k], golden['e2e_' + k], TOL_MAP, k)
        print(f'mixed {k}: max err {mx:.3e}')
    assert_mostly_close(r['z_std'], golden['e2e_z_std'], TOL_MAP, 'z_std')


@pytest.mark.parametrize('phi', [22.5, 112.5, 202.5, 292.5])
def test_mixed_render_vs_oracle_seeded(nsr, wfit, nets, mixed, phi):
    rays = camera_rays(40, phi)
    with torch.no_grad():
        ref = O.render_rays(rays, wfit[0], wfit[1], 64, 128)
        got = nsr.render_rays(rays.cuda(), nets[0], None, 64, N_importance=128, network_fine=nets[1])
    worst = 0.0
    for k in ('rgb_map', 'acc_map', 'rgb0', 'acc0'):
        worst = max(worst, assert_close(got[k], ref[k], TOL_MAP, f'{k} phi={phi}'))
    print(f'mixed phi={phi}: max map err {worst:.3e} (bar {TOL_MAP})')
    assert bool((torch.isnan(got['disp_map'].cpu()) == torch.isnan(ref['disp_map'])).all())
    assert_mostly_close(got['z_std'], ref['z_std'], TOL_MAP, f'z_std phi={phi}')


TOL_MIXED_ADVERSARIAL = 5e-3


def test_mixed_scaled_random_weights(nsr, mixed):
    """Different weight magnitudes per layer: the per-layer power-of-two scaling of the packed operands must follow.
    Measured on B200: 2.6e-3 on these 3x-scaled random (chaotic) networks -- the e4m3 residual products leave ~2^-15 per
    product, the default fp16x3 mode ~2^-22 -- which is why 'mixed' is opt-in and never the default or the headline:
    it holds 1e-3 on the fitted scene (tests above) but not on every network.  The bound here is 5e-3."""
    sdc, sdf = O.random_state_dict(21, scale=3.0), O.random_state_dict(22, scale=3.0)
    for sd in (sdc, sdf):
        sd['alpha_linear.bias'] += 2.0
        # move magnitude between consecutive layers (ReLU is positively homogeneous): same function, very different ranges
        for i, f in ((1, 64.0), (2, 1 / 64.0), (6, 1 / 512.0), (7, 512.0)):
            sd[f'pts_linears.{i}.weight'] = sd[f'pts_linears.{i}.weight'] * f
            sd[f'pts_linears.{i}.bias'] = sd[f'pts_linears.{i}.bias'] * (f if i in (1, 6) else 1.0)
    rays = camera_rays(12, 22.5)
    with torch.no_grad():
        ref = O.render_rays(rays, sdc, sdf, 64, 128)
        got = nsr.render_rays(rays.cuda(), module_from_sd(nsr, sdc), None, 64, N_importance=128,
                              network_fine=module_from_sd(nsr, sdf))
    for k in ('rgb_map', 'acc_map', 'rgb0', 'acc0'):
        mx = assert_close(got[k], ref[k], TOL_MIXED_ADVERSARIAL, k)
        print(f'mixed scaled-random {k}: max err {mx:.3e}')


def test_mixed_is_deterministic_and_chunk_invariant(nsr, nets, mixed):
    rays = camera_rays(30, 67.5).cuda()
    with torch.no_grad():
        a = nsr.render_rays(rays, nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map']
        b = nsr.render_rays(rays, nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map']
        parts = [nsr.render_rays(rays[i:i + 333], nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map']
                 for i in range(0, rays.shape[0], 333)]
    assert torch.equal(a, b)
    assert torch.equal(torch.cat(parts, 0), a)


def test_mixed_sits_between_the_other_modes(nsr, wfit, nets):
    rays = camera_rays(40, 22.5)
    errs = {}
    with torch.no_grad():
        ref = O.render_rays(rays, wfit[0], wfit[1], 64, 128)
        prev = nsr.run_nerf.PRECISION
        try:
            for mode in ('fp16x3', 'mixed', 'fp16'):
                nsr.set_precision(mode)
                got = nsr.render_rays(rays.cuda(), nets[0], None, 64, N_importance=128, network_fine=nets[1])
                errs[mode] = relerr(got['rgb_map'], ref['rgb_map'])[0].max()
        finally:
            nsr.set_precision(prev)
    print('rgb_map max err by mode:', {k: f'{v:.2e}' for k, v in errs.items()})
    assert errs['fp16x3'] <= errs['mixed'] * 1.5 + 1e-5 and errs['mixed'] < errs['fp16']
    assert errs['mixed'] <= TOL_MAP


def test_gradient_passes_ignore_mixed(nsr, nets, mixed):
    """Passes that carry gradient run the fp16 hi/lo split regardless of NSR_PRECISION (the backward kernel recomputes
    activations in that arithmetic)."""
    rays = camera_rays(6, 22.5).cuda().requires_grad_(True)
    r = nsr.render_rays(rays, nets[0], None, 64, N_importance=128, network_fine=nets[1])
    nsr.set_precision('fp16x3')
    with torch.no_grad():
        ref = nsr.render_rays(rays.detach(), nets[0], None, 64, N_importance=128, network_fine=nets[1])
    assert torch.equal(r['rgb_map'].detach(), ref['rgb_map'])
    g, = torch.autograd.grad(r['rgb_map'].sum(), rays)
    assert torch.isfinite(g).all()
