"""Pin the CPU oracle (oracle/nerf_oracle.py) to the reference: against the committed golden
vectors (made by running the unmodified reference, oracle/make_golden.py) everywhere, and
against the live reference where /root/reference exists (build container only)."""
import numpy as np
import pytest
import torch

import nerf_oracle as O
import ref_import


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, tol, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    assert np.array_equal(np.isnan(a), np.isnan(b)), f'{what}: NaN pattern differs'
    m = ~np.isnan(a)
    err = np.abs(a[m] - b[m]) / np.maximum(1.0, np.abs(b[m]))
    assert err.size == 0 or err.max() <= tol, f'{what}: max err {err.max():.3e} > {tol}'


def test_golden_has_hits_and_misses(golden):
    acc = golden['acc_map']
    assert (acc > 0.5).mean() > 0.05, 'golden rays never hit the object'
    assert np.isnan(golden['disp_map']).any(), 'golden lacks empty rays (NaN disparity, RN:381)'


def test_coarse_depths(golden):
    rays = T(golden['rays'])
    t = torch.linspace(0., 1., 64)
    z = rays[:, 6:7] * (1. - t) + rays[:, 7:8] * t
    assert np.array_equal(z.numpy(), golden['z0'])


def test_mlp_and_encoding(golden, wfit):
    sdc, sdf = wfit
    rays = T(golden['rays'])
    for z, sd, key in ((golden['z0'], sdc, 'raw0'), (golden['z1'], sdf, 'raw1')):
        z = T(z)
        pts = rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]
        raw = O.run_network(pts, rays[:, 8:11], sd)
        close(raw.numpy(), golden[key], 1e-5, key)


def test_raw2outputs(golden):
    rays = T(golden['rays'])
    for raw, z, names in ((golden['raw0'], golden['z0'], ('rgb0', 'disp0', 'acc0', 'weights0', 'depth0')),
                          (golden['raw1'], golden['z1'], ('rgb_map', 'disp_map', 'acc_map', 'weights1', 'depth_map'))):
        outs = O.raw2outputs(T(raw), T(z), rays[:, 3:6])
        for o, nme in zip(outs, names):
            close(o.numpy(), golden[nme], 1e-6, nme)
    wb = O.raw2outputs(T(golden['raw1']), T(golden['z1']), rays[:, 3:6], white_bkgd=True)[0]
    close(wb.numpy(), golden['wb_rgb_map'], 1e-6, 'white_bkgd rgb')


def test_sample_pdf_and_merge(golden):
    z0, w0 = T(golden['z0']), T(golden['weights0'])
    zs = O.sample_pdf(.5 * (z0[:, 1:] + z0[:, :-1]), w0[:, 1:-1], 128, det=True)
    close(zs.numpy(), golden['z_samples'], 1e-6, 'z_samples')
    z1, _ = torch.sort(torch.cat([z0, zs], -1), -1)
    close(z1.numpy(), golden['z1'], 1e-6, 'z1')
    close(torch.std(zs, -1, unbiased=False).numpy(), golden['z_std'], 1e-6, 'z_std')
    # stratified variant with explicit randoms
    pz0, pw0 = T(golden['p_z0']), T(golden['p_weights0'])
    pzs = O.sample_pdf(.5 * (pz0[:, 1:] + pz0[:, :-1]), pw0[:, 1:-1], 128, det=False, u=T(golden['p_u']))
    close(pzs.numpy(), golden['p_z_samples'], 1e-6, 'p_z_samples')


def test_render_rays_end_to_end(golden, wfit):
    sdc, sdf = wfit
    with torch.no_grad():
        r = O.render_rays(T(golden['rays']), sdc, sdf, 64, 128, retraw=True)
    for k in ('rgb_map', 'disp_map', 'acc_map', 'rgb0', 'disp0', 'acc0', 'z_std', 'raw'):
        close(r[k].numpy(), golden['e2e_' + k], 2e-5, 'e2e ' + k)
    # perturbed depths
    with torch.no_grad():
        rp = O.render_rays(T(golden['rays']), sdc, sdf, 64, 128, perturb=1., t_rand=T(golden['p_t_rand']),
                           u=T(golden['p_u']), return_internals=True)
    close(rp['_internals']['z0'].numpy(), golden['p_z0'], 1e-6, 'p_z0')
    close(rp['_internals']['z1'].numpy(), golden['p_z1'], 2e-5, 'p_z1')


def test_get_rays(golden):
    o, d = O.get_rays(10, 12, golden['getrays_K'], T(golden['getrays_c2w']))
    assert np.array_equal(o.numpy(), golden['getrays_o'])
    close(d.numpy(), golden['getrays_d'], 1e-7, 'rays_d')


@pytest.mark.skipif(not ref_import.available(), reason='reference tree only exists in the build container')
def test_oracle_vs_live_reference(wfit):
    """Fresh rays / pose, both weight sets, straight against RN.render on CPU."""
    RN, RH = ref_import.load()
    torch.autograd.set_detect_anomaly(False)
    sdc, sdf = wfit
    H = W = 400
    c2w = O.pose_spherical(88., 200.0 - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    sel = torch.arange(3, H * W, 1237)[:128]
    rays = torch.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0)
    for a, b in ((sdc, sdf), (O.random_state_dict(5, scale=2.0), O.random_state_dict(6, scale=2.0))):
        kw = ref_import.render_kwargs(a, b, O.YCBV_NEAR, O.YCBV_FAR)
        with torch.no_grad():
            ref = RN.render(H, W, torch.tensor(O.YCBV_K_400), chunk=512, rays=rays, retraw=True, **kw)
            mine = O.render(H, W, O.YCBV_K_400, a, b, chunk=512, rays=rays, near=O.YCBV_NEAR, far=O.YCBV_FAR, retraw=True)
        for x, y, nme in zip(ref[:3], mine[:3], ('rgb', 'disp', 'acc')):
            close(y.numpy(), x.numpy(), 1e-6, nme)
        for k in ref[3]:
            close(mine[3][k].numpy(), ref[3][k].numpy(), 1e-6, k)


def test_oracle_viewless_against_golden(viewless_golden):
    """use_viewdirs=False: the oracle against outputs of the reference's own NeRF(use_viewdirs=False) through RN.render / RN.run_network
    (oracle/make_golden_viewless.py); runs everywhere, the live comparison below only in the build container."""
    g, (sdc, sdf) = viewless_golden
    ro, rd = T(g['ro']), T(g['rd'])
    with torch.no_grad():
        raw_pts = O.run_network(T(g['pts']), None, sdf)
        r = O.render(400, 400, O.YCBV_K_400, sdc, sdf, chunk=512, rays=(ro, rd), near=float(g['near']), far=float(g['far']),
                     retraw=True, use_viewdirs=False)
    close(raw_pts.numpy(), g['raw_pts'], 1e-5, 'run_network(viewdirs=None)')
    for x, nme in zip(r[:3], ('rgb_map', 'disp_map', 'acc_map')):
        close(x.numpy(), g[nme], 2e-5, nme)
    for k in ('rgb0', 'disp0', 'acc0', 'z_std'):
        close(r[3][k].numpy(), g[k], 2e-5, k)
    close(r[3]['raw'].numpy()[::8], g['raw'], 2e-5, 'raw')


@pytest.mark.skipif(not ref_import.available(), reason='reference tree only exists in the build container')
def test_oracle_viewless_vs_live_reference(wfit):
    """use_viewdirs=False (RH:95-96, RH:119-120, RN:32, RN:111): the oracle's 8-column ray batches and `output_linear` head
    against RN.render with the reference's own NeRF(use_viewdirs=False, input_ch_views=0, output_ch=5)."""
    RN, RH = ref_import.load()
    torch.autograd.set_detect_anomaly(False)
    a, b = (O.viewless_state_dict(sd) for sd in wfit)
    H = W = 400
    c2w = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    sel = torch.arange(80000 - 2000, 80000 + 2000, 31)
    rays = torch.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0)
    kw = ref_import.render_kwargs(a, b, O.YCBV_NEAR, O.YCBV_FAR)
    assert kw['use_viewdirs'] is False
    with torch.no_grad():
        ref = RN.render(H, W, torch.tensor(O.YCBV_K_400), chunk=512, rays=rays, retraw=True, **kw)
        mine = O.render(H, W, O.YCBV_K_400, a, b, chunk=512, rays=rays, near=O.YCBV_NEAR, far=O.YCBV_FAR, retraw=True, use_viewdirs=False)
    assert float(ref[2].max()) > 0.9 and float(ref[2].min()) < 0.1, 'the sample should hold hits and misses'
    assert ref[3]['raw'].shape[-1] == 5
    for x, y, nme in zip(ref[:3], mine[:3], ('rgb', 'disp', 'acc')):
        close(y.numpy(), x.numpy(), 1e-6, nme)
    for k in ref[3]:
        close(mine[3][k].numpy(), ref[3][k].numpy(), 1e-6, k)


@pytest.mark.skipif(not ref_import.available(), reason='reference tree only exists in the build container')
def test_oracle_autograd_vs_live_reference_tape(wfit):
    """The backward tests compare against autograd through the oracle; pin that tape to the reference's own
    (RN:168-181: autograd.grad(rgb_p, batch_rays, grad_outputs=patch_grad_E)) and to loss.backward() (RN:691-707)."""
    RN, RH = ref_import.load()
    torch.autograd.set_detect_anomaly(False)
    sdc, sdf = wfit
    H = W = 400
    c2w = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    sel = torch.arange(80000 - 600, 80000 + 600, 37)
    g = torch.randn(len(sel), 3, generator=torch.Generator().manual_seed(0))
    # reference tape
    coarse, fine, query = ref_import.build_models(sdc, sdf)
    kw = ref_import.render_kwargs(sdc, sdf, O.YCBV_NEAR, O.YCBV_FAR)
    kw.update(network_fn=coarse, network_fine=fine, network_query_fn=query)
    br = torch.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0).clone().requires_grad_(True)
    rgb_ref = RN.render(H, W, torch.tensor(O.YCBV_K_400), chunk=512, rays=br, retraw=True, **kw)[0]
    (g_ref,) = torch.autograd.grad(rgb_ref, br, grad_outputs=g, retain_graph=True)
    loss = ((rgb_ref - 0.5) ** 2).mean()
    loss.backward()
    # oracle tape
    sf = {k: v.clone().requires_grad_(True) for k, v in sdf.items()}
    bo = br.detach().clone().requires_grad_(True)
    rgb = O.render(H, W, O.YCBV_K_400, sdc, sf, chunk=512, rays=bo, near=O.YCBV_NEAR, far=O.YCBV_FAR)[0]
    (g_or,) = torch.autograd.grad(rgb, bo, grad_outputs=g, retain_graph=True)
    ((rgb - 0.5) ** 2).mean().backward()
    assert (g_or - g_ref).abs().max() <= 1e-5 * g_ref.abs().max()
    for name, p in fine.named_parameters():
        assert (sf[name].grad - p.grad).abs().max() <= 1e-5 * max(1e-12, p.grad.abs().max()), name


@pytest.mark.skipif(not ref_import.available(), reason='reference tree only exists in the build container')
def test_oracle_image_stage_vs_live_reference():
    """to8b (RH:14) and the pose pull-back of RN:179-181: autograd.grad(batch_rays, <pose>, grad_outputs=dLdray) with the
    reference's own get_rays (RH:156-165) and the normalisation inside its render() (RN:97) against the oracle's."""
    RN, RH = ref_import.load()
    torch.autograd.set_detect_anomaly(False)
    x = np.random.RandomState(0).uniform(-0.3, 1.3, size=(50, 3)).astype(np.float32)
    assert np.array_equal(RH.to8b(x), O.to8b(x))
    H, W = 12, 10
    K = [[60.0, 0, 4.5], [0, 61.0, 6.5], [0, 0, 1]]
    c2w = O.pose_spherical(80., 40., 1.05)[:3, :4].clone().requires_grad_(True)
    ro, rd = RH.get_rays(H, W, torch.tensor(K), c2w)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)                      # RN:97
    g = torch.randn(H * W, 11, generator=torch.Generator().manual_seed(3))
    loss = (ro * g[:, 0:3]).sum() + (rd * g[:, 3:6]).sum() + (vd * g[:, 8:11]).sum()
    ref, = torch.autograd.grad(loss, c2w)
    mine = O.rays_grad_to_c2w(H, W, K, c2w.detach(), g)
    assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-6)
