"""use_viewdirs=False (RH:95-96, RH:119-120, RN:32, RN:91, RN:111, RN:263, RN:437) on the GPU path against the CPU oracle, whose
use_viewdirs=False branch is pinned to the reference's own NeRF(use_viewdirs=False) in tests/test_oracle_pinned.py.

The sm_100a kernels are built for the view-dependent head; an `output_linear` network is packed into that operand layout
(run_nerf._viewless_operands: feature = I, views = [A; -A], rgb = [I3, -I3], so relu(x) - relu(-x) = x), and 8-column ray batches get
three zero columns that meet zero weights.  Tolerances as in test_gpu_parity.py: 1e-3 * max(1, |ref|) on raw outputs and maps.
"""
import numpy as np
import pytest
import torch

import nerf_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope='module')
def nsr():
    import neural_sim_nerf_b200 as m
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return m


def viewless_module(nsr, sd):
    net = nsr.NeRF(input_ch_views=0, output_ch=sd['output_linear.weight'].shape[0], use_viewdirs=False)
    net.load_state_dict(sd)
    return net.cuda().requires_grad_(False)


@pytest.fixture(scope='module')
def sds(wfit):
    return tuple(O.viewless_state_dict(sd, output_ch=5) for sd in wfit)


@pytest.fixture(scope='module')
def nets(nsr, sds):
    return tuple(viewless_module(nsr, sd) for sd in sds)


def camera_rays(n_side, phi):
    H = W = 400
    c2w = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    ii = torch.linspace(0, 399, n_side).long()
    sel = (ii[:, None] * W + ii[None, :]).reshape(-1)
    return ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]


def close(got, ref, what, tol=TOL):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert bool((torch.isnan(got) == torch.isnan(ref)).all()), f'{what}: NaN pattern'
    m = ~torch.isnan(ref)
    err = ((got[m] - ref[m]).abs() / ref[m].abs().clamp(min=1.0)).max().item() if m.any() else 0.0
    assert err <= tol, f'{what}: max err {err:.3e} > {tol}'


def test_against_reference_golden(nsr, viewless_golden):
    """The CUDA path against outputs of the unmodified reference run with use_viewdirs=False (tests/golden/viewless_golden.npz)."""
    g, gsds = viewless_golden
    gnets = tuple(viewless_module(nsr, sd) for sd in gsds)
    ro, rd = torch.from_numpy(g['ro']), torch.from_numpy(g['rd'])
    kw = dict(network_fn=gnets[0], network_fine=gnets[1], network_query_fn=None, N_samples=64, N_importance=128, perturb=False,
              raw_noise_std=0., white_bkgd=False, lindisp=False, ndc=False, near=float(g['near']), far=float(g['far']), use_viewdirs=False)
    with torch.no_grad():
        got = nsr.render(400, 400, O.YCBV_K_400, chunk=512, rays=torch.stack([ro, rd], 0).cuda(), retraw=True, **kw)
        raw_pts = nsr.run_network(torch.from_numpy(g['pts']).cuda(), None, gnets[1])
    for i, nme in enumerate(('rgb_map', 'disp_map', 'acc_map')):
        close(got[i], torch.from_numpy(g[nme]), nme)
    for k in ('rgb0', 'acc0'):
        close(got[3][k], torch.from_numpy(g[k]), k)
    close(raw_pts, torch.from_numpy(g['raw_pts'][..., :4]), 'run_network(viewdirs=None)')
    # raw of the last pass: the reference carries output_linear's fifth column (RN:267), the kernel the four the compositor reads
    assert got[3]['raw'].shape == (400, 192, 4)
    empty = torch.from_numpy(g['acc_map'][::8]) < 1e-3      # depths agree there (uniform pdf); elsewhere they may sit in a different zero-weight bin
    close(got[3]['raw'][::8][empty], torch.from_numpy(g['raw'][..., :4])[empty], 'raw on empty rays')


def test_run_network_without_viewdirs(nsr, sds, nets):
    """RN:26-40 with viewdirs=None: raw [n,S,4] = the four columns of output_linear the compositor reads."""
    ro, rd = camera_rays(12, 22.5)
    z = torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, 64)
    pts = ro[:, None, :] + rd[:, None, :] * z[None, :, None]
    with torch.no_grad():
        ref = O.run_network(pts, None, sds[1])
        got = nsr.run_network(pts.cuda(), None, nets[1])
    assert ref.shape[-1] == 5 and got.shape == (*pts.shape[:-1], 4)
    assert float(ref[..., 3].max()) > 10.0, 'the sample should cross the object'
    close(got, ref[..., :4], 'raw')
    with pytest.raises(ValueError):
        nsr.run_network(pts.cuda(), None, nsr.NeRF().cuda())          # a view-dependent network needs its viewdirs


def test_nerf_forward_four_columns(nsr, wfit):
    """NeRF.forward(x) (RH:99-122) of an output_ch=4 use_viewdirs=False module on embedded points; output_ch=5 is refused (its fifth
    column is not computed)."""
    sd4 = O.viewless_state_dict(wfit[1], output_ch=4)
    net = viewless_module(nsr, sd4)
    ro, rd = camera_rays(9, 112.5)
    pts = (ro[:, None, :] + rd[:, None, :] * torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, 32)[None, :, None]).reshape(-1, 3)
    x = O.embed(pts, 10)
    with torch.no_grad():
        ref = O.mlp_forward(x, sd4)
        got = net(x.cuda())
        close(got, ref, 'NeRF.forward')
        with pytest.raises(NotImplementedError):
            viewless_module(nsr, O.viewless_state_dict(wfit[1], output_ch=5))(x.cuda())


@pytest.mark.parametrize('phi,n_side', [(22.5, 24), (200.0, 16)])
def test_render_without_viewdirs_matches_oracle(nsr, sds, nets, phi, n_side):
    """render(..., use_viewdirs=False) (RN:58-123): caller-made rays, 8-column batches inside, coarse + fine pass."""
    ro, rd = camera_rays(n_side, phi)
    kw = dict(network_fn=nets[0], network_fine=nets[1], network_query_fn=None, N_samples=64, N_importance=128, perturb=False,
              raw_noise_std=0., white_bkgd=False, lindisp=False, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, use_viewdirs=False)
    with torch.no_grad():
        ref = O.render(400, 400, O.YCBV_K_400, sds[0], sds[1], chunk=512, rays=(ro, rd), near=O.YCBV_NEAR, far=O.YCBV_FAR, use_viewdirs=False)
        got = nsr.render(400, 400, O.YCBV_K_400, chunk=4096, rays=torch.stack([ro, rd], 0).cuda(), **kw)
    assert float(ref[2].max()) > 0.9 and float(ref[2].min()) < 0.1
    for i, nme in enumerate(('rgb_map', 'disp_map', 'acc_map')):
        close(got[i], ref[i], f'{nme} phi={phi}')
    for k in ('rgb0', 'acc0'):
        close(got[3][k], ref[3][k], f'{k} phi={phi}')
    # the 8-column batch of RN:111 straight into render_rays, and through the whole-image call (c2w route)
    batch8 = O.pack_rays(ro, rd, O.YCBV_NEAR, O.YCBV_FAR, use_viewdirs=False)
    assert batch8.shape[1] == 8
    with torch.no_grad():
        direct = nsr.render_rays(batch8.cuda(), nets[0], None, 64, N_importance=128, network_fine=nets[1])
    close(direct['rgb_map'], ref[0], 'render_rays on 8 columns')
    with pytest.raises(ValueError):
        nsr.render_rays(batch8.cuda(), nsr.NeRF().cuda(), None, 64)
    with pytest.raises(ValueError):
        nsr.render(400, 400, O.YCBV_K_400, rays=torch.stack([ro, rd], 0).cuda(), **dict(kw, network_fn=nsr.NeRF().cuda(), N_importance=0))


def test_whole_image_route_and_pose_gradient(nsr, sds, nets):
    """render(c2w=...) / render_image and the pose path (RN:168-181): dL/d(ray_batch) of an `output_linear` network against autograd through
    the oracle; the three view columns carry no gradient."""
    H = W = 40
    s = H / 400.0
    K = [[O.YCBV_K_400[0][0] * s, 0.0, O.YCBV_K_400[0][2] * s], [0.0, O.YCBV_K_400[1][1] * s, O.YCBV_K_400[1][2] * s], [0.0, 0.0, 1.0]]
    c2w = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    kw = dict(network_fn=nets[0], network_fine=nets[1], network_query_fn=None, N_samples=64, N_importance=128, perturb=False,
              raw_noise_std=0., white_bkgd=False, lindisp=False, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, use_viewdirs=False)
    with torch.no_grad():
        ref = O.render(H, W, K, sds[0], sds[1], chunk=512, c2w=c2w, near=O.YCBV_NEAR, far=O.YCBV_FAR, use_viewdirs=False)
        got = nsr.render(H, W, K, chunk=1 << 20, c2w=c2w.cuda(), **kw)
        img = nsr.render_image(H, W, K, c2w, want=('rgb_map',), **kw)
    close(got[0], ref[0], 'rgb_map via c2w')
    close(img['rgb_map'], ref[0], 'rgb_map via render_image')
    # data gradient
    ro, rd = O.get_rays(H, W, K, c2w)
    batch = O.pack_rays(ro, rd, O.YCBV_NEAR, O.YCBV_FAR, use_viewdirs=False)
    g = torch.randn(H * W, 3, generator=torch.Generator().manual_seed(2))
    r_cpu = batch.clone().requires_grad_(True)
    out = O.render_rays(r_cpu, sds[0], sds[1], 64, 128)
    (ref_grad,) = torch.autograd.grad(out['rgb_map'], r_cpu, grad_outputs=g)
    r_gpu = batch.cuda().requires_grad_(True)
    mine = nsr.render_rays(r_gpu, nets[0], None, 64, N_importance=128, network_fine=nets[1])
    (got_grad,) = torch.autograd.grad(mine['rgb_map'], r_gpu, grad_outputs=g.cuda())
    assert got_grad.shape == (H * W, 8)
    # origins and directions (what RN:163 stacks into batch_rays); near / far are constants made inside render() (RN:109-110) and get
    # no gradient here (the oracle's packed batch carries d rgb / d near through the coarse depths, which the reference never asks for)
    scale = ref_grad[:, 0:6].abs().max().item()
    err = (got_grad.cpu()[:, 0:6] - ref_grad[:, 0:6]).abs().max().item()
    assert scale > 0 and err <= TOL * scale, f'dL/drays: {err:.3e} > {TOL} * {scale:.3e}'
    assert float(got_grad[:, 6:8].abs().max()) == 0.0
    # closed-form pose pull-back of the whole image
    gi = torch.randn(H * W, 3, generator=torch.Generator().manual_seed(3))
    _, d_c2w = nsr.render_image_grad(H, W, K, c2w.cuda(), gi.cuda(), **kw)
    c = c2w.clone().requires_grad_(True)
    ro2, rd2 = O.get_rays(H, W, K, c)
    rgb = O.render_rays(O.pack_rays(ro2, rd2, O.YCBV_NEAR, O.YCBV_FAR, use_viewdirs=False), sds[0], sds[1], 64, 128)['rgb_map']
    (ref_c,) = torch.autograd.grad(rgb, c, grad_outputs=gi)
    sc = ref_c.abs().max().item()
    assert (d_c2w.cpu() - ref_c).abs().max().item() <= TOL * sc, (d_c2w.cpu(), ref_c)


def test_trainable_viewless_network_is_refused(nsr, sds):
    """Parameter gradients of `output_linear` networks are not built: asking for a graph fails loudly, no silent constant."""
    net = viewless_module(nsr, sds[0]).requires_grad_(True)
    ro, rd = camera_rays(4, 22.5)
    batch8 = O.pack_rays(ro, rd, O.YCBV_NEAR, O.YCBV_FAR, use_viewdirs=False).cuda()
    with pytest.raises(NotImplementedError):
        nsr.render_rays(batch8, net, None, 64)
    with torch.no_grad():
        nsr.render_rays(batch8, net, None, 64)
