"""GPU tests of the image-level entries: nsr_to8b, nsr_make_rays_dev, nsr_rays_grad_to_c2w,
nsr_render_image_forward and the mirrors built on them (render_image, render_image_grad, the pipelined render_path)."""
import numpy as np
import pytest
import torch

import nerf_oracle as O

pytestmark = pytest.mark.gpu

K20 = [[66.0, 0, 9.5], [0, 66.0, 10.5], [0, 0, 1]]


@pytest.fixture(scope='module')
def nsr():
    import neural_sim_nerf_b200 as m
    assert torch.cuda.is_available()
    return m


@pytest.fixture(scope='module')
def nets(nsr, wfit):
    out = []
    for sd in wfit:
        m = nsr.NeRF()
        m.load_state_dict(sd)
        out.append(m.cuda())
    return out


def kwargs(nets, **over):
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1],
              use_viewdirs=True, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False)
    kw.update(over)
    return kw


@pytest.mark.parametrize('n', [0, 1, 3, 4, 5, 1023, 4096 * 3 + 2])
def test_to8b_is_bit_exact(nsr, n):
    rs = np.random.RandomState(n)
    x = rs.uniform(-0.2, 1.2, size=n).astype(np.float32)
    if n >= 5:
        x[:5] = [0.0, 1.0, 1 / 255, 0.5, np.nextafter(np.float32(1.0), np.float32(0.0))]
    if n > 1000:
        k = np.arange(256, dtype=np.float32) / np.float32(255)      # the exact bucket edges and their neighbours
        x[100:356] = k
        x[400:656] = np.nextafter(k, np.float32(-1))
        x[700:956] = np.nextafter(k, np.float32(2))
    got = nsr.to8b_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert got.dtype == np.uint8 and np.array_equal(got, O.to8b(x))


def test_to8b_unaligned_views_and_shapes(nsr):
    x = torch.rand(33, 7, 3, device='cuda') * 1.4 - 0.2
    base = torch.empty(x.numel() + 1, device='cuda')
    off = base[1:].view(33, 7, 3)            # 4-byte aligned only
    off.copy_(x)
    a, b = nsr.to8b_device(x), nsr.to8b_device(off)
    assert a.shape == x.shape and torch.equal(a, b)
    assert np.array_equal(a.cpu().numpy(), O.to8b(x.cpu().numpy()))


def test_make_rays_dev_equals_host_variant(nsr):
    import ctypes
    H, W = 37, 53
    L = nsr.lib()
    pose44 = O.pose_spherical(85., 31. - 180., 1.03)
    host = nsr.make_rays(H, W, O.YCBV_K_400, pose44[:3, :4], 0.3, 1.9)
    Kh = np.ascontiguousarray(np.asarray(O.YCBV_K_400, dtype=np.float32))
    for c, ld in ((pose44[:3, :4].contiguous().cuda(), 4), (torch.cat([pose44, pose44], 1).cuda(), 8)):
        out = torch.empty(H * W, 11, device='cuda')
        rc = L.nsr_make_rays_dev(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(c.data_ptr()), ld, 0.3, 1.9,
                                 ctypes.c_void_p(out.data_ptr()), None)
        assert rc == 0, L.nsr_last_error()
        assert torch.equal(out, host)
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, pose44[:3, :4])
    assert torch.allclose(host.cpu(), O.pack_rays(ro, rd, 0.3, 1.9), rtol=0, atol=1e-6)
    # the mirror picks the device variant for poses that live on the GPU (a [3,4] view of a [4,4] matrix, a double-precision pose)
    assert torch.equal(nsr.make_rays(H, W, O.YCBV_K_400, pose44.cuda()[:3, :4], 0.3, 1.9), host)
    assert torch.equal(nsr.make_rays(H, W, O.YCBV_K_400, pose44.double().cuda(), 0.3, 1.9), host)


@pytest.mark.parametrize('H,W', [(20, 20), (33, 47), (400, 400)])
def test_rays_grad_to_c2w_vs_autograd(nsr, H, W):
    K = O.YCBV_K_400 if H == 400 else [[66.0, 0, W / 2 - 0.3], [0, 64.0, H / 2 + 0.2], [0, 0, 1]]
    c2w = O.pose_spherical(77., 140. - 180., 1.07)[:3, :4]
    g = torch.randn(H * W, 11, generator=torch.Generator().manual_seed(H))
    ref = O.rays_grad_to_c2w(H, W, K, c2w, g)
    rays = nsr.make_rays(H, W, K, c2w, 0.3, 1.9)
    got = nsr.rays_grad_to_c2w(H, W, K, rays, g.cuda())
    again = nsr.rays_grad_to_c2w(H, W, K, rays, g.cuda())
    assert torch.equal(got, again), 'the reduction must be deterministic'
    scale = ref.abs().max().item()
    # fp32 autograd sums H*W terms too: agree to a few 1e-5 of the largest entry
    assert (got.cpu() - ref).abs().max().item() <= 1e-4 * scale, (got, ref)


def test_rays_grad_to_c2w_pixel_subset(nsr):
    H, W = 40, 30
    K = [[66.0, 0, 14.5], [0, 66.0, 20.5], [0, 0, 1]]
    c2w = O.pose_spherical(90., 10., 1.01)[:3, :4]
    sel = torch.randperm(H * W, generator=torch.Generator().manual_seed(1))[:257]
    g_sub = torch.randn(257, 11, generator=torch.Generator().manual_seed(2))
    g_full = torch.zeros(H * W, 11)
    g_full[sel] = g_sub
    ref = O.rays_grad_to_c2w(H, W, K, c2w, g_full)
    rays = nsr.make_rays(H, W, K, c2w, 0.3, 1.9)
    got = nsr.rays_grad_to_c2w(H, W, K, rays[sel.cuda()], g_sub.cuda(), pixel_idx=sel)
    assert (got.cpu() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    with pytest.raises(nsr.NsrError):
        nsr.rays_grad_to_c2w(H, W, K, rays[:100], g_sub[:100].cuda())        # a subset without pixel indices


def test_render_image_equals_render(nsr, nets):
    """One C call per image == render(c2w=...) + to8b, bit for bit; host and device poses agree."""
    H = W = 20
    pose = O.pose_spherical(90., 22.5 - 180., 1.01)
    kw = kwargs(nets)
    with torch.no_grad():
        rgb, disp, acc, extras = nsr.render(H, W, K20, chunk=1 << 16, c2w=pose[:3, :4], **kw)
        names = ('rgb8', 'rgb_map', 'disp_map', 'acc_map', 'rgb0', 'disp0', 'acc0', 'z_std')
        a = nsr.render_image(H, W, K20, pose[:3, :4], want=names, **kw)
        b = nsr.render_image(H, W, K20, pose.cuda()[:3, :4], want=names, **kw)       # strided view of a [4,4] device pose
        only8 = nsr.render_image(H, W, K20, pose[:3, :4], want=('rgb8',), **kw)
    ref = {'rgb_map': rgb, 'disp_map': disp, 'acc_map': acc, **{k: extras[k] for k in ('rgb0', 'disp0', 'acc0', 'z_std')}}
    for out in (a, b):
        for k, v in ref.items():
            assert out[k].shape == v.shape, k
            assert torch.equal(torch.nan_to_num(out[k], nan=-7.0), torch.nan_to_num(v, nan=-7.0)), k
        assert np.array_equal(out['rgb8'].cpu().numpy(), O.to8b(rgb.cpu().numpy()))
    assert torch.equal(only8['rgb8'], a['rgb8'])
    assert a['rays'].shape == (H * W, 11)
    assert float(rgb.max()) > 0.05, 'the test image must show the object'


def test_render_image_rejects_what_it_does_not_cover(nsr, nets):
    pose = O.pose_spherical(90., 22.5 - 180., 1.01)
    with pytest.raises(NotImplementedError):
        nsr.render_image(20, 20, K20, pose[:3, :4], **kwargs(nets, perturb=1.0))
    with pytest.raises(ValueError):
        nsr.render_image(20, 20, K20, pose[:3, :4], want=('rgb0',), **kwargs(nets, N_importance=0))


def test_render_image_grad_vs_oracle_autograd(nsr, wfit, nets):
    """Forward + backward of one image without an autograd tape: rgb and dL/dc2w against autograd through the oracle
    renderer and get_rays on the CPU (the chain RN:148-181 differentiates)."""
    H = W = 20
    pose = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    g = torch.randn(H * W, 3, generator=torch.Generator().manual_seed(8)) * 1e-2
    c = pose.clone().requires_grad_(True)
    ro, rd = O.get_rays(H, W, K20, c)
    out = O.render(H, W, K20, wfit[0], wfit[1], chunk=512, rays=(ro.reshape(-1, 3), rd.reshape(-1, 3)),
                   near=O.YCBV_NEAR, far=O.YCBV_FAR, N_samples=64, N_importance=128)
    ref, = torch.autograd.grad(out[0], c, grad_outputs=g)
    rgb, d_c2w = nsr.render_image_grad(H, W, K20, pose.cuda(), g.cuda(), **kwargs(nets))
    assert rgb.shape == (H, W, 3) and d_c2w.shape == (3, 4)
    assert (rgb.cpu().reshape(-1, 3) - out[0].detach()).abs().max().item() <= 1e-3
    scale = ref.abs().max().item()
    assert scale > 0
    assert (d_c2w.cpu() - ref).abs().max().item() <= 1e-3 * scale, (d_c2w, ref)


def test_render_path_pipelined_many_images(nsr, nets, tmp_path):
    """More images than staging slots: every PNG / array lands in its own index."""
    H = W = 20
    hwf = [H, W, 66.0]
    phis = [22.5 + 45.0 * k for k in range(5)]
    poses = torch.stack([O.pose_spherical(90., p - 180., 1.01) for p in phis])
    kw = kwargs(nets)
    imgs, disps = nsr.render_path(None, poses.cuda(), hwf, K20, 1 << 16, kw, savedir=str(tmp_path), object_id=7)
    imgs_host, _ = nsr.render_path(None, poses, hwf, K20, 1 << 16, kw)
    assert imgs.shape == (5, H, W, 3) and disps.shape == (5, H, W)
    assert np.array_equal(imgs, imgs_host)
    from PIL import Image
    for i in range(5):
        with torch.no_grad():
            ref = nsr.render(H, W, K20, chunk=1 << 16, c2w=poses[i, :3, :4], **kw)[0].cpu().numpy()
        assert np.array_equal(imgs[i], ref), i
        png = np.asarray(Image.open(tmp_path / '7' / f'{i:03d}.png'))
        assert np.array_equal(png, O.to8b(ref)), i
    assert not np.array_equal(imgs[0], imgs[1])


def test_render_image_grad_routes_agree(nsr, nets):
    """nsr_render_image_grad (one C call, saved sign bits) against the staged recompute route it falls back to when the bits do
    not fit in memory; non-square image, white background."""
    H, W = 24, 18
    K = [[70.0, 0, 8.5], [0, 70.0, 12.5], [0, 0, 1]]
    pose = O.pose_spherical(88., 112.5 - 180., 1.01)[:3, :4].cuda()
    g = torch.randn(H * W, 3, generator=torch.Generator().manual_seed(12)).cuda() * 1e-2
    kw = kwargs(nets, white_bkgd=True)
    out = {}
    for save in (True, False):
        nsr.run_nerf.SAVE_RELU_MASK = save
        try:
            before = nsr.lib().nsr_launch_count()
            out[save] = nsr.render_image_grad(H, W, K, pose, g, **kw)
            torch.cuda.synchronize()
        finally:
            nsr.run_nerf.SAVE_RELU_MASK = True
    assert torch.equal(out[True][0], out[False][0])
    scale = float(out[False][1].abs().max())
    assert scale > 0 and float((out[True][1] - out[False][1]).abs().max()) <= 1e-5 * scale
    with pytest.raises(ValueError):
        nsr.render_image_grad(H, W, K, pose, g[:-1], **kw)


def test_row_bands_of_an_image_add_up(nsr, nets):
    """render_image_grad(rows=...): what dist.plan_images hands to each rank of a group -- the bands' pixels are the image's rows,
    their dL/dc2w contributions add up to the whole image's gradient."""
    H = W = 48
    K = [[160.0, 0, 23.4], [0, 160.1, 24.1], [0, 0, 1]]
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
              near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
    pose = O.pose_spherical(90., 112.5 - 180., 1.01)[:3, :4].cuda()
    g = torch.randn(H * W, 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(11))
    rgb_all, d_all = nsr.run_nerf.render_image_grad(H, W, K, pose, g, **kw)
    total = torch.zeros_like(d_all)
    for r0, r1 in ((0, 13), (13, 14), (14, 48)):
        rgb_b, d_b = nsr.run_nerf.render_image_grad(H, W, K, pose, g, rows=(r0, r1), **kw)
        assert torch.equal(rgb_b, rgb_all[r0:r1])
        total += d_b
    assert float(d_all.abs().max()) > 0
    assert float((total - d_all).abs().max()) <= 1e-5 * float(d_all.abs().max())
    with pytest.raises(ValueError):
        nsr.run_nerf.render_image_grad(H, W, K, pose, g, rows=(10, 49), **kw)


def test_pack_rays_kernel_matches_the_reference_ops(nsr):
    """nsr_pack_rays = RN:97 + RN:106-112 (viewdirs = d / |d|, cat[o, d, near, far, viewdirs]) for caller-made rays; render(rays=...) uses it."""
    import ctypes
    g = torch.Generator(device='cuda').manual_seed(7)
    o = torch.randn(1003, 3, device='cuda', generator=g)
    d = torch.randn(1003, 3, device='cuda', generator=g) * 3.0
    out = torch.empty(1003, 11, device='cuda')
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    assert nsr.lib().nsr_pack_rays(P(o), P(d), 1003, 0.25, 2.5, P(out), None) == 0
    torch.cuda.synchronize()
    ref = torch.cat([o, d, torch.full((1003, 1), 0.25, device='cuda'), torch.full((1003, 1), 2.5, device='cuda'), d / torch.norm(d, dim=-1, keepdim=True)], -1)
    assert torch.equal(out[:, :8], ref[:, :8])
    assert float((out[:, 8:] - ref[:, 8:]).abs().max()) <= 2e-7
