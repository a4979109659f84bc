"""GPU parity over sample-count shapes and adversarial inputs for the exact-fp32 ray-stage kernels (and their backward),
beyond the 64+128 configuration of the BASELINE configs."""
import ctypes

import numpy as np
import pytest
import torch

import nerf_oracle as O

pytestmark = pytest.mark.gpu


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


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    m = ~torch.isnan(a)
    return ((a - b).abs()[m] / b.abs()[m].clamp(min=1.0)).max().item() if m.any() else 0.0


@pytest.mark.parametrize('S', [2, 3, 31, 32, 33, 64, 100, 192, 256])
def test_raw2outputs_shapes_and_extremes(nsr, S):
    g = torch.Generator().manual_seed(S)
    n = 257
    raw = torch.randn(n, S, 4, generator=g) * 3
    raw[:, :, 3] = torch.randn(n, S, generator=g) * 40              # densities from -120 to 120
    raw[0] = -5.0                                                   # an empty ray -> acc 0, NaN disparity
    raw[1, :, 3] = 1e4                                              # opaque at the first sample
    z = torch.sort(torch.rand(n, S, generator=g) * 1.6 + 0.3, -1).values
    if S > 2:
        z[2, 1] = z[2, 0]                                           # a zero-length interval
    d = torch.randn(n, 3, generator=g)
    for white in (False, True):
        ref = O.raw2outputs(raw, z, d, white_bkgd=white)
        got = nsr.raw2outputs(raw.cuda(), z.cuda(), d.cuda(), white_bkgd=white)
        for a, b, name in zip(got, ref, ('rgb', 'disp', 'acc', 'weights', 'depth')):
            assert rel(a, b) <= 2e-5, (name, S, white)


@pytest.mark.parametrize('B,N', [(2, 1), (3, 7), (32, 64), (63, 128), (64, 33), (129, 200)])
def test_sample_pdf_shapes(nsr, B, N):
    g = torch.Generator().manual_seed(B * 1000 + N)
    n = 130
    bins = torch.sort(torch.rand(n, B, generator=g), -1).values
    w = torch.rand(n, B - 1, generator=g) ** 4
    w[0] = 0.0                                                      # all-empty -> uniform
    w[1, : (B - 1) // 2] = 0.0                                      # half empty
    u = torch.rand(n, N, generator=g)
    ref = O.sample_pdf(bins, w, N, det=False, u=u)
    L = nsr.lib()
    out = torch.empty(n, N, device='cuda')
    b, ww, uu = bins.cuda(), w.cuda(), u.cuda()
    assert L.nsr_sample_pdf(b.data_ptr(), ww.data_ptr(), n, B, N, uu.data_ptr(), out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs()
    width = (bins[:, 1:] - bins[:, :-1]).max().item()
    # RH:239's `denom < 1e-5` branch may flip on a last-bit difference of the normaliser (sum order differs from ATen's
    # for B-1 >= 512 only; below that the kernel reproduces it): allow a handful of in-bin moves, nothing larger
    assert (err > 1e-5).sum().item() <= 3 and err.max().item() <= width + 1e-6
    det = nsr.sample_pdf(b, ww, N, det=True).cpu()
    assert (det - O.sample_pdf(bins, w, N, det=True)).abs().max().item() <= width + 1e-6
    assert ((det - O.sample_pdf(bins, w, N, det=True)).abs() > 1e-5).sum().item() <= 3


@pytest.mark.parametrize('S,Ni', [(8, 8), (16, 64), (64, 32), (96, 96), (128, 128)])
def test_render_rays_other_sample_counts(nsr, wfit, nets, S, Ni):
    H = W = 400
    c2w = O.pose_spherical(90., 292.5 - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    sel = torch.arange(150 * 400 + 100, 150 * 400 + 300, 2)
    rays = O.pack_rays(ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel], O.YCBV_NEAR, O.YCBV_FAR)
    with torch.no_grad():
        ref = O.render_rays(rays, wfit[0], wfit[1], S, Ni)
        got = nsr.render_rays(rays.cuda(), nets[0], None, S, N_importance=Ni, network_fine=nets[1])
    for k in ('rgb_map', 'acc_map', 'rgb0', 'acc0'):
        assert rel(got[k], ref[k]) <= 1e-3, (k, S, Ni)


def test_unsupported_sizes_are_errors_not_garbage(nsr, nets):
    rays = torch.zeros(4, 11, device='cuda')
    rays[:, 5] = -1.0
    rays[:, 7] = 1.0
    rays[:, 10] = -1.0
    with pytest.raises(nsr.NsrError):
        nsr.render_rays(rays, nets[0], None, 300, N_importance=0)        # > 256 samples per ray: not built
    with pytest.raises(nsr.NsrError):
        nsr.render_rays(rays, nets[0], None, 200, N_importance=100, network_fine=nets[1])
    with pytest.raises(nsr.NsrError):   # one sample: the reference's dists tensor is empty (RN:358-359), not reproduced
        nsr.raw2outputs(torch.zeros(4, 1, 4, device='cuda'), torch.ones(4, 1, device='cuda'), rays[:, 3:6].contiguous())


@pytest.mark.parametrize('S', [5, 64, 192])
def test_raw2outputs_backward_kernel_vs_autograd(nsr, S):
    """The compositing backward alone (dL/draw, dL/d||d||) against autograd through the oracle's raw2outputs."""
    g = torch.Generator().manual_seed(S + 7)
    n = 150
    raw = (torch.randn(n, S, 4, generator=g) * 2).requires_grad_(True)
    with torch.no_grad():
        raw[:, :, 3] *= 10
    z = torch.sort(torch.rand(n, S, generator=g) * 1.6 + 0.3, -1).values
    rays = torch.randn(n, 11, generator=g)
    dnorm = rays[:, 3:6].norm(dim=-1).clone().requires_grad_(True)
    unit = (rays[:, 3:6] / rays[:, 3:6].norm(dim=-1, keepdim=True)).detach()
    gout = torch.randn(n, 3, generator=g)
    rgb = O.raw2outputs(raw, z, unit * dnorm[:, None])[0]
    ref_raw, ref_dn = torch.autograd.grad(rgb, (raw, dnorm), grad_outputs=gout)
    L = nsr.lib()
    d_raw = torch.empty(n, S, 4, device='cuda')
    # reach the kernel through the public backward entry: needs a network, so call the stage through ctypes-free torch path:
    # nsr_render_rays_backward runs compositing-backward first and leaves dL/draw at the head of its workspace.
    import neural_sim_nerf_b200 as m
    net = m.NeRF().cuda()
    T = S
    wsb = L.nsr_render_backward_workspace_bytes(n, T)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    d_rays = torch.empty(n, 11, device='cuda')
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    keep = [rays.cuda(), z.cuda(), raw.detach().cuda().contiguous(), gout.cuda(), m.packed_weights(net)]   # keep the device buffers alive
    rc = L.nsr_render_rays_backward(P(keep[0]), P(keep[1]), P(keep[2]), n, T, P(keep[4]), 0,
                                    P(keep[3]), P(d_rays), None, None, None, P(ws), wsb, None)
    assert rc == 0, L.nsr_last_error()
    torch.cuda.synchronize()
    got_raw = ws[:n * T * 16].view(torch.float32).view(n, T, 4).cpu()
    scale = ref_raw.abs().max().item()
    assert (got_raw - ref_raw).abs().max().item() <= 2e-5 * scale
