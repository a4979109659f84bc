"""GPU parity of the backward pass: dL/d(rays) from dL/d(rgb_map), against PyTorch autograd
through the CPU oracle (the reference's own tape, RN:168-181, is autograd over the same ops).

Tolerance (check_grad): every element within 1e-3 * (max|grad| + |grad|) of the reference gradient, the largest deviation within
1e-3 * max|grad| of the same tensor (gradients of a sharp scene span many decades), cosine similarity > 0.99999.
"""
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


def camera_rays(n_side, phi):
    H = W = 400
    c2w = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    ii = torch.linspace(0, 399, n_side).long()
    sel = (ii[:, None] * W + ii[None, :]).reshape(-1)
    return O.pack_rays(ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel], O.YCBV_NEAR, O.YCBV_FAR)


def check_grad(got, ref, what, tol=1e-3):
    """north_star's 1e-3 on gradients: every element within tol * (max|ref| + |ref|) of fp32 autograd, the largest deviation within
    tol * max|ref|, and the direction right (cosine)."""
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    scale = ref.abs().max().item()
    diff = (got - ref).abs()
    err = diff.max().item()
    cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), ref.reshape(1, -1)).item()
    print(f'{what}: max|ref| {scale:.3e}  max err {err:.3e} ({err / max(scale, 1e-30):.2e} rel)  cos {cos:.8f}')
    assert err <= tol * scale, f'{what}: {err:.3e} > {tol} * {scale:.3e}'
    assert bool((diff <= tol * scale + tol * ref.abs()).all()), f'{what}: element-wise bound'
    assert cos > 0.99999, f'{what}: cosine {cos}'


@pytest.mark.parametrize('phi,n_side,white', [(22.5, 14, False), (200.0, 10, True)])
def test_render_rays_backward_matches_autograd(nsr, wfit, nets, phi, n_side, white):
    rays = camera_rays(n_side, phi)
    n = rays.shape[0]
    g = torch.randn(n, 3, generator=torch.Generator().manual_seed(1))
    # reference: autograd through the oracle (fine pass only carries gradient: z_samples is detached, RN:475)
    r_cpu = rays.clone().requires_grad_(True)
    ref = O.render_rays(r_cpu, wfit[0], wfit[1], 64, 128, white_bkgd=white)
    (ref_grad,) = torch.autograd.grad(ref['rgb_map'], r_cpu, grad_outputs=g)
    r_gpu = rays.cuda().requires_grad_(True)
    out = nsr.render_rays(r_gpu, nets[0], None, 64, N_importance=128, network_fine=nets[1], white_bkgd=white, retraw=True)
    (got_grad,) = torch.autograd.grad(out['rgb_map'], r_gpu, grad_outputs=g.cuda())
    assert got_grad.shape == (n, 11)
    check_grad(got_grad[:, 0:3], ref_grad[:, 0:3], 'dL/d rays_o')
    check_grad(got_grad[:, 3:6], ref_grad[:, 3:6], 'dL/d rays_d')
    check_grad(got_grad[:, 8:11], ref_grad[:, 8:11], 'dL/d viewdirs')
    assert float(got_grad[:, 6:8].abs().max()) == 0.0


def test_render_backward_through_get_rays_to_c2w(nsr, wfit, nets):
    """The reference's use (RN:148-181): rays come from get_rays(c2w(psi)); chain dL/drays on to the pose."""
    H = W = 20
    K = [[66.0, 0, 9.5], [0, 66.0, 10.5], [0, 0, 1]]
    pose = O.pose_spherical(90., 157.5 - 180., 1.01)[:3, :4]
    g = torch.randn(H * W, 3, generator=torch.Generator().manual_seed(2))
    kw = dict(N_samples=64, N_importance=128)

    c_cpu = pose.clone().requires_grad_(True)
    ro, rd = O.get_rays(H, W, K, c_cpu)
    ref = O.render(H, W, K, wfit[0], wfit[1], chunk=512, rays=(ro.reshape(-1, 3), rd.reshape(-1, 3)), near=O.YCBV_NEAR, far=O.YCBV_FAR, **kw)
    (ref_grad,) = torch.autograd.grad(ref[0], c_cpu, grad_outputs=g)

    c_gpu = pose.cuda().requires_grad_(True)
    ro, rd = nsr.get_rays(H, W, K, c_gpu)
    batch_rays = torch.stack([ro.reshape(-1, 3), rd.reshape(-1, 3)], 0)                      # RN:163
    rgb, _, _, _ = nsr.render(H, W, K, chunk=512, rays=batch_rays, retraw=True, network_fn=nets[0], network_query_fn=None,
                              network_fine=nets[1], use_viewdirs=True, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, **kw)
    dLdray = torch.autograd.grad(rgb, batch_rays, grad_outputs=g.cuda(), retain_graph=True)  # RN:177-178
    (got_grad,) = torch.autograd.grad(batch_rays, c_gpu, grad_outputs=dLdray)                 # RN:179-181
    check_grad(got_grad, ref_grad, 'dL/d c2w')


def test_coarse_only_backward_and_unsupported_grads(nsr, wfit, nets):
    rays = camera_rays(8, 60.0)
    g = torch.randn(rays.shape[0], 3, generator=torch.Generator().manual_seed(3))
    r_cpu = rays.clone().requires_grad_(True)
    ref = O.render_rays(r_cpu, wfit[0], None, 64, 0)
    (ref_grad,) = torch.autograd.grad(ref['rgb_map'], r_cpu, grad_outputs=g)
    r_gpu = rays.cuda().requires_grad_(True)
    out = nsr.render_rays(r_gpu, nets[0], None, 64, N_importance=0)
    (got_grad,) = torch.autograd.grad(out['rgb_map'], r_gpu, grad_outputs=g.cuda())
    check_grad(got_grad[:, 0:6], ref_grad[:, 0:6], 'coarse-only dL/d(o,d)')
    assert not out['acc_map'].requires_grad and not out['disp_map'].requires_grad   # not built -> not differentiable, loudly


def decode_relu_bits(mask_u8, n_points):
    """The forward pass's saved ReLU states (common.cuh MASK_*: per 128-point tile 68 words x 128 rows; word l*8+w = columns
    [32w, 32w+32) of pts_linears.l, word 64+w of views_linears.0; bit j = column 2j, bit 16+j = column 2j+1)
    -> (layers [P,8,256], views [P,128]) as float64 0/1."""
    w = mask_u8.view(torch.int32).view(-1, 68, 128)
    j = torch.arange(16, device=w.device)
    even = (w[..., None] >> j) & 1
    odd = (w[..., None] >> (16 + j)) & 1
    bits = torch.stack([even, odd], -1).reshape(w.shape[0], 68, 128, 32).permute(0, 2, 1, 3).reshape(w.shape[0] * 128, 68 * 32)[:n_points]
    return bits[:, :2048].reshape(n_points, 8, 256).double(), bits[:, 2048:2176].double()


def conditioned_render(rays, z, sd, relu_layers, relu_views, sigma_on, white=False):
    """RN:26-40 + RH:99-122 + RN:343-387 in float64 with every ReLU's on/off state GIVEN (h = a * state) instead of decided by the
    sign of a: the function the backward kernels differentiate.  Returns rgb_map, the pre-activations and their noise scales
    (sum |w||h| + |b|, what an fp32 rounding error of the dot product is proportional to)."""
    lin = torch.nn.functional.linear
    n, T = z.shape
    r = rays.double()
    pts = (r[:, None, 0:3] + r[:, None, 3:6] * z.double()[:, :, None]).reshape(-1, 3)
    ex = O.embed(pts, O.N_FREQ_XYZ)
    ev = O.embed(r[:, None, 8:11].expand(n, T, 3).reshape(-1, 3), O.N_FREQ_DIR)
    pre, scale = [], []
    h = ex
    for i in range(8):
        W, b = sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias']
        a = lin(h, W, b)
        pre.append(a)
        scale.append(lin(h.detach().abs(), W.detach().abs(), b.detach().abs()))
        h = a * relu_layers[:, i]
        if i == O.SKIP_AFTER:
            h = torch.cat([ex, h], -1)
    sigma = lin(h, sd['alpha_linear.weight'], sd['alpha_linear.bias'])
    feat = lin(h, sd['feature_linear.weight'], sd['feature_linear.bias'])
    hin = torch.cat([feat, ev], -1)
    W, b = sd['views_linears.0.weight'], sd['views_linears.0.bias']
    av = lin(hin, W, b)
    pre.append(av)
    scale.append(lin(hin.detach().abs(), W.detach().abs(), b.detach().abs()))
    rgb = lin(av * relu_views, sd['rgb_linear.weight'], sd['rgb_linear.bias'])
    raw = torch.cat([rgb, sigma * sigma_on.reshape(-1, 1)], -1).reshape(n, T, 4)      # RN:356's relu with its state given, too
    out = O.raw2outputs(raw, z.double(), r[:, 3:6], None, white)
    return out[0], pre, scale, sigma.reshape(n, T)


def test_parameter_gradients_match_autograd(nsr, wfit):
    """loss = mse(rgb, target) + mse(rgb0, target) (RN:691-696), loss.backward() -> dL/dMLP for BOTH networks,
    every element within 1e-3 * max|grad| of the float64 gradient.

    What "the gradient" is needs care.  dL/dW is discontinuous in the last bits of the forward pass: a ReLU whose pre-activation
    is within fp32 rounding noise of zero is ON in one correct fp32 evaluation and OFF in another, and one such unit at a point that
    carries a large dL/draw moves dL/dW by ~2e-3 of its maximum on this very batch -- torch's own fp32 autograd on the CPU (MKL) and
    on the GPU (cuBLAS) differ from each other by that amount (tools/dump_check.py prints it).  So the test pins exactly what is
    well defined: (1) our gradient equals the float64 gradient of the network WITH THE RELU STATES OUR FORWARD PASS SAW (its saved
    sign bits) to 1e-3 everywhere; (2) those states differ from the float64 signs only on units whose |pre-activation| is below
    1e-5 of its rounding-noise scale sum|w||h|+|b| (measured: 9e-7), and on fewer than 1 unit in 10^5 (measured: 17 of 60 M); (3) the loss itself agrees to 1e-4."""
    import ctypes
    rays = camera_rays(12, 22.5)
    n = rays.shape[0]
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(5))
    nets = []
    for sd in wfit:
        m = nsr.NeRF()
        m.load_state_dict(sd)
        nets.append(m.cuda())
    L = nsr.lib()
    rg = rays.cuda()
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    new = lambda *s: torch.empty(*s, device='cuda')
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])

    def forward_states(S, Ni):
        T = S + Ni
        rgb, raw, zv = new(n, 3), new(n, T, 4), new(n, T)
        ws = torch.empty(L.nsr_render_workspace_bytes(n, S, Ni), dtype=torch.uint8, device='cuda')
        mask = torch.empty(L.nsr_relu_mask_bytes(n, T), dtype=torch.uint8, device='cuda')
        rc = L.nsr_render_rays_forward_ex(P(rg), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw), P(zv),
                                          None, P(mask), None, None, P(ws), ws.numel(), None)
        assert rc == 0, L.nsr_last_error()
        torch.cuda.synchronize()
        return zv, raw, decode_relu_bits(mask, n * T)
    z0, raw0, st0 = forward_states(64, 0)          # the coarse pass (carries gradient through rgb0)
    zf, rawf, stf = forward_states(64, 128)        # the last pass: the fine network on the merged depths
    # ours, through the public render_rays + loss.backward()
    out = nsr.render_rays(rg, nets[0], None, 64, N_importance=128, network_fine=nets[1])
    loss = ((out['rgb_map'] - target.cuda()) ** 2).mean() + ((out['rgb0'] - target.cuda()) ** 2).mean()
    loss.backward()
    # float64 with the given ReLU states
    sdc = {k: v.cuda().double().requires_grad_(True) for k, v in wfit[0].items()}
    sdf = {k: v.cuda().double().requires_grad_(True) for k, v in wfit[1].items()}
    rgb0_r, pre0, sc0, sig0 = conditioned_render(rg, z0, sdc, st0[0], st0[1], (raw0[..., 3] > 0).double())
    rgb_r, pref, scf, sigf = conditioned_render(rg, zf, sdf, stf[0], stf[1], (rawf[..., 3] > 0).double())
    loss_ref = ((rgb_r - target.cuda().double()) ** 2).mean() + ((rgb0_r - target.cuda().double()) ** 2).mean()
    loss_ref.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * max(1.0, abs(loss_ref.item()))
    # (2) the states are the float64 signs except at rounding level
    for tag, pre, sc, st, sig, raw in (('coarse', pre0, sc0, st0, sig0, raw0), ('fine', pref, scf, stf, sigf, rawf)):
        flips, units, worst = 0, 0, 0.0
        for l in range(9):
            a, s = pre[l].detach(), sc[l]
            state = st[0][:, l] if l < 8 else st[1]
            bad = (a > 0).double() != state
            flips += int(bad.sum())
            units += a.numel()
            if bool(bad.any()):
                worst = max(worst, float((a.abs() / s)[bad].max()))
        bad = (sig.detach() > 0) != (raw[..., 3] > 0)
        if bool(bad.any()):
            worst = max(worst, float((sig.detach().abs()[bad]).max() / 60.0))
        print(f'{tag}: {flips + int(bad.sum())} ReLU states of {units} differ from the float64 signs; largest |pre-activation| / noise scale among them {worst:.2e}')
        assert worst <= 1e-5 and flips <= 1e-5 * units
    # (1) gradients
    failures = []
    for net, sd, tag in ((nets[0], sdc, 'coarse'), (nets[1], sdf, 'fine')):
        for name, prm in net.named_parameters():
            assert prm.grad is not None, f'{tag}.{name} got no gradient'
            try:
                check_grad(prm.grad, sd[name].grad, f'{tag}.{name}', tol=1e-3)
            except AssertionError as e:
                failures.append(str(e))
    assert not failures, failures
    # for the record: plain fp32 autograd (the reference's arithmetic, CPU) against the same float64 gradient
    s32 = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in wfit]
    ref = O.render_rays(rays, s32[0], s32[1], 64, 128, z_fine=zf.cpu())
    (((ref['rgb_map'] - target) ** 2).mean() + ((ref['rgb0'] - target) ** 2).mean()).backward()
    worst32 = max(float((s32[1][k].grad.double() - sdf[k].grad.cpu()).abs().max() / sdf[k].grad.abs().max()) for k in sdf)
    worst_us = max(float((p_.grad.double() - sdf[k].grad).abs().max() / sdf[k].grad.abs().max()) for k, p_ in nets[1].named_parameters())
    print(f'fine network, worst tensor: torch fp32 CPU autograd vs state-conditioned float64 {worst32:.2e}; ours {worst_us:.2e}')


def test_parameter_gradients_many_samples_no_resampling(nsr, wfit):
    """Same check with 192 coarse samples and no hierarchical step: isolates operand precision from RH:239 flips."""
    rays = camera_rays(10, 22.5)
    n = rays.shape[0]
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(6))
    sdf = {k: v.clone().requires_grad_(True) for k, v in wfit[1].items()}
    ref = O.render_rays(rays, sdf, None, 192, 0)
    ((ref['rgb_map'] - target) ** 2).mean().backward()
    m = nsr.NeRF()
    m.load_state_dict(wfit[1])
    m = m.cuda()
    out = nsr.render_rays(rays.cuda(), m, None, 192, N_importance=0)
    ((out['rgb_map'] - target.cuda()) ** 2).mean().backward()
    failures = []
    for name, prm in m.named_parameters():
        try:
            check_grad(prm.grad, sdf[name].grad, f'S192.{name}', tol=1e-3)
        except AssertionError as e:
            failures.append(str(e))
    assert not failures, failures


@pytest.mark.parametrize('n_side,S,Ni', [(17, 64, 128), (3, 64, 128), (9, 64, 0), (11, 16, 40)])
def test_saved_sign_bits_backward_equals_recompute(nsr, nets, n_side, S, Ni):
    """nsr_render_rays_forward_ex saves one bit per ReLU; nsr_render_rays_backward_ex with those bits (no forward recompute,
    12 instead of 22 GEMM steps) must return what the recompute route returns."""
    import ctypes
    L = nsr.lib()
    rays = camera_rays(n_side, 22.5).cuda()
    n, T = rays.shape[0], S + Ni
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    new = lambda *s: torch.empty(*s, device='cuda')
    rgb, raw, zv = new(n, 3), new(n, T, 4), new(n, T)
    ws = torch.empty(L.nsr_render_workspace_bytes(n, S, Ni), dtype=torch.uint8, device='cuda')
    mask = torch.full((L.nsr_relu_mask_bytes(n, T),), 0xAA, dtype=torch.uint8, device='cuda')
    assert mask.numel() == ((n * T + 127) // 128) * 68 * 128 * 4
    rc = L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw), P(zv),
                                      None, P(mask), None, None, P(ws), ws.numel(), None)
    assert rc == 0, L.nsr_last_error()
    g = torch.randn(n, 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(3))
    bws = torch.empty(L.nsr_render_backward_workspace_bytes(n, T), dtype=torch.uint8, device='cuda')
    net = pf if Ni > 0 else pc
    d_ref, d_got = new(n, 11), new(n, 11)
    assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(net), 0, P(g), P(d_ref), None, None, None, None, None, P(bws), bws.numel(), None) == 0
    assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(net), 0, P(g), P(d_got), None, None, None, P(mask), None, P(bws), bws.numel(), None) == 0, L.nsr_last_error()
    torch.cuda.synchronize()
    scale = float(d_ref.abs().max())
    err = float((d_got - d_ref).abs().max())
    print(f'saved-bits vs recompute (n={n}, S={S}, Ni={Ni}): max|ref| {scale:.3e}, max diff {err:.3e}, bit-equal {torch.equal(d_got, d_ref)}')
    assert scale > 0 and err <= 1e-6 * scale
    # the bits themselves: a fitted network has both live and dead units
    words = mask.view(torch.int32)
    assert int((words != 0).sum()) > 0 and int((words != -1).sum()) > 0
    # parameter gradients without recompute: the forward pass also dumps the activations (dump_out), the backward pass adds the
    # gradients to the same buffer; dL/dW, dL/db against the recompute route (whose backward kernel fills the whole dump itself)
    def param_grads(use_saved):
        dump = torch.empty(L.nsr_mlp_dump_bytes(n, T), dtype=torch.uint8, device='cuda')
        if use_saved:
            m2 = torch.empty_like(mask)
            rc = L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw),
                                              P(zv), None, P(m2), P(dump), None, P(ws), ws.numel(), None)
            assert rc == 0, L.nsr_last_error()
            assert torch.equal(m2, mask)
        gw = [torch.zeros(s_, device='cuda') for s_ in nsr.run_nerf._EXPECTED_SHAPES]
        gb = [torch.zeros(s_[0], device='cuda') for s_ in nsr.run_nerf._EXPECTED_SHAPES]
        dWp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in gw])
        dBp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in gb])
        d = new(n, 11)
        rc = L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(net), 0, P(g), P(d), P(dump), dWp, dBp, P(mask) if use_saved else None, None,
                                           P(bws), bws.numel(), None)
        assert rc == 0, L.nsr_last_error()
        torch.cuda.synchronize()
        return d, gw, gb
    d_a, gw_a, gb_a = param_grads(False)
    d_b, gw_b, gb_b = param_grads(True)
    assert float((d_a - d_b).abs().max()) <= 1e-6 * float(d_a.abs().max())
    for i, (x, y) in enumerate(list(zip(gw_a, gw_b)) + list(zip(gb_a, gb_b))):
        sc = float(x.abs().max())
        assert sc > 0, i
        assert float((x - y).abs().max()) <= 1e-4 * sc, (i, float((x - y).abs().max()), sc)   # split-K atomics: summation order differs run to run
    with pytest.raises(AssertionError):     # dump_out without relu_mask is refused
        rc = L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw), P(zv),
                                          None, None, P(mask), None, P(ws), ws.numel(), None)
        assert rc == 0


def test_pose_only_autograd_takes_the_saved_bits_route(nsr, wfit, nets):
    """With frozen networks (nothing but the rays asks for gradient) render_rays keeps the sign bits and the backward skips the
    recompute; same gradient as with NSR_SAVE_RELU_MASK=0 and as the oracle's autograd."""
    import copy
    frozen = [copy.deepcopy(m).requires_grad_(False) for m in nets]
    rays = camera_rays(12, 202.5)
    g = torch.randn(rays.shape[0], 3, generator=torch.Generator().manual_seed(5))
    grads = {}
    launches = {}
    nsr.packed_weights(frozen[0]), nsr.packed_weights(frozen[1])      # pack outside the counted region
    for save in (True, False):
        nsr.run_nerf.SAVE_RELU_MASK = save
        try:
            r = rays.cuda().requires_grad_(True)
            before = nsr.lib().nsr_launch_count()
            out = nsr.render_rays(r, frozen[0], None, 64, N_importance=128, network_fine=frozen[1])
            (grads[save],) = torch.autograd.grad(out['rgb_map'], r, grad_outputs=g.cuda())
            launches[save] = nsr.lib().nsr_launch_count() - before
        finally:
            nsr.run_nerf.SAVE_RELU_MASK = True
    assert launches[True] == launches[False]
    scale = float(grads[False].abs().max())
    assert float((grads[True] - grads[False]).abs().max()) <= 1e-6 * scale
    r_cpu = rays.clone().requires_grad_(True)
    ref = O.render_rays(r_cpu, wfit[0], wfit[1], 64, 128)
    (ref_grad,) = torch.autograd.grad(ref['rgb_map'], r_cpu, grad_outputs=g)
    check_grad(grads[True][:, 0:6], ref_grad[:, 0:6], 'dL/d rays (saved sign bits)')


def test_full_size_backward_properties(nsr, nets):
    """BASELINE config 3 at full size (160 000 rays, 64+128 samples), through size-independent properties: the saved-bits and the
    recompute routes agree, the backward is linear in dL/drgb (a power-of-two multiple is exact), rays whose dL/drgb is zero get a
    zero gradient, near/far get none, and everything is finite."""
    import ctypes
    L = nsr.lib()
    H = W = 400
    S, Ni = 64, 128
    T = S + Ni
    rays = nsr.make_rays(H, W, O.YCBV_K_400, O.pose_spherical(90., 67.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
    n = rays.shape[0]
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    new = lambda *s: torch.empty(*s, device='cuda')
    rgb, raw, zv = new(n, 3), new(n, T, 4), new(n, T)
    ws = torch.empty(L.nsr_render_workspace_bytes(n, S, Ni), dtype=torch.uint8, device='cuda')
    mask = torch.empty(L.nsr_relu_mask_bytes(n, T), dtype=torch.uint8, device='cuda')
    assert L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, Ni, 0, None, None, P(rgb), None, None, None, None, None, None, P(raw), P(zv),
                                        None, P(mask), None, None, P(ws), ws.numel(), None) == 0, L.nsr_last_error()
    g = torch.randn(n, 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(9))
    g[::7] = 0.0
    bws = torch.empty(L.nsr_render_backward_workspace_bytes(n, T), dtype=torch.uint8, device='cuda')

    def bwd(gg, m):
        d = new(n, 11)
        assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(pf), 0, P(gg), P(d), None, None, None, P(m), None, P(bws), bws.numel(), None) == 0
        torch.cuda.synchronize()
        return d
    d_saved, d_rec = bwd(g, mask), bwd(g, None)
    assert bool(torch.isfinite(d_saved).all())
    scale = float(d_rec.abs().max())
    assert scale > 0 and float((d_saved - d_rec).abs().max()) <= 2e-6 * scale
    assert torch.equal(bwd(g * 4.0, mask), d_saved * 4.0)            # per-row power-of-two scaling: exact
    assert float(d_saved[::7].abs().max()) == 0.0 and float(d_saved[:, 6:8].abs().max()) == 0.0
    hit = rgb.sum(-1) > 0.05
    assert float(d_saved[hit].abs().max()) > 0 and int(hit.sum()) > 1000
    # closed-form pose pull-back of the full image: deterministic, finite, and it responds to the gradient
    c1 = nsr.rays_grad_to_c2w(H, W, O.YCBV_K_400, rays, d_saved)
    assert torch.equal(c1, nsr.rays_grad_to_c2w(H, W, O.YCBV_K_400, rays, d_saved)) and bool(torch.isfinite(c1).all())
    assert float(c1.abs().max()) > 0
