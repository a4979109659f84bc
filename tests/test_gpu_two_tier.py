"""Two-tier evaluation (include/nsr_b200.h "active set"): a cheap first tier certifies the empty sample points (sigma <= 0 gives
alpha == 0 and weight == 0 EXACTLY, RN:356 / RN:376-377), the default fp16 hi/lo arithmetic runs on the others only.

The contract tested here is stronger than the 1e-3 parity bar: every map output is BIT-IDENTICAL to evaluating every point
with the default arithmetic (NSR_FLAG_DENSE), on the fitted scene, on scaled random networks (where nearly every point is
active), when the run-time verification fails and everything is re-evaluated, and when tier 1 is skipped.  The dense route
itself is what tests/test_gpu_parity.py pins to the reference at 1e-3.
"""
import ctypes

import numpy as np
import pytest
import torch

import nerf_oracle as O

pytestmark = pytest.mark.gpu

DENSE = 32
S, NI = 64, 128
T = S + NI


@pytest.fixture(scope='module')
def nsr():
    import neural_sim_nerf_b200 as m
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return m


def module_from_sd(nsr, sd):
    net = nsr.NeRF()
    net.load_state_dict(sd)
    return net.cuda()


@pytest.fixture(scope='module')
def nets(nsr, wfit):
    return module_from_sd(nsr, wfit[0]), module_from_sd(nsr, wfit[1])


@pytest.fixture()
def knobs(nsr):
    """restores the process-wide two-tier knobs after the test"""
    L = nsr.lib()
    en, tau, ver, frc = ctypes.c_int(), ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    L.nsr_get_two_tier(ctypes.byref(en), ctypes.byref(tau), ctypes.byref(ver), ctypes.byref(frc))
    yield L
    assert L.nsr_set_two_tier(en.value, tau.value, ver.value, frc.value) == 0


def camera_rays(n_side, phi):
    H = W = 400
    c2w = O.pose_spherical(90., phi - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, O.YCBV_K_400, c2w)
    ii = torch.linspace(0, 399, n_side).long()
    sel = (ii[:, None] * W + ii[None, :]).reshape(-1)
    return O.pack_rays(ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel], O.YCBV_NEAR, O.YCBV_FAR).cuda()


P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())


def forward(nsr, rays, pc, pf, flags, want_raw=False, with_active=False, s=S, ni=NI, t_rand=None, u=None):
    """nsr_render_rays_forward_ex through ctypes -> dict of outputs (+ ctrl words of the last pass when with_active)"""
    L = nsr.lib()
    n = rays.shape[0]
    t = s + ni
    new = lambda *sh: torch.full(sh, float('nan'), device='cuda')
    out = dict(rgb=new(n, 3), disp=new(n), acc=new(n), rgb0=new(n, 3), disp0=new(n), acc0=new(n), zstd=new(n), w=new(n, t), z=new(n, t))
    raw = new(n, t, 4) if want_raw else None
    ws = torch.zeros(L.nsr_render_workspace_bytes(n, s, ni), dtype=torch.uint8, device='cuda')
    aset = torch.zeros(L.nsr_active_set_bytes(n, t), dtype=torch.uint8, device='cuda') if with_active else None
    rc = L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), s, ni, flags, P(t_rand), P(u), P(out['rgb']), P(out['disp']), P(out['acc']),
                                      P(out['rgb0']) if ni else None, P(out['disp0']) if ni else None, P(out['acc0']) if ni else None,
                                      P(out['zstd']) if ni else None, P(raw), P(out['z']), P(out['w']), None, None, P(aset), P(ws), ws.numel(), None)
    assert rc == 0, L.nsr_last_error()
    torch.cuda.synchronize()
    if ni == 0:
        for k in ('rgb0', 'disp0', 'acc0', 'zstd'):
            out.pop(k)
    if want_raw:
        out['raw'] = raw
    if with_active:
        out['ctrl'] = aset[:64].view(torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        out['aset'] = aset
    # control words of the coarse / fine passes kept in the workspace
    off = (ctypes.c_size_t * 8)()
    assert L.nsr_render_workspace_layout(n, s, ni, off, 8) == 8
    out['ctrl_ws'] = [ws[o:o + 64].view(torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF for o in (off[5], off[6])]
    return out


def assert_maps_identical(a, b, what):
    for k in ('rgb', 'disp', 'acc', 'rgb0', 'disp0', 'acc0', 'zstd', 'w', 'z'):
        if k not in a:
            continue
        x, y = a[k], b[k]
        same = (x == y) | (torch.isnan(x) & torch.isnan(y))
        assert bool(same.all()), f'{what}: {k} differs in {int((~same).sum())} of {same.numel()} entries (max |d| {float(torch.nan_to_num(x - y).abs().max()):.3e})'


def bits_to_float(u):
    return float(np.array([u], dtype=np.uint32).view(np.float32)[0])


def test_maps_bit_identical_to_dense_on_the_fitted_scene(nsr, nets, knobs):
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    for phi in (22.5, 202.5):
        rays = camera_rays(48, phi)
        n = rays.shape[0]
        dense = forward(nsr, rays, pc, pf, DENSE)
        two = forward(nsr, rays, pc, pf, 0)
        assert_maps_identical(two, dense, f'phi={phi}')
        c0, c1 = two['ctrl_ws']
        f0, f1 = c0[0] / (n * S), c1[0] / (n * T)
        v0, v1 = bits_to_float(c0[1]), bits_to_float(c1[1])
        print(f'phi={phi}: active fraction coarse {f0:.4f} fine {f1:.4f}; max |sigma~ - sigma| on active points coarse {v0:.3e} fine {v1:.3e}')
        assert 0 < f0 < 0.2 and 0 < f1 < 0.6, 'the fitted scene is mostly empty space'
        assert v0 < 1.0 and v1 < 1.0, 'tier 1 must stay far inside tau = 4 on the points where both tiers ran'
        assert c0[2] == 0 and c0[3] == 0 and c1[2] == 0 and c1[3] == 0, 'no dense fallback expected here'
        assert dense['ctrl_ws'][1][0] == 0, 'NSR_FLAG_DENSE must not touch the active set'


def test_tier1_sigma_is_a_safe_certificate(nsr, nets, knobs):
    """Directly: for every point the two-tier route left at (0,0,0,sigma~) the dense sigma is negative as well, with a wide margin
    (the certificate is sigma~ <= -4); active points carry exactly the dense raw."""
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rays = camera_rays(40, 112.5)
    dense = forward(nsr, rays, pc, pf, DENSE, want_raw=True)
    two = forward(nsr, rays, pc, pf, 0, want_raw=True, with_active=True)
    assert_maps_identical(two, dense, 'with raw + active set')
    n = rays.shape[0]
    count = int(two['ctrl'][0])
    lst = two['aset'][256:256 + 4 * count].view(torch.int32).long()
    assert lst.numel() == count and int(lst.unique().numel()) == count, 'active list holds each point once'
    active = torch.zeros(n * T, dtype=torch.bool, device='cuda')
    active[lst] = True
    rd, rt = dense['raw'].reshape(-1, 4), two['raw'].reshape(-1, 4)
    assert torch.equal(rt[active], rd[active]), 'tier 2 = the dense arithmetic, bit for bit'
    empty = ~active
    assert bool((rt[empty, :3] == 0).all()) and bool((rt[empty, 3] <= -4.0).all())
    worst = float(rd[empty, 3].max())
    gap = float((rt[empty, 3] - rd[empty, 3]).abs().max())
    print(f'{int(empty.sum())} certified-empty points of {n * T}: largest dense sigma among them {worst:.3f}, max |sigma~ - sigma| {gap:.3e}')
    assert worst < -3.0 and gap < 0.5


def test_verification_failure_re_evaluates_everything(nsr, nets, knobs):
    L = knobs
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rays = camera_rays(32, 67.5)
    assert L.nsr_set_two_tier(1, 4.0, 1e-12, 10.0) == 0          # any difference between the tiers counts as a failure
    # hierarchical: the coarse pass fails its verification and is re-evaluated densely; that in turn makes the fine pass skip tier 1
    dense = forward(nsr, rays, pc, pf, DENSE, want_raw=True)
    two = forward(nsr, rays, pc, pf, 0, want_raw=True, with_active=True)
    assert_maps_identical(two, dense, 'after the dense re-evaluation')
    assert two['ctrl_ws'][0][3] == 1 and two['ctrl_ws'][0][0] > 0, 'the coarse pass must have been re-evaluated'
    assert two['ctrl'][2] == 1 and two['ctrl'][3] == 0 and two['ctrl'][0] == 0, 'the fine pass must have run densely from the start'
    assert torch.equal(two['raw'], dense['raw']), 'raw is the dense raw everywhere'
    # single pass (N_importance = 0): the re-evaluation kernel itself
    dense = forward(nsr, rays, pc, None, DENSE, want_raw=True, ni=0)
    two = forward(nsr, rays, pc, None, 0, want_raw=True, with_active=True, ni=0)
    assert_maps_identical(two, dense, 'single pass, re-evaluated')
    assert two['ctrl'][3] == 1 and two['ctrl'][2] == 0 and 0 < two['ctrl'][0] < rays.shape[0] * S
    assert torch.equal(two['raw'], dense['raw'])


def test_fine_pass_skips_tier1_when_the_coarse_pass_is_mostly_active(nsr, nets, knobs):
    L = knobs
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rays = camera_rays(32, 157.5)
    dense = forward(nsr, rays, pc, pf, DENSE)
    assert L.nsr_set_two_tier(1, 4.0, 1.0, 0.0) == 0             # any active coarse point forces the fine pass dense
    two = forward(nsr, rays, pc, pf, 0, with_active=True)
    assert_maps_identical(two, dense, 'forced-dense fine pass')
    assert two['ctrl'][2] == 1 and two['ctrl'][0] == 0 and two['ctrl_ws'][0][2] == 0


def test_random_networks_where_nearly_everything_is_active(nsr, knobs):
    sdc, sdf = O.random_state_dict(21, scale=3.0), O.random_state_dict(22, scale=3.0)
    for sd in (sdc, sdf):
        sd['alpha_linear.bias'] += 2.0
    mc, mf = module_from_sd(nsr, sdc), module_from_sd(nsr, sdf)
    pc, pf = nsr.packed_weights(mc), nsr.packed_weights(mf)
    rays = camera_rays(24, 22.5)
    dense = forward(nsr, rays, pc, pf, DENSE)
    two = forward(nsr, rays, pc, pf, 0)
    assert_maps_identical(two, dense, 'scaled random networks')
    c0, c1 = two['ctrl_ws']
    print(f'random x3: coarse active {c0[0] / (rays.shape[0] * S):.3f}, fine dense flag {c1[2]}, coarse max |dsigma| {bits_to_float(c0[1]):.3e}')
    # a huge tau makes every point active through the list route as well
    assert knobs.nsr_set_two_tier(1, 1e30, 1.0, 10.0) == 0
    two = forward(nsr, rays, pc, pf, 0, with_active=True)
    assert_maps_identical(two, dense, 'tau = 1e30')
    assert two['ctrl'][0] == rays.shape[0] * T


def test_other_geometries_and_flags(nsr, nets, knobs):
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rays = camera_rays(20, 292.5)[:333]                          # ragged: 333 rays, last tiles partly filled
    n = rays.shape[0]
    g = torch.Generator(device='cuda').manual_seed(4)
    cases = [dict(s=64, ni=0), dict(s=33, ni=17), dict(s=64, ni=128, t_rand=torch.rand(n, 64, device='cuda', generator=g), u=torch.rand(n, 128, device='cuda', generator=g))]
    for kw in cases:
        for fl in (0, 2, 1):                                     # plain, white_bkgd, lindisp
            dense = forward(nsr, rays, pc, pf, DENSE | fl, **kw)
            two = forward(nsr, rays, pc, pf, fl, **kw)
            assert_maps_identical(two, dense, f'{ {k: v for k, v in kw.items() if k in ("s", "ni")} } flags={fl}')
    # zero rays / disabling
    assert nsr.lib().nsr_render_rays_forward_ex(None, 0, P(pc), P(pf), 64, 128, 0, *([None] * 15), None, 0, None) == 0
    assert knobs.nsr_set_two_tier(0, 4.0, 1.0, 0.3) == 0
    off = forward(nsr, rays, pc, pf, 0)
    assert off['ctrl_ws'][1][0] == 0 and off['ctrl_ws'][0][0] == 0
    assert knobs.nsr_set_two_tier(1, 1.0, 2.0, 0.3) != 0, 'verify_max >= tau must be refused'


def test_backward_over_the_active_set_equals_the_dense_backward(nsr, nets, knobs):
    """dL/draw is exactly 0 at every certified-empty point, so back-propagating the active points only gives the same dL/d(rays)."""
    L = knobs
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rays = camera_rays(36, 247.5)
    n = rays.shape[0]
    new = lambda *sh: torch.empty(*sh, device='cuda')
    g = torch.randn(n, 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(11))

    def run(flags, use_mask, use_active, ni=NI):
        t = S + ni
        net = pf if ni else pc
        rgb, raw, zv = new(n, 3), new(n, t, 4), new(n, t)
        ws = torch.empty(L.nsr_render_workspace_bytes(n, S, ni), dtype=torch.uint8, device='cuda')
        bws = torch.empty(L.nsr_render_backward_workspace_bytes(n, t), dtype=torch.uint8, device='cuda')
        mask = torch.empty(L.nsr_relu_mask_bytes(n, t), dtype=torch.uint8, device='cuda') if use_mask else None
        aset = torch.zeros(L.nsr_active_set_bytes(n, t), dtype=torch.uint8, device='cuda') if use_active else None
        assert L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf) if ni else None, S, ni, flags, None, None, P(rgb), None, None, None, None, None, None,
                                            P(raw), P(zv), None, P(mask), None, P(aset), P(ws), ws.numel(), None) == 0, L.nsr_last_error()
        d = new(n, 11)
        assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, t, P(net), 0, P(g), P(d), None, None, None, P(mask), P(aset), P(bws), bws.numel(),
                                             None) == 0, L.nsr_last_error()
        torch.cuda.synchronize()
        ctrl = aset[:16].view(torch.int32).cpu().tolist() if use_active else None
        return rgb, d, ctrl

    def same(d, ref, what):
        scale = float(ref.abs().max())
        err = float((d - ref).abs().max())
        print(f'{what}: max|ref| {scale:.3e}, max diff {err:.3e}, bit-equal {torch.equal(d, ref)}')
        assert scale > 0 and err <= 1e-6 * scale, what

    for ni in (NI, 0):
        t = S + ni
        rgb_d, d_dense, _ = run(DENSE, False, False, ni)
        for use_mask in (True, False):
            rgb_a, d_act, ctrl = run(0, use_mask, True, ni)
            assert torch.equal(rgb_a, rgb_d)
            assert 0 < ctrl[0] < 0.6 * n * t and ctrl[2] == 0 and ctrl[3] == 0
            same(d_act, d_dense, f'active-set backward, Ni={ni}, {"saved bits" if use_mask else "recompute"}, active fraction {ctrl[0] / (n * t):.3f}')
        # failed verification -> the backward pass must follow the forward pass to the dense point order
        assert L.nsr_set_two_tier(1, 4.0, 1e-12, 10.0) == 0
        try:
            for use_mask in (True, False):
                _, d_redo, ctrl = run(0, use_mask, True, ni)
                assert (ctrl[2] == 1 and ctrl[3] == 0) if ni else (ctrl[2] == 0 and ctrl[3] == 1), ctrl   # fine: forced dense by the coarse failure
                same(d_redo, d_dense, f'dense fallback, Ni={ni}, {"saved bits" if use_mask else "recompute"}')
        finally:
            assert L.nsr_set_two_tier(1, 4.0, 1.0, 0.3) == 0
    # parameter gradients are refused on the active-set route
    dW = (ctypes.c_void_p * 12)(*[0] * 12)
    raw, zv = new(n, T, 4), new(n, T)
    bws = torch.empty(L.nsr_render_backward_workspace_bytes(n, T), dtype=torch.uint8, device='cuda')
    aset = torch.zeros(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device='cuda')
    assert L.nsr_render_rays_backward_ex(P(rays), P(zv), P(raw), n, T, P(pf), 0, P(g), P(new(n, 11)), P(bws), dW, dW, None, P(aset), P(bws), bws.numel(), None) != 0


def test_python_surface(nsr, nets, wfit, knobs):
    """render_rays(): default precision = two-tier; retraw=True hands out the dense raw; the pose-only autograd route
    back-propagates over the active set and matches the dense route and the oracle."""
    rays = camera_rays(16, 202.5)
    with torch.no_grad():
        a = nsr.render_rays(rays, nets[0], None, S, N_importance=NI, network_fine=nets[1], retraw=True)
        b = nsr.render_rays(rays, nets[0], None, S, N_importance=NI, network_fine=nets[1])
        nsr.set_precision('fp16x3-dense')
        try:
            c = nsr.render_rays(rays, nets[0], None, S, N_importance=NI, network_fine=nets[1], retraw=True)
        finally:
            nsr.set_precision('fp16x3')
    for k in b:
        same = (a[k] == b[k]) | (torch.isnan(a[k]) & torch.isnan(b[k]))
        assert bool(same.all()), k
    assert torch.equal(a['raw'], c['raw'])
    import copy
    frozen = [copy.deepcopy(m).requires_grad_(False) for m in nets]
    g = torch.randn(rays.shape[0], 3, device='cuda', generator=torch.Generator(device='cuda').manual_seed(5))
    grads = {}
    for mode in ('fp16x3', 'fp16x3-dense'):
        nsr.set_precision(mode)
        try:
            r = rays.clone().requires_grad_(True)
            out = nsr.render_rays(r, frozen[0], None, S, N_importance=NI, network_fine=frozen[1])
            (grads[mode],) = torch.autograd.grad(out['rgb_map'], r, grad_outputs=g)
        finally:
            nsr.set_precision('fp16x3')
    scale = float(grads['fp16x3-dense'].abs().max())
    assert scale > 0 and float((grads['fp16x3'] - grads['fp16x3-dense']).abs().max()) <= 1e-6 * scale


def test_full_size_image_bit_identical(nsr, nets, knobs):
    """BASELINE config 2 at full size (160 000 rays, 64+128): two-tier == dense on every map, and the active fractions of this
    workload (what bench.py's speed-up rests on)."""
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
    n = rays.shape[0]
    dense = forward(nsr, rays, pc, pf, DENSE)
    two = forward(nsr, rays, pc, pf, 0)
    assert_maps_identical(two, dense, '400x400')
    c0, c1 = two['ctrl_ws']
    print(f'400x400: active fraction coarse {c0[0] / (n * S):.4f} fine {c1[0] / (n * T):.4f}; verification {bits_to_float(c0[1]):.3e} / {bits_to_float(c1[1]):.3e}')
    assert c1[2] == 0 and c1[3] == 0


@pytest.mark.parametrize('n_side', [23, 40])
def test_tier1_as_cta_pairs_gives_the_same_bits(nsr, nets, n_side):
    """nsr_set_tier1_pair(1): tier 1 runs as clusters of two CTAs (tcgen05 cta_group::2, half of every weight chunk per SM, the
    leader CTA issuing for both).  Same arithmetic, so the same sigma~, the same active set (as a set) and the same maps, bit for bit;
    an odd number of tiles exercises the padded last pair."""
    L = nsr.lib()
    rays = camera_rays(n_side, 22.5)
    n = rays.shape[0]
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    res = {}
    old = L.nsr_set_tier1_pair(0)
    try:
        for pair in (0, 1):
            L.nsr_set_tier1_pair(pair)
            outs = [torch.empty(n, 3, device='cuda'), torch.empty(n, device='cuda'), torch.empty(n, device='cuda'), torch.empty(n, device='cuda')]
            raw, zv = torch.empty(n, T, 4, device='cuda'), torch.empty(n, T, device='cuda')
            aset = torch.zeros(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device='cuda')
            ws = torch.empty(L.nsr_render_workspace_bytes(n, S, NI), dtype=torch.uint8, device='cuda')
            rc = L.nsr_render_rays_forward_ex(P(rays), n, P(pc), P(pf), S, NI, 0, None, None, P(outs[0]), P(outs[1]), P(outs[2]), None, None, None,
                                              P(outs[3]), P(raw), P(zv), None, None, None, P(aset), P(ws), ws.numel(), None)
            assert rc == 0, L.nsr_last_error()
            torch.cuda.synchronize()
            cnt = int(aset[:4].view(torch.int32).item())
            lst = torch.sort(aset[256:256 + 4 * cnt].view(torch.int32)).values
            res[pair] = (outs, raw, zv, cnt, lst)
    finally:
        L.nsr_set_tier1_pair(old)
    a, b = res[0], res[1]
    assert a[3] == b[3] and a[3] > 0 and torch.equal(a[4], b[4])
    assert ((n * S + 127) // 128) % 2 == (1 if n_side == 23 else 0)          # 23 x 23 rays: the coarse pass has an odd tile count
    for x, y in zip(a[0], b[0]):
        assert bool(((x == y) | (torch.isnan(x) & torch.isnan(y))).all())
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
