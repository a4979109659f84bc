"""GPU tests of the fused optimisation step: nsr_random_uniform / nsr_add_sigma_noise (device Philox),
nsr_train_step and its mirror train_step -- against the oracle generator, against render() + autograd + torch.optim.Adam on
the same kernels, and against the CPU oracle renderer + autograd + torch.optim.Adam."""
import copy
import ctypes

import numpy as np
import pytest
import torch

import nerf_oracle as O
from test_gpu_parity import camera_rays, module_from_sd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nsr():
    import neural_sim_nerf_b200 as m
    assert torch.cuda.is_available()
    return m


def P(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize('count', [1, 3, 4, 5, 4099])
def test_random_uniform_equals_oracle_philox(nsr, count):
    L = nsr.lib()
    out = torch.full((count + 2,), -1.0, device='cuda')
    for seed, stream in ((0, 0), (0x123456789abcdef, 1), (2 ** 64 - 1, 3)):
        out.fill_(-1.0)
        assert L.nsr_random_uniform(seed, stream, P(out), count, None) == 0, L.nsr_last_error()
        got = out.cpu().numpy()
        assert np.array_equal(got[:count], O.philox_uniform(seed, stream, count))
        assert (got[count:] == -1.0).all()            # nothing written past `count`


def test_random_uniform_statistics(nsr):
    L = nsr.lib()
    n = 1 << 20
    a, b = torch.empty(n, device='cuda'), torch.empty(n, device='cuda')
    L.nsr_random_uniform(7, 0, P(a), n, None)
    L.nsr_random_uniform(7, 1, P(b), n, None)
    assert 0.0 <= float(a.min()) and float(a.max()) < 1.0
    assert abs(float(a.mean()) - 0.5) < 2e-3 and abs(float(a.var()) - 1 / 12) < 1e-3
    hist = torch.histc(a, bins=64, min=0, max=1)
    chi2 = float(((hist - n / 64) ** 2 / (n / 64)).sum())
    assert chi2 < 130.0                                # 63 degrees of freedom: P(chi2 > 130) ~ 1e-6
    corr = float(((a - 0.5) * (b - 0.5)).mean() * 12)
    assert abs(corr) < 5e-3                            # streams are independent
    assert abs(float(((a[:-1] - 0.5) * (a[1:] - 0.5)).mean() * 12)) < 5e-3


def test_sigma_noise(nsr):
    L = nsr.lib()
    n = 1 << 18
    raw = torch.zeros(n, 4, device='cuda')
    raw[:, :3] = 5.0
    assert L.nsr_add_sigma_noise(11, 2, P(raw), n, 0.5, None) == 0
    assert bool((raw[:, :3] == 5.0).all())
    s = raw[:, 3]
    assert abs(float(s.mean())) < 5e-3 and abs(float(s.std()) - 0.5) < 5e-3
    kurt = float(((s / 0.5) ** 4).mean())
    assert abs(kurt - 3.0) < 0.1                       # Gaussian tails
    again = torch.zeros(n, 4, device='cuda')
    L.nsr_add_sigma_noise(11, 2, P(again), n, 0.5, None)
    assert torch.equal(again[:, 3], s)
    L.nsr_add_sigma_noise(12, 2, P(again), n, 0.5, None)
    assert not torch.equal(again[:, 3], 2 * s)


def train_kwargs(nets, **over):
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True,
              ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=0., lindisp=False)
    kw.update(over)
    return kw


def batch(n_side=12, phi=22.5, seed=0):
    rays = camera_rays(n_side, phi)
    target = torch.rand(rays.shape[0], 3, generator=torch.Generator().manual_seed(seed))
    return torch.stack([rays[:, 0:3], rays[:, 3:6]], 0), target


def max_rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def test_train_step_equals_autograd_and_torch_adam(nsr, wfit):
    """Same kernels underneath: render() + autograd + torch.optim.Adam vs one nsr_train_step call per iteration."""
    nets_a = [module_from_sd(nsr, sd) for sd in wfit]
    nets_b = copy.deepcopy(nets_a)
    opt_a = torch.optim.Adam([p for m in nets_a for p in m.parameters()], lr=5e-4, betas=(0.9, 0.999))
    opt_b = torch.optim.Adam([p for m in nets_b for p in m.parameters()], lr=5e-4, betas=(0.9, 0.999))
    start = [p.detach().clone() for m in nets_a for p in m.parameters()]
    for it in range(3):
        br, tgt = batch(12, 22.5 + 90 * it, it)
        br, tgt = br.cuda(), tgt.cuda()
        kw = train_kwargs(nets_a)
        rgb, _, _, extras = nsr.render(400, 400, O.YCBV_K_400, chunk=1 << 15, rays=br, retraw=True, **kw)
        opt_a.zero_grad()
        img_loss, img_loss0 = nsr.img2mse(rgb, tgt), nsr.img2mse(extras['rgb0'], tgt)
        (img_loss + img_loss0).backward()
        opt_a.step()
        out = nsr.train_step(br, tgt, opt_b, **train_kwargs(nets_b))
        assert abs(float(out['img_loss']) - float(img_loss)) <= 1e-5 * max(1.0, float(img_loss))
        assert abs(float(out['img_loss0']) - float(img_loss0)) <= 1e-5 * max(1.0, float(img_loss0))
        if it == 0:
            assert torch.equal(out['rgb'], rgb.detach())                # identical parameters, identical kernels
        else:
            assert torch.allclose(out['rgb'], rgb.detach(), atol=5e-4)   # parameters differ by summation order of the weight gradients
        assert float(out['psnr']) == pytest.approx(float(nsr.mse2psnr(img_loss.detach())), rel=1e-5)
        if it == 0:     # identical parameters so far: the gradients both routes fed to Adam agree to rounding
            for pa, pb in zip([p for m in nets_a for p in m.parameters()], [p for m in nets_b for p in m.parameters()]):
                assert max_rel(opt_b.state[pb]['exp_avg'], opt_a.state[pa]['exp_avg']) <= 2e-3
                assert max_rel(opt_b.state[pb]['exp_avg_sq'], opt_a.state[pa]['exp_avg_sq']) <= 4e-3
    moved = 0.0
    for (na, pa), (nb, pb), p0 in zip(nets_a[0].named_parameters(), nets_b[0].named_parameters(), start[:24]):
        # the weight-gradient kernels add split-K partial sums with atomics: the two routes agree to rounding, and an Adam update is
        # lr * m / sqrt(v) <= ~lr per step, so compare against the distance travelled
        diff = (pa - pb).abs()
        moved = max(moved, float((pa - p0).abs().max()))
        # (coordinates whose gradient is ~eps-sized turn rounding differences of the two routes -- the split-K atomics add in a different
        # order every run -- into a visible fraction of one lr-sized step: bound the worst case by one step, the bulk tightly)
        assert float(diff.max()) <= 5e-4 and float((diff > 2e-5).float().mean()) <= 0.01, (na, float(diff.max()))
    for pa, pb in zip(nets_a[1].parameters(), nets_b[1].parameters()):
        diff = (pa - pb).abs()
        assert float(diff.max()) <= 5e-4 and float((diff > 2e-5).float().mean()) <= 0.01
    assert moved > 5e-4                                             # the parameters did move (3 steps of lr = 5e-4)
    sa, sb = opt_a.state_dict()['state'], opt_b.state_dict()['state']
    assert set(sa) == set(sb)
    for k in sa:
        assert float(sb[k]['step']) == 3.0 == float(sa[k]['step'])
        # two independent three-step trajectories of a sharp-edged scene: the moments stay close, not identical
        assert max_rel(sb[k]['exp_avg'], sa[k]['exp_avg']) <= 5e-2
        assert max_rel(sb[k]['exp_avg_sq'], sa[k]['exp_avg_sq']) <= 1e-1


def test_train_step_keeps_the_packed_weights_in_step(nsr, wfit):
    """After a fused step the renderer must see the updated parameters (the blobs are re-packed inside the call)."""
    nets = [module_from_sd(nsr, sd) for sd in wfit]
    opt = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=1e-2)
    br, tgt = batch(10, 67.5, 3)
    rays = camera_rays(10, 67.5).cuda()
    with torch.no_grad():
        before = nsr.render_rays(rays, nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map'].clone()
    nsr.train_step(br.cuda(), tgt.cuda(), opt, **train_kwargs(nets))
    with torch.no_grad():
        after = nsr.render_rays(rays, nets[0], None, 64, N_importance=128, network_fine=nets[1])['rgb_map']
        fresh = [module_from_sd(nsr, {k: v.detach().cpu() for k, v in m.state_dict().items()}) for m in nets]
        ref = nsr.render_rays(rays, fresh[0], None, 64, N_importance=128, network_fine=fresh[1])['rgb_map']
    assert not torch.equal(before, after)
    assert torch.equal(after, ref)


def test_train_step_vs_cpu_oracle(nsr, wfit):
    """Two iterations against the oracle renderer + autograd + torch.optim.Adam on the CPU."""
    nets = [module_from_sd(nsr, sd) for sd in wfit]
    opt = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=5e-4)
    sds = [{k: v.clone().requires_grad_(True) for k, v in sd.items()} for sd in wfit]
    names = [n for n, _ in nets[0].named_parameters()]
    opt_c = torch.optim.Adam([sd[n] for sd in sds for n in names], lr=5e-4)
    for it in range(2):
        br, tgt = batch(8, 22.5 + 45 * it, 10 + it)
        rays = O.pack_rays(br[0], br[1], O.YCBV_NEAR, O.YCBV_FAR)
        ref = O.render_rays(rays, sds[0], sds[1], 64, 128)
        opt_c.zero_grad()
        l, l0 = ((ref['rgb_map'] - tgt) ** 2).mean(), ((ref['rgb0'] - tgt) ** 2).mean()
        (l + l0).backward()
        g_ref = {(k, n): sds[k][n].grad.clone() for k in range(2) for n in names}
        opt_c.step()
        out = nsr.train_step(br.cuda(), tgt.cuda(), opt, **train_kwargs(nets))
        assert float(out['img_loss']) == pytest.approx(float(l), rel=2e-3)
        assert float(out['img_loss0']) == pytest.approx(float(l0), rel=2e-3)
        # first step: exp_avg = (1 - beta1) * grad exactly -> read the gradient the kernel used out of the optimiser state
        if it == 0:
            for k in range(2):
                for n, p in nets[k].named_parameters():
                    g = opt.state[p]['exp_avg'] / 0.1
                    scale = float(g_ref[(k, n)].abs().max())
                    if scale > 0:
                        # (the oracle resamples its own fine depths here: RH:239 bin flips add to the arithmetic error, so this end-to-end
                        # bound is looser than the 1e-3 of tests/test_gpu_backward.py, which compares on identical sample positions)
                        assert float((g.cpu() - g_ref[(k, n)]).abs().max()) <= 3e-3 * scale, (k, n)
    # an Adam step moves every coordinate by at most ~lr: after two steps the two trajectories may differ by a fraction of that where
    # a tiny gradient changes sign between the two arithmetic routes; bound the bulk tightly and the worst case by the step size
    for k in range(2):
        for n, p in nets[k].named_parameters():
            d = (p.detach().cpu() - sds[k][n].detach()).abs()
            assert float(d.max()) <= 2.1 * 5e-4, (k, n, float(d.max()))
            assert float((d > 5e-5).float().mean()) <= 0.02, (k, n)


def test_train_step_random_draws_and_options(nsr, wfit):
    def run(seed, **over):
        nets = [module_from_sd(nsr, sd) for sd in wfit]
        opt = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=5e-4)
        br, tgt = batch(9, 112.5, 5)
        out = nsr.train_step(br.cuda(), tgt.cuda(), opt, seed=seed, **train_kwargs(nets, **over))
        return float(out['loss']), out['rgb'].clone(), nets
    l_det, rgb_det, _ = run(1)
    l_a, rgb_a, nets_a = run(1, perturb=1.0, raw_noise_std=1.0)
    l_b, rgb_b, nets_b = run(1, perturb=1.0, raw_noise_std=1.0)
    l_c, rgb_c, _ = run(2, perturb=1.0, raw_noise_std=1.0)
    assert np.isfinite([l_det, l_a, l_c]).all()
    assert torch.equal(rgb_a, rgb_b) and l_a == l_b                # same seed, same step -> same draws
    assert not torch.equal(rgb_a, rgb_c) and not torch.equal(rgb_a, rgb_det)
    assert float((rgb_a - rgb_det).abs().max()) < 0.6              # jitter + sigma noise perturb the image, they do not wreck it
    # coarse-only (N_importance = 0) and white background run and update only the coarse network
    nets = [module_from_sd(nsr, sd) for sd in wfit]
    fine0 = [p.detach().clone() for p in nets[1].parameters()]
    opt = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=5e-4)
    br, tgt = batch(9, 112.5, 5)
    out = nsr.train_step(br.cuda(), tgt.cuda(), opt, **train_kwargs(nets, N_importance=0, white_bkgd=True))
    assert float(out['img_loss0']) == 0.0 and np.isfinite(float(out['loss']))
    assert all(torch.equal(a, b) for a, b in zip(fine0, nets[1].parameters()))
    with pytest.raises(NotImplementedError):
        nsr.train_step(br.cuda(), tgt.cuda(), torch.optim.SGD(nets[0].parameters(), lr=0.1), **train_kwargs(nets))
    with pytest.raises(NotImplementedError):
        nsr.train_step(br.cuda(), tgt.cuda(), torch.optim.Adam(list(nets[0].parameters()) + list(nets[1].parameters()), weight_decay=0.1),
                       **train_kwargs(nets))
