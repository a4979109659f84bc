"""GPU tests of the flag combinations INTEGRATION.md lists as supported and of the two image loops
(render_path RN:213-255, render_path_grad RN:126-210) that call the renderer."""
import math
import os

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


def kwargs(nets, **over):
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1],
              use_viewdirs=True, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False)
    kw.update(over)
    return kw


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    m = ~(torch.isnan(a) | torch.isnan(b))
    return ((a - b).abs()[m] / b.abs()[m].clamp(min=1.0)).max().item()


K24 = [[80.0, 0, 11.5], [0, 80.0, 12.5], [0, 0, 1]]


def test_perturb_with_pytest_seed_matches_oracle(nsr, wfit, nets):
    """perturb=1 with the reference's pytest hooks (RN:455-459, RH:213-222): numpy seed 0 draws for t_rand and u."""
    H = W = 24
    pose = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    ro, rd = O.get_rays(H, W, K24, pose)
    rays = O.pack_rays(ro.reshape(-1, 3), rd.reshape(-1, 3), O.YCBV_NEAR, O.YCBV_FAR)
    n = rays.shape[0]
    np.random.seed(0)
    t_rand = torch.Tensor(np.random.rand(n, 64))
    np.random.seed(0)
    u = torch.Tensor(np.random.rand(n, 128))
    with torch.no_grad():
        ref = O.render_rays(rays, wfit[0], wfit[1], 64, 128, perturb=1., t_rand=t_rand, u=u)
        got = nsr.render_rays(rays.cuda(), nets[0], None, 64, N_importance=128, network_fine=nets[1], perturb=1., pytest=True)
    for k in ('rgb_map', 'acc_map', 'rgb0', 'acc0'):
        assert rel(got[k], ref[k]) <= 1e-3, k


def test_ndc_and_staticcam_paths_are_consistent(nsr, nets):
    """ndc=True (RN:101-103, RH:178-195) and c2w_staticcam (RN:94-96) only reshape the ray batch: render() must equal
    render_rays() on rays prepared with the same formulas."""
    H = W = 16
    Kc = [[60.0, 0, 7.5], [0, 60.0, 8.5], [0, 0, 1]]
    pose = O.pose_spherical(60., 30., 1.3)[:3, :4].cuda()
    other = O.pose_spherical(75., 40., 1.2)[:3, :4].cuda()
    with torch.no_grad():
        # ndc
        a = nsr.render(H, W, Kc, chunk=512, c2w=pose, **kwargs(nets, ndc=True, near=0., far=1.))
        ro, rd = nsr.get_rays(H, W, Kc, pose)
        vd = (rd / rd.norm(dim=-1, keepdim=True)).reshape(-1, 3)
        no, nd = nsr.ndc_rays(H, W, Kc[0][0], 1., ro, rd)
        packed = torch.cat([no.reshape(-1, 3), nd.reshape(-1, 3), torch.zeros(H * W, 1, device='cuda'), torch.ones(H * W, 1, device='cuda'), vd], -1)
        b = nsr.render_rays(packed, nets[0], None, 64, N_importance=128, network_fine=nets[1])
        assert torch.equal(a[0].reshape(-1, 3), b['rgb_map'])
        # static camera: rays from `other`, view directions from `pose`
        c = nsr.render(H, W, Kc, chunk=512, c2w=pose, c2w_staticcam=other, **kwargs(nets))
        so, sd = nsr.get_rays(H, W, Kc, other)
        packed = torch.cat([so.reshape(-1, 3), sd.reshape(-1, 3), torch.full((H * W, 1), O.YCBV_NEAR, device='cuda'),
                            torch.full((H * W, 1), O.YCBV_FAR, device='cuda'), vd], -1)
        d = nsr.render_rays(packed, nets[0], None, 64, N_importance=128, network_fine=nets[1])
        assert torch.equal(c[0].reshape(-1, 3), d['rgb_map'])
    assert torch.isfinite(a[0]).all() and torch.isfinite(c[0]).all()


def test_raw_noise_perturbs_density_only_where_it_matters(nsr, nets):
    H = W = 16
    pose = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4].cuda()
    with torch.no_grad():
        clean = nsr.render(H, W, K24, chunk=512, c2w=pose, **kwargs(nets))
        torch.manual_seed(0)
        noisy = nsr.render(H, W, K24, chunk=512, c2w=pose, **kwargs(nets, raw_noise_std=1.0))
    assert noisy[0].shape == clean[0].shape and torch.isfinite(noisy[0]).all()
    d = (noisy[0] - clean[0]).abs()
    assert 0 < float(d.max()) < 0.5          # sigma noise of 1.0 moves colours a little, never wildly
    assert set(noisy[3]) == set(clean[3])


def diff_pose(phi_deg, radius=1.01):
    """pose_spherical(theta=90, phi, r) with torch ops so that d(pose)/d(phi) exists (LL:25-71 in spirit)."""
    ph = phi_deg / 180. * math.pi
    c, s = torch.cos(ph), torch.sin(ph)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    rot_phi = torch.stack([torch.stack([one, zero, zero, zero]), torch.stack([zero, c, -s, zero]),
                           torch.stack([zero, s, c, zero]), torch.stack([zero, zero, zero, one])])
    trans = torch.eye(4, device=ph.device)
    trans[2, 3] = radius
    th = torch.tensor(math.pi / 2, device=ph.device)
    rot_th = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]],
                          device=ph.device)
    flip = torch.tensor([[-1., 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], device=ph.device)
    return flip @ rot_th @ rot_phi @ trans


def test_render_path_and_render_path_grad(nsr, nets, tmp_path):
    """render_path writes the PNGs render_images expects; render_path_grad's single-pass gradient has the same MEAN
    (MAIN:191) as the reference's loop over `chunk`-ray patches (emulated here with this renderer, chunk = 100)."""
    H = W = 20
    Kc = [[66.0, 0, 9.5], [0, 66.0, 10.5], [0, 0, 1]]
    hwf = [H, W, 66.0]
    degrees = torch.tensor([0, 45, 90, 135, 180, 225, 270, 315], dtype=torch.float32, device='cuda') + 22.5
    prob = torch.tensor([0.05, 0.05, 0.05, 0.05, 0.6, 0.1, 0.05, 0.05], device='cuda', requires_grad=True)
    offsets = [0.0, 7.0]
    poses = [diff_pose((prob * degrees).sum() + o - 180.0) for o in offsets]
    g = torch.Generator().manual_seed(4)
    grad_E = [{'image_index': i, 'grad_E': torch.randn(1, 3, H, W, generator=g) * 1e-3} for i in range(2)]
    kw = kwargs(nets)
    chunk = 100
    # --- reference loop (RN:141-194), patch by patch
    ref_list = []
    for i_pose, c2w in enumerate(poses):
        ro, rd = nsr.get_rays(H, W, Kc, c2w[:3, :4])
        gi = grad_E[i_pose]['grad_E'][0].permute(1, 2, 0).reshape(-1, 3).cuda()
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        for i in range(0, H * W, chunk):
            batch_rays = torch.stack([ro[i:i + chunk], rd[i:i + chunk]], 0)
            rgb_p, _, _, _ = nsr.render(H, W, Kc, chunk=chunk, rays=batch_rays, retraw=True, **kw)
            dLdray = torch.autograd.grad(rgb_p, batch_rays, grad_outputs=gi[i:i + chunk])
            dLdpsi = torch.autograd.grad(batch_rays, prob, grad_outputs=dLdray, retain_graph=True)
            ref_list.append(dLdpsi[0].cpu())
    ref_mean = torch.mean(torch.stack(ref_list), 0)
    # --- one pass per image
    rgbs, dLdpsis = nsr.render_path_grad(prob, poses, hwf, Kc, chunk, grad_E, kw, savedir=str(tmp_path), object_id=2)
    got_mean = torch.mean(torch.stack(dLdpsis), 0)
    assert rgbs.shape == (2, H, W, 3) and len(dLdpsis) == 2
    scale = ref_mean.abs().max().item()
    assert (got_mean - ref_mean).abs().max().item() <= 1e-3 * scale, (got_mean, ref_mean)
    assert os.path.exists(tmp_path / '2' / 'withgrad' / '001.png')
    # --- render_path
    imgs, disps = nsr.render_path(prob, [p.detach() for p in poses], hwf, Kc, chunk, kw, savedir=str(tmp_path), object_id=2)
    assert imgs.shape == (2, H, W, 3) and disps.shape == (2, H, W)
    assert np.allclose(imgs, rgbs, atol=1e-5)
    from PIL import Image
    png = np.asarray(Image.open(tmp_path / '2' / '000.png'))
    assert png.shape == (H, W, 3) and np.array_equal(png, nsr.to8b(imgs[0]))


def test_forward_is_cuda_graph_capturable(nsr, nets):
    """Nothing in nsr_render_rays_forward synchronises the host or allocates: the whole 6-kernel sequence can be captured
    once and replayed (how a serving loop removes launch overhead for small ray chunks)."""
    import ctypes
    L = nsr.lib()
    H = W = 400
    pose = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    n = 4096
    rays_all = nsr.make_rays(H, W, O.YCBV_K_400, pose, O.YCBV_NEAR, O.YCBV_FAR)
    rays = rays_all[60000:60000 + n].clone()
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    new = lambda *s: torch.empty(*s, device='cuda')
    outs = [new(n, 3), new(n), new(n), new(n, 3), new(n), new(n), new(n)]
    wsb = L.nsr_render_workspace_bytes(n, 64, 128)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def launch(stream):
        rc = L.nsr_render_rays_forward(P(rays), n, P(pc), P(pf), 64, 128, 0, None, None, *[P(t) for t in outs], None, None, None,
                                       P(ws), wsb, ctypes.c_void_p(stream.cuda_stream))
        assert rc == 0, L.nsr_last_error()

    launch(torch.cuda.current_stream())
    torch.cuda.synchronize()
    eager = outs[0].clone()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            launch(side)
    rays.copy_(rays_all[90000:90000 + n])          # new input, same buffers
    graph.replay()
    torch.cuda.synchronize()
    replayed = outs[0].clone()
    launch(torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(replayed, outs[0]) and not torch.equal(replayed, eager)


def test_bilevel_psi_gradient_end_to_end(nsr, wfit, nets):
    """MAIN:134-191 on the device: psi -> softmax -> sample_pose (all K poses in one op) -> render_path_grad -> mean
    dL/dpsi, against the same chain evaluated on the CPU with the oracle renderer and autograd."""
    H = W = 20
    Kc = [[66.0, 0, 9.5], [0, 66.0, 10.5], [0, 0, 1]]
    hwf = [H, W, 66.0]
    n_k = 3
    psi = torch.tensor([0.1, -0.2, 0.3, 0.0, 0.5, -0.1, 0.2, -0.3])
    _, log = nsr.sample_pose_nograd(torch.softmax(psi / 0.25, 0), n_k, 0.5, seed=11)
    gen = torch.Generator().manual_seed(6)
    grad_E = [{'image_index': i, 'grad_E': torch.randn(1, 3, H, W, generator=gen) * 1e-3} for i in range(n_k)]
    # --- oracle chain on the CPU
    prob_c = torch.softmax(psi / 0.25, 0).requires_grad_()
    poses_c = nsr.sample_pose(prob_c, n_k, 0.5, log)
    ref = []
    for i in range(n_k):
        ro, rd = O.get_rays(H, W, Kc, poses_c[i, :3, :4])
        out = O.render(H, W, Kc, wfit[0], wfit[1], chunk=512, rays=(ro.reshape(-1, 3), rd.reshape(-1, 3)),
                       near=O.YCBV_NEAR, far=O.YCBV_FAR, N_samples=64, N_importance=128)
        gi = grad_E[i]['grad_E'][0].permute(1, 2, 0).reshape(-1, 3)
        ref.append(torch.autograd.grad(out[0], prob_c, grad_outputs=gi, retain_graph=True)[0])
    ref_mean = torch.stack(ref).mean(0)
    # --- device chain
    prob_g = torch.softmax(psi.cuda() / 0.25, 0).requires_grad_()
    poses_g = nsr.sample_pose(prob_g, n_k, 0.5, log)
    assert poses_g.is_cuda and poses_g.shape == (n_k, 4, 4)
    assert (poses_g.detach().cpu() - poses_c.detach()).abs().max().item() < 1e-5
    _, dLdpsis = nsr.render_path_grad(prob_g, poses_g, hwf, Kc, H * W, grad_E, kwargs(nets))
    got_mean = torch.stack(dLdpsis).mean(0)
    scale = ref_mean.abs().max().item()
    assert scale > 0
    assert (got_mean - ref_mean).abs().max().item() <= 1e-3 * scale, (got_mean, ref_mean)


_STAND_IN = '''
def render(*args, **kwargs):
    raise RuntimeError("the stand-in module's own render() must have been replaced")


def render_path(render_poses, hwf, K, chunk, render_kwargs):
    """shaped like the reference's image loop: `render` is looked up in THIS module's globals at call time (RN:233)"""
    H, W, focal = hwf
    frames = []
    for c2w in render_poses:
        rgb, disp, acc, extras = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)
        frames.append(rgb)
    return frames
'''


def test_install_on_a_stand_in_module_runs_on_the_gpu(nsr, nets):
    """install() (INTEGRATION.md's two-line binding) on a minimal module whose image loop resolves `render` through its own
    globals like RN:233 does: after the patch the loop renders on the GPU, with the same pixels as calling nsr.render directly
    (the CPU test tests/test_host_logic.py does the same on the unmodified reference module)."""
    import types
    mod = types.ModuleType('stand_in_run_nerf')
    exec(_STAND_IN, mod.__dict__)
    poses = torch.stack([O.pose_spherical(90., 22.5 - 180., 1.01), O.pose_spherical(90., 202.5 - 180., 1.01)]).cuda()
    with pytest.raises(RuntimeError):
        mod.render_path(poses, [24, 24, 80.], K24, 512, kwargs(nets))
    assert nsr.install(mod) is mod
    assert mod.render is nsr.render and mod.render_path.__globals__['render'] is nsr.render
    before = nsr.lib().nsr_launch_count()
    with torch.no_grad():
        frames = mod.render_path(poses, [24, 24, 80.], K24, 512, kwargs(nets))
        direct = [nsr.render(24, 24, K24, chunk=512, c2w=p[:3, :4], **kwargs(nets))[0] for p in poses]
    assert nsr.lib().nsr_launch_count() > before
    assert len(frames) == 2 and frames[0].shape == (24, 24, 3) and frames[0].is_cuda
    for a, b in zip(frames, direct):
        assert torch.equal(a, b)
    assert float(frames[0].max()) > 0.05, 'the fitted object must be visible'
