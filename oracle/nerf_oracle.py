"""CPU oracle for the NeRF per-ray render hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (fp32, CPU) restatement of the algorithm the
reference implements in

    RN = /root/reference/optimization/utils/run_nerf_noscale.py
    RH = /root/reference/optimization/utils/run_nerf_helpers.py

It exists so that the CUDA path can be checked on the GPU box, where
/root/reference does not exist.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it; the product
package (neural-sim-nerf_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against the reference ITSELF, imported unmodified in
the build container (oracle/ref_import.py) -- see tests/test_oracle_pinned.py
and the committed fixtures in tests/golden/ made by oracle/make_golden.py.

Weights are handled as a state-dict (name -> tensor) with the reference's
parameter names (RH:82-97): pts_linears.{0..7}.{weight,bias},
views_linears.0.*, feature_linear.*, alpha_linear.*, rgb_linear.*.
"""
import math

import numpy as np
import torch

N_FREQ_XYZ = 10   # multires        (MAIN:1266)  -> 3 + 3*2*10 = 63 channels
N_FREQ_DIR = 4    # multires_views  (MAIN:1268)  -> 3 + 3*2*4  = 27 channels
NET_DEPTH = 8
NET_WIDTH = 256
SKIP_AFTER = 4    # skips=[4] (RN:268)


# --------------------------------------------------------------------------
# positional encoding, RH:18-66
# --------------------------------------------------------------------------
def embed(x, n_freq):
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].

    RH:23-48: include_input=True, log_sampling=True so the bands are exact
    powers of two (2**linspace(0, L-1, L)), periodic_fns=[sin, cos], each block
    3 wide.  x: [..., 3] -> [..., 3 + 6 L].
    """
    parts = [x]
    for k in range(n_freq):
        xf = x * float(2.0 ** k)
        parts.append(torch.sin(xf))
        parts.append(torch.cos(xf))
    return torch.cat(parts, dim=-1)


# --------------------------------------------------------------------------
# the MLP, RH:99-122  (D=8, W=256, skips=[4]; use_viewdirs=True, or False when the state-dict holds `output_linear`)
# --------------------------------------------------------------------------
def mlp_forward(x, sd):
    """x: [P, 63+27] -> [P, 4] = (rgb_raw[3], sigma_raw[1]).  RH:99-122.
    use_viewdirs=False (a state-dict with `output_linear.*`, RH:95-96): x [P, 63] -> [P, output_ch], RH:119-120; the
    compositor reads columns 0..3 only (RN:361-364), a fifth column (output_ch = 5 when N_importance > 0, RN:267) is carried."""
    lin = torch.nn.functional.linear
    enc_xyz = x[..., :63]
    enc_dir = x[..., 63:]
    h = enc_xyz
    for i in range(NET_DEPTH):
        h = torch.relu(lin(h, sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias']))
        if i == SKIP_AFTER:
            h = torch.cat([enc_xyz, h], dim=-1)          # RH:105-106 (input first)
    if 'output_linear.weight' in sd:
        return lin(h, sd['output_linear.weight'], sd['output_linear.bias'])     # RH:119-120
    sigma = lin(h, sd['alpha_linear.weight'], sd['alpha_linear.bias'])          # RH:109
    feat = lin(h, sd['feature_linear.weight'], sd['feature_linear.bias'])       # RH:110 (no activation)
    h = torch.cat([feat, enc_dir], dim=-1)                                      # RH:111
    h = torch.relu(lin(h, sd['views_linears.0.weight'], sd['views_linears.0.bias']))  # RH:113-115
    rgb = lin(h, sd['rgb_linear.weight'], sd['rgb_linear.bias'])                # RH:117
    return torch.cat([rgb, sigma], dim=-1)                                      # RH:118


def run_network(pts, viewdirs, sd, netchunk=1024 * 64):
    """RN:26-40.  pts [n,S,3], viewdirs [n,3] (None with use_viewdirs=False, RN:32) -> raw [n,S,4 (or output_ch)]."""
    flat = pts.reshape(-1, 3)
    emb = embed(flat, N_FREQ_XYZ)
    if viewdirs is not None:                                                  # RN:32
        dirs = viewdirs[:, None, :].expand(pts.shape).reshape(-1, 3)          # RN:33-34
        emb = torch.cat([emb, embed(dirs, N_FREQ_DIR)], dim=-1)               # RN:35-36
    outs = [mlp_forward(emb[i:i + netchunk], sd) for i in range(0, emb.shape[0], netchunk)]  # RN:14-23
    outs = torch.cat(outs, 0)
    return outs.reshape(*pts.shape[:-1], outs.shape[-1])


# --------------------------------------------------------------------------
# alpha compositing, RN:343-387
# --------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, noise=None, white_bkgd=False):
    """RN:343-387.  `noise` (optional, same shape as raw[...,3]) stands in for
    randn*raw_noise_std (RN:365-366).  Returns rgb_map, disp_map, acc_map,
    weights, depth_map."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)   # RN:358-359
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)                # RN:361
    rgb = torch.sigmoid(raw[..., :3])                                       # RN:363
    sig = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-torch.relu(sig) * dists)                        # RN:356,374
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]  # RN:376
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)                       # RN:378
    depth_map = torch.sum(weights * z_vals, -1)                             # RN:380
    acc_map = torch.sum(weights, -1)                                        # RN:382
    # RN:381 -- torch.max propagates the 0/0 NaN when acc == 0
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])                       # RN:384-385
    return rgb_map, disp_map, acc_map, weights, depth_map


# --------------------------------------------------------------------------
# inverse-CDF resampling, RH:199-243
# --------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples, det=True, u=None):
    """RH:199-243.  bins [n,B], weights [n,B-1] -> samples [n,n_samples].
    `u` overrides the uniform draws (RH:211) for seeded tests."""
    w = weights + 1e-5                                                       # RH:201
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)               # RH:204
    if u is None:
        if det:
            u = torch.linspace(0., 1., steps=n_samples).expand(list(cdf.shape[:-1]) + [n_samples])  # RH:208-209
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                            # RH:227
    below = torch.clamp(inds - 1, min=0)                                     # RH:228
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)                         # RH:229
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)         # RH:239
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b)                                       # RH:241


# --------------------------------------------------------------------------
# render_rays, RN:390-501
# --------------------------------------------------------------------------
def render_rays(ray_batch, sd_coarse, sd_fine, N_samples=64, N_importance=128,
                retraw=False, lindisp=False, perturb=0., white_bkgd=False,
                netchunk=1024 * 64, t_rand=None, u=None, return_internals=False, z_fine=None):
    """ray_batch [n,11] = o(3) d(3) near far viewdir(3)  (RN:433-437).
    perturb>0 needs t_rand [n,N_samples] and u [n,N_importance] given explicitly
    (the reference draws torch.rand there, RN:453 / RH:211).  `z_fine` [n, N_samples+N_importance] replaces the merged
    depths of RN:477 (tests use it to compare gradients on identical sample positions: sample_pdf's `denom < 1e-5`
    branch, RH:239, is discontinuous in the last bits of the coarse weights)."""
    n = ray_batch.shape[0]
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None      # RN:437
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    t_vals = torch.linspace(0., 1., steps=N_samples)                          # RN:439
    if not lindisp:
        z_vals = near * (1. - t_vals) + far * t_vals                          # RN:441
    else:
        z_vals = 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)         # RN:443
    z_vals = z_vals.expand([n, N_samples])
    if perturb > 0.:                                                          # RN:447-461
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        z_vals = lower + (upper - lower) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]  # RN:463
    raw = run_network(pts, viewdirs, sd_coarse, netchunk)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, None, white_bkgd)
    internals = {'raw0': raw, 'z0': z_vals, 'weights0': weights}
    ret = {}
    if N_importance > 0:
        rgb0, disp0, acc0 = rgb_map, disp_map, acc_map
        z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])                     # RN:473
        z_samples = sample_pdf(z_mid, weights[..., 1:-1], N_importance,
                               det=(perturb == 0.), u=u).detach()             # RN:474-475
        z_vals, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)        # RN:477
        if z_fine is not None:
            z_vals = z_fine
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
        raw = run_network(pts, viewdirs, sd_fine if sd_fine is not None else sd_coarse, netchunk)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, None, white_bkgd)
        internals.update({'z_samples': z_samples, 'z1': z_vals, 'raw1': raw, 'weights1': weights})
    ret.update({'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map})
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'], ret['disp0'], ret['acc0'] = rgb0, disp0, acc0
        ret['z_std'] = torch.std(z_samples, dim=-1, unbiased=False)           # RN:495
    if return_internals:
        ret['_internals'] = internals
    return ret


# --------------------------------------------------------------------------
# get_rays (RH:156-165) and render (RN:58-123)
# --------------------------------------------------------------------------
def get_rays(H, W, K, c2w):
    """RH:156-165: pixel (i=x, j=y) -> dir [(i-cx)/fx, -(j-cy)/fy, -1] rotated by c2w[:3,:3]."""
    K = torch.as_tensor(K, dtype=torch.float32)
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    j, i = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing='ij')
    dirs = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def pack_rays(rays_o, rays_d, near, far, use_viewdirs=True):
    """RN:91-112 with ndc=False -> [N,11] fp32 ([N,8] with use_viewdirs=False, RN:111)."""
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    viewdirs = viewdirs.reshape(-1, 3).float()
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    nr = near * torch.ones_like(rays_d[..., :1])
    fr = far * torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, nr, fr] + ([viewdirs] if use_viewdirs else []), -1)


def render(H, W, K, sd_coarse, sd_fine, chunk=512, rays=None, c2w=None, near=0., far=1., use_viewdirs=True,
           **kw):
    """RN:58-123 (ndc=False): returns [rgb_map, disp_map, acc_map, extras]."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    sh = rays_d.shape
    packed = pack_rays(rays_o, rays_d, near, far, use_viewdirs)
    chunks = {}
    for i in range(0, packed.shape[0], chunk):                                # RN:43-55
        r = render_rays(packed[i:i + chunk], sd_coarse, sd_fine, **kw)
        for k, v in r.items():
            chunks.setdefault(k, []).append(v)
    out = {k: torch.cat(v, 0) for k, v in chunks.items()}
    out = {k: v.reshape(list(sh[:-1]) + list(v.shape[1:])) for k, v in out.items()}
    keys = ['rgb_map', 'disp_map', 'acc_map']
    return [out[k] for k in keys] + [{k: v for k, v in out.items() if k not in keys}]


def to8b(x):
    """RH:14."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def rays_grad_to_c2w(H, W, K, c2w, d_packed, near=0., far=1.):
    """What autograd computes for RN:179-181 restricted to the pose: the pull-back of a cotangent on the packed rays
    [H*W,11] (RN:106-112) through RN:97 and get_rays (RH:156-165) to c2w [3,4]."""
    c = torch.as_tensor(c2w, dtype=torch.float32)[:3, :4].clone().requires_grad_(True)
    ro, rd = get_rays(H, W, K, c)
    packed = pack_rays(ro, rd, near, far)
    g, = torch.autograd.grad(packed, c, grad_outputs=torch.as_tensor(d_packed, dtype=torch.float32))
    return g


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123 library's
    philox4x32 with 10 rounds) on numpy uint32 arrays: counter [...,4], key [...,2] -> [...,4].  Not part of the reference
    (which draws with torch.rand / torch.randn, RN:455, RH:211, RN:366): it is the checker of the device generator that
    replaces those draws in nsr_train_step, pinned by Random123's published known-answer vectors (tests/test_host_logic.py)."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return np.stack(c, -1).astype(np.uint32)


def philox_uniform(seed, stream, count):
    """The device generator's addressing (include/nsr_b200.h nsr_random_uniform): element 4q+j = component j of
    philox(counter = (q lo, q hi, stream, 0), key = (seed lo, seed hi)), mapped to [0,1) as (x >> 8) * 2^-24."""
    q = np.arange((count + 3) // 4, dtype=np.uint64)
    ctr = np.stack([q & np.uint64(0xFFFFFFFF), q >> np.uint64(32), np.full_like(q, stream), np.zeros_like(q)], -1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64), (len(q), 2))
    r = philox4x32_10(ctr, key).reshape(-1)[:count]
    return ((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


# --------------------------------------------------------------------------
# helpers shared by tests / bench (not part of the reference)
# --------------------------------------------------------------------------
PARAM_SHAPES = ([(f'pts_linears.{i}', (256, 63 if i == 0 else (319 if i == 5 else 256))) for i in range(8)]
                + [('views_linears.0', (128, 283)), ('feature_linear', (256, 256)),
                   ('alpha_linear', (1, 256)), ('rgb_linear', (3, 128))])


def random_state_dict(seed, scale=1.0):
    """nn.Linear-style init (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) from a numpy
    RandomState so it is reproducible across torch versions.  `scale` widens the
    weights to push sigma / rgb away from zero."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, (o, i) in PARAM_SHAPES:
        b = 1.0 / math.sqrt(i)
        sd[name + '.weight'] = torch.from_numpy(rs.uniform(-b, b, size=(o, i)).astype(np.float32) * scale)
        sd[name + '.bias'] = torch.from_numpy(rs.uniform(-b, b, size=(o,)).astype(np.float32))
    return sd


def viewless_state_dict(sd, output_ch=5):
    """A use_viewdirs=False network (RH:95-96: `output_linear` [output_ch, 256] instead of the feature / alpha / rgb heads; RN:263
    leaves input_ch_views = 0, so views_linears.0 is [128, 256] and unused) made from a view-dependent state-dict: same trunk, sigma row
    = alpha_linear, colour rows = the view-dependent head with its ReLU and its view columns dropped (some fixed linear map of h).
    Test fixture only: the density field stays the one the trunk was fitted to."""
    out = {k: v.clone() for k, v in sd.items() if k.startswith('pts_linears.')}
    wv = sd['views_linears.0.weight'][:, :NET_WIDTH]
    out['views_linears.0.weight'] = wv.clone()
    out['views_linears.0.bias'] = sd['views_linears.0.bias'].clone()
    w = torch.zeros(output_ch, NET_WIDTH)
    b = torch.zeros(output_ch)
    w[0:3] = sd['rgb_linear.weight'] @ wv @ sd['feature_linear.weight']
    b[0:3] = sd['rgb_linear.bias'] + sd['rgb_linear.weight'] @ (wv @ sd['feature_linear.bias'] + sd['views_linears.0.bias'])
    w[3] = sd['alpha_linear.weight'][0]
    b[3] = sd['alpha_linear.bias'][0]
    if output_ch > 4:
        w[4:] = 0.01
    out['output_linear.weight'], out['output_linear.bias'] = w, b
    return out


# camera used by BASELINE configs (logs/nerfdata/nerf_traindata_info.json; LL:185-198)
YCBV_K_400 = [[1333.3333740234375, 0.0, 195.43], [0.0, 1334.22, 200.63], [0.0, 0.0, 1.0]]
YCBV_NEAR = 0.8103964843749999 - 0.5
YCBV_FAR = 1.4297681884765627 + 0.5


def pose_spherical(theta_deg, phi_deg, radius):
    """LL:89-94 pose_spherical_nograd: camera on a sphere looking at the origin."""
    th, ph = theta_deg / 180. * np.pi, phi_deg / 180. * np.pi
    t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=np.float32)
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]], dtype=np.float32)
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], dtype=np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    return torch.from_numpy(flip @ rt @ rp @ t)
