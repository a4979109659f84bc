"""Host-side mirror of the reference renderer surface, backed by libnsr_b200 (CUDA, sm_100a).

Same names, argument meaning and return structure as the reference's
  RN = optimization/utils/run_nerf_noscale.py   render (RN:58), batchify_rays (RN:43), render_rays (RN:390),
                                                run_network (RN:26), batchify (RN:14), raw2outputs (RN:343)
  RH = optimization/utils/run_nerf_helpers.py   sample_pdf (RH:199), get_rays (RH:156), ndc_rays (RH:178),
                                                Embedder/get_embedder (RH:18-66), NeRF (RH:70-122)
so `neural_sim_main.py` keeps calling render()/render_rays()/run_network() unchanged.

PyTorch is used for device memory, streams and the nn.Module that owns the weights; every
numerical stage runs in the CUDA library.  There is no CPU / eager fallback: unsupported
configurations raise NotImplementedError.
"""
import ctypes
import math
import os
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (RN:6 star-exports it: `from utils.run_nerf_noscale import *` users may rely on `F`)

from . import _lib
from ._lib import FLAG_DENSE, FLAG_EMBEDDED_INPUT, FLAG_FAST_FP16, FLAG_LINDISP, FLAG_MIXED_F8, FLAG_PTS_INPUT, FLAG_WHITE_BKGD, check, lib, ptr

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

# render() may merge the caller's ray chunks up to this many rays per launch: the reference's
# `chunk` only bounds memory ("Does not affect final results", RN:67-68) and 512-ray chunks
# (CFG:25) would leave most of a B200 idle.
MIN_RAYS_PER_LAUNCH = int(os.environ.get('NSR_MIN_CHUNK', 1 << 16))

# MLP arithmetic of the forward pass (DESIGN.md "precision"):
#   'fp16x3'  error-compensated fp16 hi/lo split (3 MMAs per product): ~1e-5 from the fp32 reference.  The render entry points
#             apply it to the ACTIVE sample points only -- a cheap first tier certifies the empty ones (sigma <= 0: weight exactly
#             0), see include/nsr_b200.h "two-tier evaluation"; the maps are bit-identical to evaluating every point;
#   'fp16x3-dense'  the same arithmetic on every point (NSR_FLAG_DENSE);
#   'mixed'   the split for the first three layers, fp16 + e4m3 residual products (NSR_FLAG_MIXED_F8) for the rest:
#             2.25 tensor passes per product, ~2e-4 from the reference -- inside the 1e-3 bar;
#   'fp16'    single fp16 MMA per product (NSR_FLAG_FAST_FP16): misses the bar on silhouette rays, opt-in only.
# Passes that carry gradient always run 'fp16x3' (the backward kernel recomputes activations in that arithmetic).
PRECISION = os.environ.get('NSR_PRECISION', 'fp16x3')
_PREC_FLAGS = {'fp16x3': 0, 'fp16x3-dense': FLAG_DENSE, 'mixed': FLAG_MIXED_F8, 'fp16': FLAG_FAST_FP16}


def set_precision(mode):
    global PRECISION
    if mode not in _PREC_FLAGS:
        raise ValueError(f"precision must be one of {sorted(_PREC_FLAGS)}")
    PRECISION = mode


def _prec_flag():
    if PRECISION not in _PREC_FLAGS:
        raise ValueError(f"NSR_PRECISION={PRECISION!r}: must be one of {sorted(_PREC_FLAGS)}")
    return _PREC_FLAGS[PRECISION]


img2mse = lambda x, y: torch.mean((x - y) ** 2)                                  # RH:12
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.Tensor([10.]).to(x.device))  # RH:13
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)                       # RH:14


# ----------------------------------------------------------------------------- modules (RH:18-122)
class Embedder:
    """RH:18-48.  Kept for API compatibility; the CUDA kernel has multires=10 / 4 built in."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs['input_dims']
        n = kwargs['num_freqs']
        if kwargs['log_sampling']:
            self.freq_bands = 2. ** torch.linspace(0., kwargs['max_freq_log2'], steps=n)
        else:
            self.freq_bands = torch.linspace(2. ** 0., 2. ** kwargs['max_freq_log2'], steps=n)
        self.out_dim = (d if kwargs['include_input'] else 0) + d * n * len(kwargs['periodic_fns'])

    def embed(self, inputs):
        out = [inputs] if self.kwargs['include_input'] else []
        for f in self.freq_bands:
            for fn in self.kwargs['periodic_fns']:
                out.append(fn(inputs * f))
        return torch.cat(out, -1)


def get_embedder(multires, i=0):
    """RH:51-66."""
    if i == -1:
        return nn.Identity(), 3
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return (lambda x, eo=eo: eo.embed(x)), eo.out_dim


class NeRF(nn.Module):
    """Parameter container with the reference's layout and names (RH:70-97), so reference
    checkpoints (`network_fn_state_dict` / `network_fine_state_dict`, RN:306-314) load unchanged."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=(4,), use_viewdirs=True):
        super().__init__()
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs = list(skips), use_viewdirs
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W + input_ch, W) if i in self.skips else nn.Linear(W, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)

    def forward(self, x):
        """RH:99-122: x [..., input_ch + input_ch_views] already embedded -> [..., 4] = (rgb, sigma) raw, on the tensor-core kernel
        (nsr_mlp_forward with NSR_FLAG_EMBEDDED_INPUT: the encoder warps copy the 90 channels instead of computing them).
        Forward only: the differentiable routes are render() / render_rays() / train_step(), whose backward kernels start from
        points, not from embeddings -- asking this call for a graph raises instead of silently returning a constant."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError('NeRF.forward(x) is forward-only here (wrap it in torch.no_grad()); gradients flow through '
                                      'render() / render_rays() / train_step()')
        if _is_viewless(self):
            # RH:119-120; the kernel returns the four columns the compositor reads (RN:361-364)
            if self.output_linear.out_features != 4:
                raise NotImplementedError('NeRF.forward(x) with use_viewdirs=False returns four columns; output_ch = '
                                          f'{self.output_linear.out_features} (RN:267) has a fifth that no caller reads and the kernel does not compute')
            if x.shape[-1] != 63:
                raise NotImplementedError(f'embedded input must have 63 channels (input_ch_views = 0, RN:263), got {x.shape[-1]}')
            x90 = torch.zeros(*x.shape[:-1], 90, dtype=torch.float32, device=x.device)
            x90[..., :63] = x
            x = x90
        if x.shape[-1] != 90:
            raise NotImplementedError(f'embedded input must have 63 + 27 = 90 channels, got {x.shape[-1]}')
        xf = _f32c(x.reshape(-1, 90), 'x')
        raw = torch.empty(xf.shape[0], 4, dtype=torch.float32, device=xf.device)
        check(lib().nsr_mlp_forward(None, ptr(xf), xf.shape[0], 1, ptr(packed_weights(self)), FLAG_EMBEDDED_INPUT | _prec_flag(), ptr(raw), _stream()),
              'nsr_mlp_forward')
        return raw.reshape(*x.shape[:-1], 4)


# ----------------------------------------------------------------------------- packed-weight cache
_EXPECTED_SHAPES = [(256, 63)] + [(256, 256)] * 4 + [(256, 319)] + [(256, 256)] * 2 + [(128, 283), (256, 256), (1, 256), (3, 128)]
_pack_cache = weakref.WeakKeyDictionary()


def _is_viewless(net):
    """A NeRF(use_viewdirs=False) module: `output_linear` instead of the feature / alpha / rgb heads (RH:95-96)."""
    return not getattr(net, 'use_viewdirs', True) and hasattr(net, 'output_linear')


def _viewless_operands(net):
    """use_viewdirs=False (RH:119-120: outputs = output_linear(h); columns 0..2 colour, 3 sigma, RN:361-364) on the kernel built for the
    view-dependent head.  With A, a = rows 0..2 of output_linear, the twelve operand tensors are
        feature_linear = I, 0          views_linears.0 = [A; -A; 0] on the 256 feature columns (0 on the 27 view columns), [a; -a; 0]
        rgb_linear = [I3, -I3, 0], 0   alpha_linear = row 3 of output_linear
    so that rgb = relu(A h + a) - relu(-(A h + a)) = A h + a.  (I and the +-1 are exact in fp16; the feature vector is h to 2^-22.)
    Returns (parameters the blob depends on, weights, biases)."""
    pts = list(net.pts_linears)
    out = net.output_linear
    if len(pts) != 8:
        raise NotImplementedError(f'netdepth {len(pts)} != 8 is not supported by the sm_100a kernel')
    for l, shp in zip(pts, _EXPECTED_SHAPES[:8]):
        if tuple(l.weight.shape) != shp:
            raise NotImplementedError(f'layer shape {tuple(l.weight.shape)} != {shp}: only D=8, W=256, multires=10 is built')
    if out.weight.shape[1] != 256 or out.weight.shape[0] < 4:
        raise NotImplementedError(f'output_linear {tuple(out.weight.shape)}: expected [4 or 5, 256] (RN:267)')
    params = [l.weight for l in pts] + [out.weight] + [l.bias for l in pts] + [out.bias]
    dev = out.weight.device
    A, a = out.weight.detach()[:3], out.bias.detach()[:3]
    wv = torch.zeros(128, 283, dtype=torch.float32, device=dev)
    bv = torch.zeros(128, dtype=torch.float32, device=dev)
    wv[0:3, :256], wv[3:6, :256] = A, -A
    bv[0:3], bv[3:6] = a, -a
    wr = torch.zeros(3, 128, dtype=torch.float32, device=dev)
    for c in range(3):
        wr[c, c], wr[c, 3 + c] = 1.0, -1.0
    ws = [l.weight.detach().contiguous() for l in pts] + [wv, torch.eye(256, dtype=torch.float32, device=dev),
                                                          out.weight.detach()[3:4].contiguous(), wr]
    bs = [l.bias.detach().contiguous() for l in pts] + [bv, torch.zeros(256, dtype=torch.float32, device=dev),
                                                        out.bias.detach()[3:4].contiguous(), torch.zeros(3, dtype=torch.float32, device=dev)]
    return params, ws, bs


def _net_tensors(net):
    try:
        layers = list(net.pts_linears) + [net.views_linears[0], net.feature_linear, net.alpha_linear, net.rgb_linear]
    except AttributeError as e:
        raise NotImplementedError('network must be a NeRF(D=8, W=256, skips=[4], use_viewdirs=True) module (RH:70-97)') from e
    if len(layers) != 12:
        raise NotImplementedError(f'netdepth {len(net.pts_linears)} != 8 is not supported by the sm_100a kernel')
    for l, shp in zip(layers, _EXPECTED_SHAPES):
        if tuple(l.weight.shape) != shp:
            raise NotImplementedError(f'layer shape {tuple(l.weight.shape)} != {shp}: only D=8, W=256, multires=10/4 is built')
    return layers


def packed_weights(net):
    """Device blob of `net` in the kernel's operand layout; re-packed (on the GPU) whenever a
    parameter's storage or version counter changes."""
    viewless = _is_viewless(net)
    if viewless:
        layers = list(net.pts_linears) + [net.output_linear]
    else:
        layers = _net_tensors(net)
    params = [l.weight for l in layers] + [l.bias for l in layers]
    key = tuple((p.data_ptr(), p._version, p.device.index) for p in params)
    hit = _pack_cache.get(net)
    if hit is not None and hit[0] == key:
        return hit[1]
    for p in params:
        if not p.is_cuda or p.dtype != torch.float32:
            raise _lib.NsrError('network parameters must be fp32 CUDA tensors (no CPU fallback)')
    if viewless:
        _, ws, bs = _viewless_operands(net)
    else:
        ws = [l.weight.detach().contiguous() for l in layers]
        bs = [l.bias.detach().contiguous() for l in layers]
    L = lib()
    blob = torch.empty(L.nsr_packed_net_bytes(), dtype=torch.uint8, device=params[0].device)
    wp = (ctypes.c_void_p * 12)(*[w.data_ptr() for w in ws])
    bp = (ctypes.c_void_p * 12)(*[b.data_ptr() for b in bs])
    check(L.nsr_pack_net(wp, bp, ptr(blob), _stream()), 'nsr_pack_net')
    _pack_cache[net] = (key, blob)
    return blob


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, name):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if not t.is_cuda:
        raise _lib.NsrError(f'{name} must be a CUDA tensor (no CPU fallback)')
    return t.detach().to(torch.float32).contiguous()


def _no_grad_inputs(*ts):
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in ts):
        raise NotImplementedError('only render()/render_rays() carry a backward (dL/d rays, SURVEY.md a-11); '
                                  'call this stage under torch.no_grad() or detach its inputs')


# ----------------------------------------------------------------------------- RN:14-40
def batchify(fn, chunk):
    """RN:14-23 (kept for API compatibility)."""
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)


def run_network(inputs, viewdirs, fn, embed_fn=None, embeddirs_fn=None, netchunk=1024 * 64):
    """RN:26-40: inputs [n,S,3], viewdirs [n,3], fn = NeRF module -> raw [n,S,4].
    embed_fn / embeddirs_fn / netchunk are accepted and ignored: encoding (multires 10 / 4) and
    chunking live inside the kernel."""
    if viewdirs is None and not _is_viewless(fn):
        raise ValueError('viewdirs=None needs a NeRF(use_viewdirs=False) network (RN:32, RH:119-120)')
    _no_grad_inputs(inputs, *([] if viewdirs is None else [viewdirs]))
    sh = inputs.shape
    pts = _f32c(inputs, 'inputs').reshape(-1, sh[-2], 3) if inputs.dim() >= 3 else _f32c(inputs, 'inputs').reshape(-1, 1, 3)
    n, S = pts.shape[0], pts.shape[1]
    rays = torch.zeros(n, 11, dtype=torch.float32, device=pts.device)
    if viewdirs is not None:
        vd = _f32c(viewdirs, 'viewdirs').reshape(-1, 3)
        if vd.shape[0] != n:
            raise ValueError(f'viewdirs {tuple(viewdirs.shape)} does not match inputs {tuple(inputs.shape)}')
        rays[:, 8:11] = vd
    raw = torch.empty(n, S, 4, dtype=torch.float32, device=pts.device)
    check(lib().nsr_mlp_forward(ptr(rays), ptr(pts), n, S, ptr(packed_weights(fn)), FLAG_PTS_INPUT | _prec_flag(), ptr(raw), _stream()),
          'nsr_mlp_forward')
    return raw.reshape(list(sh[:-1]) + [4])


# ----------------------------------------------------------------------------- RN:343-387
def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """RN:343-387 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    _no_grad_inputs(raw, z_vals, rays_d)
    raw = _f32c(raw, 'raw')
    if raw_noise_std > 0.:
        raw = raw.clone()
        raw[..., 3] += torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std  # RN:365-366
    z = _f32c(z_vals, 'z_vals')
    d = _f32c(rays_d, 'rays_d')
    n, S = z.shape
    dev = raw.device
    rgb = torch.empty(n, 3, device=dev)
    disp, acc, depth = (torch.empty(n, device=dev) for _ in range(3))
    w = torch.empty(n, S, device=dev)
    flags = FLAG_WHITE_BKGD if white_bkgd else 0
    check(lib().nsr_raw2outputs(ptr(raw), ptr(z), ptr(d), 3, n, S, flags, ptr(rgb), ptr(disp), ptr(acc), ptr(w), ptr(depth),
                                _stream()), 'nsr_raw2outputs')
    return rgb, disp, acc, w, depth


# ----------------------------------------------------------------------------- RH:199-243
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """RH:199-243: bins [n,B], weights [n,B-1] -> samples [n,N_samples]."""
    _no_grad_inputs(bins, weights)
    b = _f32c(bins, 'bins')
    w = _f32c(weights, 'weights')
    lead = b.shape[:-1]
    b2, w2 = b.reshape(-1, b.shape[-1]), w.reshape(-1, w.shape[-1])
    n, B = b2.shape
    if w2.shape != (n, B - 1):
        raise ValueError(f'weights {tuple(weights.shape)} must be bins {tuple(bins.shape)} minus one column')
    u = None
    if pytest:  # RH:213-222
        np.random.seed(0)
        if not det:
            u = torch.Tensor(np.random.rand(n, N_samples)).to(b.device)
    elif not det:
        u = torch.rand(n, N_samples, device=b.device)                              # RH:211
    out = torch.empty(n, N_samples, device=b.device)
    check(lib().nsr_sample_pdf(ptr(b2), ptr(w2), n, B, N_samples, ptr(u), ptr(out), _stream()), 'nsr_sample_pdf')
    return out.reshape(list(lead) + [N_samples])


# ----------------------------------------------------------------------------- RN:390-501
def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False):
    """Volumetric rendering of a ray batch, RN:390-501.  ray_batch [n,11] = o d near far viewdir.
    `network_query_fn` is accepted for signature compatibility; the embedders it closes over
    (RN:281-284) are the fixed multires=10 / 4 ones compiled into the kernel."""
    nets_used = [network_fn] + ([network_fine] if (network_fine is not None and int(N_importance) > 0) else [])
    if ray_batch.shape[-1] <= 8:
        # use_viewdirs=False (RN:111, RN:437): eight columns.  The kernels take eleven; the view columns meet zero weights.
        if not all(_is_viewless(m) for m in nets_used):
            raise ValueError('an 8-column ray batch (use_viewdirs=False, RN:111) needs NeRF(use_viewdirs=False) networks')
        ray_batch = torch.cat([ray_batch, torch.zeros_like(ray_batch[:, :3])], -1)
    if len({_is_viewless(m) for m in nets_used}) > 1:
        raise ValueError('network_fn and network_fine must both be use_viewdirs=True or both use_viewdirs=False (RN:268-278 builds them alike)')
    if torch.is_grad_enabled() and any(_is_viewless(m) and any(p.requires_grad for p in m.parameters()) for m in nets_used):
        raise NotImplementedError('parameter gradients of use_viewdirs=False networks are not built: freeze them '
                                  '(requires_grad_(False)) for the pose path, or train with use_viewdirs=True (CFG:8)')
    params = _params_of(network_fn) + (_params_of(network_fine) if (network_fine is not None and int(N_importance) > 0) else [])
    needs_grad = torch.is_grad_enabled() and (ray_batch.requires_grad or any(p.requires_grad for p in params))
    rays = _f32c(ray_batch, 'ray_batch')
    n = rays.shape[0]
    dev = rays.device
    S, Ni = int(N_samples), int(N_importance)
    flags = (FLAG_LINDISP if lindisp else 0) | (FLAG_WHITE_BKGD if white_bkgd else 0) | _prec_flag()
    pc = packed_weights(network_fn)
    pf = packed_weights(network_fine) if (network_fine is not None and Ni > 0) else None
    t_rand = u = None
    if perturb > 0.:                                                               # RN:447-461, RH:211
        if pytest:
            np.random.seed(0)
            t_rand = torch.Tensor(np.random.rand(n, S)).to(dev)
            np.random.seed(0)
            u = torch.Tensor(np.random.rand(n, Ni)).to(dev) if Ni > 0 else None
        else:
            t_rand = torch.rand(n, S, device=dev)
            u = torch.rand(n, Ni, device=dev) if Ni > 0 else None
    if raw_noise_std > 0.:
        if needs_grad:
            raise NotImplementedError('backward with raw_noise_std > 0 is not built')
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ret = {'rgb_map': new(n, 3), 'disp_map': new(n), 'acc_map': new(n)}
        if Ni > 0:
            ret.update({'rgb0': new(n, 3), 'disp0': new(n), 'acc0': new(n), 'z_std': new(n)})
        # noise is injected between the MLP and the compositor (RN:365-374): run the stages one by one
        return _render_rays_staged(rays, pc, pf, S, Ni, flags, t_rand, u, retraw, raw_noise_std, white_bkgd, ret)
    cfg = dict(pc=pc, pf=pf, S=S, Ni=Ni, flags=flags, t_rand=t_rand, u=u, retraw=retraw)
    if needs_grad:
        if flags & FLAG_FAST_FP16:
            raise NotImplementedError("backward is not built for NSR_PRECISION='fp16'")
        cfg['flags'] = flags & ~FLAG_MIXED_F8      # gradient-carrying passes: same arithmetic as the backward kernel's recompute
        outs = _RenderRaysFn.apply(ray_batch, cfg, *params)
    else:
        outs = _forward_impl(rays, cfg, keep_for_backward=False)[0]
    keys = ['rgb_map', 'disp_map', 'acc_map'] + (['rgb0', 'disp0', 'acc0', 'z_std'] if Ni > 0 else []) + (['raw'] if retraw else [])
    return dict(zip(keys, outs))


# Pose-gradient passes (no parameter gradient wanted) keep one bit per ReLU of the last network pass -- 272 B per sample point --
# so that the backward kernel skips its forward recompute (DESIGN.md "backward").  NSR_SAVE_RELU_MASK=0 turns it off; it is
# also skipped when the bits would not fit comfortably in the free device memory (the recompute path needs no extra memory).
SAVE_RELU_MASK = os.environ.get('NSR_SAVE_RELU_MASK', '1') != '0'


_fit_cache = {}          # device index -> (monotonic time of the query, bytes that fit)


def _mask_fits(n_bytes, dev):
    """Is there comfortably room for n_bytes more?  cudaMemGetInfo costs milliseconds, so one answer serves for two seconds
    (the allocations themselves are guarded: an out-of-memory error falls back to the recompute route)."""
    import time
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    now = time.monotonic()
    hit = _fit_cache.get(idx)
    if hit is None or now - hit[0] > 2.0:
        free, _ = torch.cuda.mem_get_info(idx)
        hit = (now, free // 2)
        _fit_cache[idx] = hit
    return n_bytes <= hit[1]


def _forward_impl(rays, cfg, keep_for_backward, save_mask=False, save_dump=False):
    """One call of nsr_render_rays_forward(_ex).  Returns (outputs tuple, saved) with saved = (z_vals [n,T], raw [n,T,4],
    z0 [n,S], raw0 [n,S,4], relu_mask, dump, active_set) of the last and (when N_importance > 0) the coarse pass, or Nones.
    save_dump (training batches): the last pass also writes its activations for the weight-gradient GEMMs (~5 KB per point).
    A caller-visible raw (retraw=True) forces the dense evaluation: the two-tier route leaves (0,0,0,sigma~) at empty points."""
    L = lib()
    n, dev = rays.shape[0], rays.device
    S, Ni = cfg['S'], cfg['Ni']
    T = S + Ni
    new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    rgb, disp, acc = new(n, 3), new(n), new(n)
    rgb0 = disp0 = acc0 = zstd = None
    if Ni > 0:
        rgb0, disp0, acc0, zstd = new(n, 3), new(n), new(n), new(n)
    raw = new(n, T, 4) if (cfg['retraw'] or keep_for_backward) else None
    zv = new(n, T) if keep_for_backward else None
    ws_bytes = L.nsr_render_workspace_bytes(n, S, Ni)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    mask = dump = aset = None
    flags = cfg['flags'] | (FLAG_DENSE if cfg['retraw'] else 0)
    if keep_for_backward and not save_dump and n > 0 and not (flags & (FLAG_FAST_FP16 | FLAG_MIXED_F8 | FLAG_DENSE)):
        # pose path: the backward pass only visits the active points of the last pass (dL/draw == 0 exactly elsewhere)
        aset = torch.empty(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device=dev)
    if (save_mask or save_dump) and keep_for_backward and SAVE_RELU_MASK and n > 0 and not (cfg['flags'] & (FLAG_FAST_FP16 | FLAG_MIXED_F8)):
        mb = L.nsr_relu_mask_bytes(n, T)
        db = L.nsr_mlp_dump_bytes(n, T) if save_dump else 0
        if _mask_fits(mb + db, dev):
            try:
                mask = torch.empty(mb, dtype=torch.uint8, device=dev)
                if save_dump:
                    dump = torch.empty(db, dtype=torch.uint8, device=dev)
            except torch.cuda.OutOfMemoryError:
                mask = dump = None               # the recompute route needs neither
    check(L.nsr_render_rays_forward_ex(ptr(rays), n, ptr(cfg['pc']), ptr(cfg['pf']), S, Ni, flags, ptr(cfg['t_rand']), ptr(cfg['u']),
                                       ptr(rgb), ptr(disp), ptr(acc), ptr(rgb0), ptr(disp0), ptr(acc0), ptr(zstd),
                                       ptr(raw), ptr(zv), None, ptr(mask), ptr(dump), ptr(aset), ptr(ws), ws_bytes, _stream()), 'nsr_render_rays_forward')
    z0 = raw0 = None
    if keep_for_backward and Ni > 0:
        # the coarse pass's depths and raw outputs are left in the workspace (nsr_render_workspace_layout: z0 | w0 | raw0 | ...).
        # (After a two-tier pass raw0 holds (0,0,0,sigma~) at the certified-empty points: alpha == 0 there either way, so a
        # backward pass through rgb0 gets the same -- zero -- dL/draw for them.)
        off = (ctypes.c_size_t * 8)()
        if L.nsr_render_workspace_layout(n, S, Ni, off, 8) != 8:
            check(-1, 'nsr_render_workspace_layout')
        z0 = ws[off[0]:off[0] + n * S * 4].view(torch.float32).view(n, S).clone()
        raw0 = ws[off[2]:off[2] + n * S * 16].view(torch.float32).view(n, S, 4).clone()
    outs = [rgb, disp, acc] + ([rgb0, disp0, acc0, zstd] if Ni > 0 else []) + ([raw] if cfg['retraw'] else [])
    return tuple(outs), (zv, raw, z0, raw0, mask, dump, aset)


def _params_of(net):
    if _is_viewless(net):
        return []                 # forward and dL/drays only (render_rays refuses trainable use_viewdirs=False networks)
    layers = _net_tensors(net)
    return [l.weight for l in layers] + [l.bias for l in layers]


class _RenderRaysFn(torch.autograd.Function):
    """render_rays as an autograd node.  Differentiable outputs: rgb_map and rgb0; differentiable inputs: ray_batch
    (the pose path, RN:177-178) and the parameters of both networks (the training step, RN:691-707).  disp / acc /
    z_std / raw are marked non-differentiable, so asking for their gradients fails loudly instead of returning zeros."""

    @staticmethod
    def forward(ctx, ray_batch, cfg, *params):
        rays = ray_batch.detach().to(torch.float32).contiguous()
        pose_only = not any(torch.is_tensor(p) and p.requires_grad for p in params)     # RN:168-181: no dL/dMLP wanted
        # pose path: keep the ReLU sign bits; training batches: the activations of the last pass too -- either way the backward
        # kernel of that pass recomputes nothing (whole images with parameter gradients do not fit: _mask_fits decides)
        # (rays that require grad = the pose path even if the modules' parameters carry the nn.Module default requires_grad=True, as
        # in the unpatched render_path_grad loop, RN:168-181: no activation dump then; should the caller ask for parameter gradients
        # after all, the backward pass recomputes what it needs)
        outs, saved = _forward_impl(rays, cfg, keep_for_backward=True, save_mask=True, save_dump=not pose_only and not ray_batch.requires_grad)
        ctx.relu_mask, ctx.dump, ctx.active_set = saved[4], saved[5], saved[6]
        ctx.save_for_backward(rays, *[t for t in saved[:4] if t is not None])
        ctx.have_coarse = saved[2] is not None
        ctx.cfg = cfg
        ctx.in_dtype = ray_batch.dtype
        ctx.set_materialize_grads(False)
        nondiff = [o for i, o in enumerate(outs) if not (i == 0 or (cfg['Ni'] > 0 and i == 3))]
        ctx.mark_non_differentiable(*nondiff)
        return outs

    @staticmethod
    def _one_pass(rays, zv, raw, net_blob, flags, g, want_dump, relu_mask=None, fwd_dump=None, active_set=None):
        L = lib()
        n, T = zv.shape
        d_rays = torch.empty(n, 11, dtype=torch.float32, device=rays.device)
        ws_bytes = L.nsr_render_backward_workspace_bytes(n, T)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=rays.device)
        dump = grads = dWp = dBp = None
        if want_dump:   # parameter gradients: the kernels ADD this pass's dL/dW, dL/db into zero-initialised fp32 tensors
            try:
                dump = fwd_dump if fwd_dump is not None else torch.empty(L.nsr_mlp_dump_bytes(n, T), dtype=torch.uint8, device=rays.device)
            except torch.cuda.OutOfMemoryError as e:
                raise _lib.NsrError(f'parameter gradients of {n} rays x {T} samples need {L.nsr_mlp_dump_bytes(n, T) >> 20} MiB of scratch '
                                    '(~10 KB per sample point): render in smaller chunks') from e
            shapes = _EXPECTED_SHAPES
            grads = [torch.zeros(s, dtype=torch.float32, device=rays.device) for s in shapes] + \
                    [torch.zeros(s[0], dtype=torch.float32, device=rays.device) for s in shapes]
            dWp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in grads[:12]])
            dBp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in grads[12:]])
        if want_dump and fwd_dump is None:
            relu_mask = active_set = None        # parameter gradients need every activation: recompute them (densely) unless the forward pass dumped them
        check(L.nsr_render_rays_backward_ex(ptr(rays), ptr(zv), ptr(raw), n, T, ptr(net_blob), flags, ptr(g), ptr(d_rays), ptr(dump),
                                            dWp, dBp, ptr(relu_mask), ptr(active_set), ptr(ws), ws_bytes, _stream()), 'nsr_render_rays_backward')
        return d_rays, grads

    @staticmethod
    def backward(ctx, *gouts):
        cfg = ctx.cfg
        saved = list(ctx.saved_tensors)
        rays, zv, raw = saved[0], saved[1], saved[2]
        z0, raw0 = (saved[3], saved[4]) if ctx.have_coarse else (None, None)
        d_rgb = gouts[0]
        d_rgb0 = gouts[3] if cfg['Ni'] > 0 else None
        n_par = 24
        need_rays = ctx.needs_input_grad[0]
        need_c = any(ctx.needs_input_grad[2:2 + n_par])
        need_f = any(ctx.needs_input_grad[2 + n_par:2 + 2 * n_par])
        fine_is_coarse = cfg['pf'] is None           # RN:481: no fine network -> the coarse one evaluates the last pass too
        wflag = cfg['flags'] & FLAG_WHITE_BKGD
        d_rays = None
        g_c = g_f = None
        if d_rgb is not None:
            blob = cfg['pc'] if fine_is_coarse else cfg['pf']
            want = need_c if fine_is_coarse else need_f
            dr, gr = _RenderRaysFn._one_pass(rays, zv, raw, blob, wflag, d_rgb.detach().float().contiguous(), want, ctx.relu_mask,
                                             ctx.dump if (want and ctx.relu_mask is not None) else None, ctx.active_set)
            d_rays = dr
            if fine_is_coarse:
                g_c = gr
            else:
                g_f = gr
        if d_rgb0 is not None and (need_rays or need_c):
            dr, gr = _RenderRaysFn._one_pass(rays, z0, raw0, cfg['pc'], wflag, d_rgb0.detach().float().contiguous(), need_c)
            d_rays = dr if d_rays is None else d_rays + dr
            if gr is not None:
                g_c = gr if g_c is None else [a + b for a, b in zip(g_c, gr)]
        out = [d_rays.to(ctx.in_dtype) if (d_rays is not None and need_rays) else None, None]
        out += g_c if g_c is not None else [None] * n_par
        if len(ctx.needs_input_grad) > 2 + n_par:
            out += g_f if g_f is not None else [None] * n_par
        return tuple(out[:len(ctx.needs_input_grad)])     # use_viewdirs=False networks pass no parameters (_params_of)


def _render_rays_staged(rays, pc, pf, S, Ni, flags, t_rand, u, retraw, raw_noise_std, white_bkgd, ret):
    """render_rays with raw_noise_std > 0: same kernels, launched stage by stage so that the noise
    (torch.randn * std, RN:366) can be added to sigma before each compositing step."""
    L = lib()
    n, dev, st = rays.shape[0], rays.device, _stream()
    new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    cflag = flags & FLAG_WHITE_BKGD
    # coarse z via the fused entry with Ni=0 would also composite; build z through resample-free path
    ws_bytes = L.nsr_render_workspace_bytes(n, S, 0)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    z0, raw0, w0 = new(n, S), new(n, S, 4), new(n, S)
    check(L.nsr_render_rays_forward(ptr(rays), n, ptr(pc), None, S, 0, flags, ptr(t_rand), None, None, None, None,
                                    None, None, None, None, ptr(raw0), ptr(z0), None, ptr(ws), ws_bytes, st), 'coarse pass')
    raw0[..., 3] += torch.randn(n, S, device=dev) * raw_noise_std
    tgt = ('rgb0', 'disp0', 'acc0') if Ni > 0 else ('rgb_map', 'disp_map', 'acc_map')
    check(L.nsr_raw2outputs(ptr(raw0), ptr(z0), ptr(rays[:, 3:6].contiguous()), 3, n, S, cflag, ptr(ret[tgt[0]]),
                            ptr(ret[tgt[1]]), ptr(ret[tgt[2]]), ptr(w0), None, st), 'nsr_raw2outputs')
    if Ni == 0:
        if retraw:
            ret['raw'] = raw0
        return ret
    T = S + Ni
    z1, raw1 = new(n, T), new(n, T, 4)
    check(L.nsr_resample_merge(ptr(z0), ptr(w0), n, S, Ni, ptr(u), ptr(z1), None, ptr(ret['z_std']), st), 'nsr_resample_merge')
    check(L.nsr_mlp_forward(ptr(rays), ptr(z1), n, T, ptr(pf if pf is not None else pc), flags & (FLAG_FAST_FP16 | FLAG_MIXED_F8), ptr(raw1), st), 'nsr_mlp_forward')
    raw1[..., 3] += torch.randn(n, T, device=dev) * raw_noise_std
    check(L.nsr_raw2outputs(ptr(raw1), ptr(z1), ptr(rays[:, 3:6].contiguous()), 3, n, T, cflag, ptr(ret['rgb_map']),
                            ptr(ret['disp_map']), ptr(ret['acc_map']), None, None, st), 'nsr_raw2outputs')
    if retraw:
        ret['raw'] = raw1
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """RN:43-55."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


# ----------------------------------------------------------------------------- RH:156-195
def get_rays(H, W, K, c2w):
    """RH:156-165 (torch ops on c2w's device; differentiable w.r.t. c2w for the psi path)."""
    dev = c2w.device if torch.is_tensor(c2w) else device
    c2w = torch.as_tensor(c2w, dtype=torch.float32, device=dev)
    j, i = torch.meshgrid(torch.linspace(0, H - 1, H, device=dev), torch.linspace(0, W - 1, W, device=dev), indexing='ij')
    dirs = torch.stack([(i - float(K[0][2])) / float(K[0][0]), -(j - float(K[1][2])) / float(K[1][1]), -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """RH:178-195: move the origins onto the near plane, then map origins / directions to normalised device
    coordinates (forward-facing scenes).  Same operation order as the reference, so the floats agree."""
    shift = -(near + rays_o[..., 2]) / rays_d[..., 2]
    org = rays_o + shift[..., None] * rays_d
    ox, oy, oz = org[..., 0], org[..., 1], org[..., 2]
    dx, dy, dz = rays_d[..., 0], rays_d[..., 1], rays_d[..., 2]
    sx, sy = -1. / (W / (2. * focal)), -1. / (H / (2. * focal))
    o_ndc = torch.stack([sx * ox / oz, sy * oy / oz, 1. + 2. * near / oz], -1)
    d_ndc = torch.stack([sx * (dx / dz - ox / oz), sy * (dy / dz - oy / oz), -2. * near / oz], -1)
    return o_ndc, d_ndc


def make_rays(H, W, K, c2w, near, far):
    """get_rays + viewdir normalisation + packing (RH:156-165, RN:91-112) in one kernel -> [H*W,11].  A pose that already lives
    on the GPU (the device pose sampler's output) is read there (nsr_make_rays_dev): no device->host copy, no synchronisation."""
    Kh = _K9(K)
    if torch.is_tensor(c2w) and c2w.is_cuda:
        c = c2w.detach()
        if c.dtype != torch.float32 or c.dim() != 2 or c.stride(-1) != 1 or c.stride(0) < 4 or c.shape[0] < 3 or c.shape[1] < 4:
            c = c.float()[:3, :4].contiguous()
        rays = torch.empty(H * W, 11, dtype=torch.float32, device=c.device)
        check(lib().nsr_make_rays_dev(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ptr(c), c.stride(0), float(near), float(far), ptr(rays), _stream()),
              'nsr_make_rays_dev')
        return rays
    c = c2w.detach().float().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w, dtype=np.float32)
    ch = np.ascontiguousarray(c[:3, :4], dtype=np.float32)
    rays = torch.empty(H * W, 11, dtype=torch.float32, device=device)
    check(lib().nsr_make_rays(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ch.ctypes.data_as(ctypes.c_void_p),
                              float(near), float(far), ptr(rays), _stream()), 'nsr_make_rays')
    return rays


# ----------------------------------------------------------------------------- RN:58-123
def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """RN:58-123 -> [rgb_map, disp_map, acc_map, extras].  Same arguments as the reference."""
    if not use_viewdirs:
        # RN:91/111: no view directions in the ray batch; the networks must be the RH:95-96 kind (the reference fails in RH:100 otherwise)
        nets_used = [kwargs.get('network_fn')] + ([kwargs['network_fine']] if kwargs.get('network_fine') is not None and kwargs.get('N_importance', 0) > 0 else [])
        if not all(m is not None and _is_viewless(m) for m in nets_used):
            raise ValueError('use_viewdirs=False needs NeRF(use_viewdirs=False) networks (RH:95-96)')
    scalar_bounds = not torch.is_tensor(near) and not torch.is_tensor(far)
    if c2w is not None and not ndc and c2w_staticcam is None and scalar_bounds and not (torch.is_tensor(c2w) and c2w.requires_grad and torch.is_grad_enabled()):
        packed = make_rays(H, W, K, c2w, near, far)          # fused ray generation
        sh = (H, W, 3)
    elif (c2w is None and not ndc and c2w_staticcam is None and scalar_bounds and torch.is_tensor(rays[1]) and rays[1].is_cuda and
          not (torch.is_grad_enabled() and (rays[0].requires_grad or rays[1].requires_grad))):
        # caller-made rays, nothing to differentiate: RN:97 + RN:106-112 in one kernel instead of five tensor ops
        rays_o, rays_d = rays
        sh = rays_d.shape
        o, d = _f32c(rays_o.reshape(-1, 3), 'rays_o'), _f32c(rays_d.reshape(-1, 3), 'rays_d')
        packed = torch.empty(o.shape[0], 11, dtype=torch.float32, device=d.device)
        check(lib().nsr_pack_rays(ptr(o), ptr(d), o.shape[0], float(near), float(far), ptr(packed), _stream()), 'nsr_pack_rays')
    else:
        if c2w is not None:
            rays_o, rays_d = get_rays(H, W, K, c2w)
        else:
            rays_o, rays_d = rays
        if use_viewdirs:                                                        # RN:91
            viewdirs = rays_d
            if c2w_staticcam is not None:
                rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)               # RN:94-96
            viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)    # RN:97
            viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
        sh = rays_d.shape
        if ndc:
            rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)        # RN:101-103
        rays_o = torch.reshape(rays_o, [-1, 3]).float()
        rays_d = torch.reshape(rays_d, [-1, 3]).float()
        nr, fr = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
        packed = torch.cat([rays_o, rays_d, nr, fr] + ([viewdirs] if use_viewdirs else []), -1)   # RN:109-112
        if not packed.is_cuda:
            raise _lib.NsrError('rays must live on the GPU (no CPU fallback)')
    all_ret = batchify_rays(packed, max(int(chunk), MIN_RAYS_PER_LAUNCH), **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))  # RN:116-118
    k_extract = ['rgb_map', 'disp_map', 'acc_map']
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


# ----------------------------------------------------------------------------- RN:126-255 (callers of render)
def _imwrite(filename, rgb8):
    """PNG writer: imageio when present (RN:206/250), else Pillow."""
    try:
        import imageio
        imageio.imwrite(filename, rgb8)
    except ImportError:
        from PIL import Image
        Image.fromarray(rgb8).save(filename)


def to8b_device(x):
    """RH:14 on the device: float tensor -> uint8 tensor of the same shape (255*clip(x,0,1), truncated)."""
    x = _f32c(x, 'x')
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(lib().nsr_to8b(ptr(x), x.numel(), ptr(out), _stream()), 'nsr_to8b')
    return out


def _K9(K):
    return np.ascontiguousarray(np.asarray([[float(K[r][c]) for c in range(3)] for r in range(3)], dtype=np.float32))


def _image_path_ok(kw):
    """Can render(H, W, K, c2w=..., **kw) go through the one-call image entry?  (what MAIN:109-114 + create_nerf's
    render_kwargs_test amount to: use_viewdirs, no NDC, deterministic sampling, scalar bounds)"""
    viewless = kw.get('network_fn') is not None and _is_viewless(kw['network_fn']) and (kw.get('network_fine') is None or _is_viewless(kw['network_fine']))
    return (bool(kw.get('use_viewdirs', False)) != viewless and not kw.get('ndc', True) and kw.get('c2w_staticcam') is None
            and not kw.get('perturb', 0.) and not kw.get('raw_noise_std', 0.) and not kw.get('retraw', False)
            and not torch.is_tensor(kw.get('near', 0.)) and not torch.is_tensor(kw.get('far', 1.))
            and kw.get('network_fn') is not None)


def render_image(H, W, K, c2w, want=('rgb8', 'rgb_map', 'disp_map'), **kw):
    """One image through nsr_render_image_forward: ray generation (c2w on the host or on the device), coarse + fine pass,
    compositing and to8b in one C call; returns {name: device tensor} for the names in `want`
    ('rgb8' [H,W,3] uint8, 'rgb_map' [H,W,3], 'disp_map', 'acc_map', 'rgb0', 'disp0', 'acc0', 'z_std' [H,W]) plus
    'rays' [H*W,11] (a view into the call's workspace).  `kw` = the reference's render_kwargs (RN:318-338 + near/far)."""
    if not _image_path_ok(kw):
        raise NotImplementedError('render_image covers ndc=False, perturb=0, raw_noise_std=0, scalar near/far (use_viewdirs matching the networks)')
    L = lib()
    net_c, net_f = kw['network_fn'], kw.get('network_fine')
    S, Ni = int(kw['N_samples']), int(kw.get('N_importance', 0))
    flags = (FLAG_LINDISP if kw.get('lindisp', False) else 0) | (FLAG_WHITE_BKGD if kw.get('white_bkgd', False) else 0) | _prec_flag()
    pc = packed_weights(net_c)
    pf = packed_weights(net_f) if (net_f is not None and Ni > 0) else None
    dev = pc.device
    n = H * W
    shapes = {'rgb8': ((H, W, 3), torch.uint8), 'rgb_map': ((H, W, 3), torch.float32), 'rgb0': ((H, W, 3), torch.float32)}
    out = {}
    for name in want:
        if name not in ('rgb8', 'rgb_map', 'disp_map', 'acc_map', 'rgb0', 'disp0', 'acc0', 'z_std'):
            raise ValueError(f'render_image: unknown output {name!r}')
        if Ni == 0 and name in ('rgb0', 'disp0', 'acc0', 'z_std'):
            raise ValueError(f'render_image: {name!r} needs N_importance > 0')
        shape, dt = shapes.get(name, ((H, W), torch.float32))
        out[name] = torch.empty(shape, dtype=dt, device=dev)
    ws_bytes = L.nsr_render_image_workspace_bytes(H, W, S, Ni)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    Kh = _K9(K)
    c2w_host = c2w_dev = None
    ld = 4
    keep = None
    if torch.is_tensor(c2w) and c2w.is_cuda:
        c = c2w.detach()
        if c.dtype != torch.float32 or c.stride(-1) != 1 or c.stride(0) < 4:
            c = c.float().contiguous()
        keep, c2w_dev, ld = c, ptr(c), c.stride(0)
    else:
        c = c2w.detach().float().numpy() if torch.is_tensor(c2w) else np.asarray(c2w, dtype=np.float32)
        keep = np.ascontiguousarray(c[:3, :4], dtype=np.float32)
        c2w_host = keep.ctypes.data_as(ctypes.c_void_p)
    g = lambda k: ptr(out.get(k))
    check(L.nsr_render_image_forward(H, W, Kh.ctypes.data_as(ctypes.c_void_p), c2w_host, c2w_dev, ld, float(kw['near']), float(kw['far']),
                                     ptr(pc), ptr(pf), S, Ni, flags, g('rgb8'), g('rgb_map'), g('disp_map'), g('acc_map'), g('rgb0'),
                                     g('disp0'), g('acc0'), g('z_std'), ptr(ws), ws_bytes, _stream()), 'nsr_render_image_forward')
    del keep
    out['rays'] = ws[:n * 44].view(torch.float32).view(n, 11)
    return out


def rays_grad_to_c2w(H, W, K, rays, d_rays, pixel_idx=None):
    """dL/d(ray_batch) [n,11] -> dL/dc2w [3,4] in closed form (nsr_rays_grad_to_c2w): the get_rays (RH:156-165) + RN:97 part
    of `torch.autograd.grad(batch_rays, categorical_prob, grad_outputs=dLdray)` (RN:179-181)."""
    L = lib()
    rays, d_rays = _f32c(rays, 'rays'), _f32c(d_rays, 'd_rays')
    n = rays.shape[0]
    if pixel_idx is not None:
        pixel_idx = pixel_idx.to(device=rays.device, dtype=torch.int32).contiguous()
    ws = torch.empty(L.nsr_c2w_grad_workspace_bytes(), dtype=torch.uint8, device=rays.device)
    out = torch.empty(3, 4, dtype=torch.float32, device=rays.device)
    Kh = _K9(K)
    check(L.nsr_rays_grad_to_c2w(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ptr(rays), ptr(d_rays), ptr(pixel_idx), n, ptr(out), 0,
                                 ptr(ws), _stream()), 'nsr_rays_grad_to_c2w')
    return out


def render_image_grad(H, W, K, c2w, g_rgb, rows=None, **kw):
    """One image forward + backward for the pose path (RN:148-181 for all H*W rays at once): returns (rgb_map [H,W,3],
    dL/dc2w [3,4]) given g_rgb = dL/drgb_map [H*W,3].  No autograd graph is built: rays come from the device-resident c2w,
    the fine pass is back-propagated by nsr_render_rays_backward and dL/d(ray_batch) is folded to the pose in closed form.
    rows = (r0, r1): only image rows [r0, r1) are rendered and back-propagated (g_rgb still covers the whole image); returns
    (rgb_map [r1-r0,W,3], this slice's contribution to dL/dc2w) -- contributions of disjoint slices add up to the image's
    gradient, which is how dist.py shards one image over several GPUs."""
    if not _image_path_ok(kw):
        raise NotImplementedError('render_image_grad covers ndc=False, perturb=0, raw_noise_std=0, scalar near/far (use_viewdirs matching the networks)')
    if PRECISION == 'fp16':
        raise NotImplementedError("backward is not built for NSR_PRECISION='fp16'")
    L = lib()
    net_c, net_f = kw['network_fn'], kw.get('network_fine')
    S, Ni = int(kw['N_samples']), int(kw.get('N_importance', 0))
    flags = (FLAG_LINDISP if kw.get('lindisp', False) else 0) | (FLAG_WHITE_BKGD if kw.get('white_bkgd', False) else 0)
    pc = packed_weights(net_c)
    pf = packed_weights(net_f) if (net_f is not None and Ni > 0) else None
    dev = pc.device
    c = c2w.detach().to(device=dev, dtype=torch.float32)[:3, :4].contiguous()
    Kh = _K9(K)
    g = g_rgb.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()
    if g.shape[0] != H * W:
        raise ValueError(f'g_rgb {tuple(g_rgb.shape)} does not match a {H}x{W} image')
    if rows is not None:
        r0, r1 = int(rows[0]), int(rows[1])
        if not (0 <= r0 < r1 <= H):
            raise ValueError(f'render_image_grad: rows {rows} outside a {H}-row image')
        rays = torch.empty(H * W, 11, dtype=torch.float32, device=dev)
        check(L.nsr_make_rays_dev(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ptr(c), 4, float(kw['near']), float(kw['far']), ptr(rays), _stream()),
              'nsr_make_rays_dev')
        part = rays[r0 * W:r1 * W]
        cfg = dict(pc=pc, pf=pf, S=S, Ni=Ni, flags=flags, t_rand=None, u=None, retraw=False)
        with torch.no_grad():
            outs, saved = _forward_impl(part, cfg, keep_for_backward=True, save_mask=True)
            d_rays, _ = _RenderRaysFn._one_pass(part, saved[0], saved[1], pf if pf is not None else pc, flags & FLAG_WHITE_BKGD,
                                                g[r0 * W:r1 * W].contiguous(), False, saved[4], None, saved[6])
            d_c2w = rays_grad_to_c2w(H, W, K, part, d_rays, pixel_idx=torch.arange(r0 * W, r1 * W, device=dev, dtype=torch.int32))
        return outs[0].view(r1 - r0, W, 3), d_c2w
    ws_bytes = L.nsr_render_image_grad_workspace_bytes(H, W, S, Ni)
    ws = None
    if SAVE_RELU_MASK and _mask_fits(ws_bytes, dev):
        try:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        except torch.cuda.OutOfMemoryError:
            ws = None                            # the staged recompute route below needs far less
    if ws is not None:
        # one C call: rays, both passes (saving the ReLU sign bits), backward without recompute, closed-form dL/dc2w
        rgb = torch.empty(H * W, 3, dtype=torch.float32, device=dev)
        d_c2w = torch.empty(3, 4, dtype=torch.float32, device=dev)
        check(L.nsr_render_image_grad(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ptr(c), 4, float(kw['near']), float(kw['far']), ptr(pc), ptr(pf),
                                      S, Ni, flags, ptr(g), ptr(rgb), ptr(d_c2w), 0, ptr(ws), ws_bytes, _stream()), 'nsr_render_image_grad')
        return rgb.view(H, W, 3), d_c2w
    # the recompute route, stage by stage (no extra memory for the sign bits)
    rays = torch.empty(H * W, 11, dtype=torch.float32, device=dev)
    check(L.nsr_make_rays_dev(H, W, Kh.ctypes.data_as(ctypes.c_void_p), ptr(c), 4, float(kw['near']), float(kw['far']), ptr(rays), _stream()),
          'nsr_make_rays_dev')
    cfg = dict(pc=pc, pf=pf, S=S, Ni=Ni, flags=flags, t_rand=None, u=None, retraw=False)
    with torch.no_grad():
        outs, saved = _forward_impl(rays, cfg, keep_for_backward=True, save_mask=False)
        d_rays, _ = _RenderRaysFn._one_pass(rays, saved[0], saved[1], pf if pf is not None else pc, flags & FLAG_WHITE_BKGD, g, False, None,
                                            None, saved[6])
        d_c2w = rays_grad_to_c2w(H, W, K, rays, d_rays)
    return outs[0].view(H, W, 3), d_c2w


class _AsyncImageWriter:
    """Pinned staging buffers + events: the D2H copy and the PNG encode of image i overlap the kernels of image i+1
    (SURVEY §8f N3)."""

    def __init__(self, depth=2):
        self.depth, self.slots, self.pending = depth, {}, []

    def _pinned(self, key, t, slot):
        k = (key, tuple(t.shape), t.dtype, slot)
        if k not in self.slots:
            self.slots[k] = torch.empty(t.shape, dtype=t.dtype, device='cpu', pin_memory=True)
        return self.slots[k]

    def submit(self, index, tensors, on_ready):
        """tensors: {name: device tensor}; on_ready(index, {name: numpy}) runs once the copies have landed."""
        slot = index % self.depth
        while len(self.pending) >= self.depth:
            self._retire()
        host = {}
        for name, t in tensors.items():
            h = self._pinned(name, t, slot)
            h.copy_(t, non_blocking=True)
            host[name] = h
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((index, host, ev, on_ready))

    def _retire(self):
        index, host, ev, on_ready = self.pending.pop(0)
        ev.synchronize()
        on_ready(index, {k: v.numpy().copy() for k, v in host.items()})

    def drain(self):
        while self.pending:
            self._retire()


def render_path(categorical_prob, render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, object_id=2,
                render_factor=0):
    """RN:213-255: render every pose under no_grad, optionally write <savedir>/<object_id>/NNN.png.  One C call per
    image (nsr_render_image_forward: rays from c2w, both passes, to8b on the device); the device->host copies and the
    PNG encode of an image overlap the kernels of the next one."""
    H, W, focal = hwf
    if render_factor != 0:                                         # RN:217-221
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    n_img = len(render_poses)
    rgbs, disps = [None] * n_img, [None] * n_img
    if savedir is not None:
        os.makedirs(os.path.join(savedir, str(object_id)), exist_ok=True)

    def on_ready(i, host):
        rgbs[i], disps[i] = host['rgb_map'], host['disp_map']
        if savedir is not None:
            _imwrite(os.path.join(savedir, str(object_id), '{:03d}.png'.format(i)), host['rgb8'])       # RN:245-250

    fast = _image_path_ok(render_kwargs)
    writer = _AsyncImageWriter()
    with torch.no_grad():
        if fast and torch.is_tensor(render_poses) and not render_poses.is_cuda:
            render_poses = render_poses.float()
        for i, c2w in enumerate(render_poses):
            if fast:
                want = ('rgb8', 'rgb_map', 'disp_map') if savedir is not None else ('rgb_map', 'disp_map')
                out = render_image(H, W, K, c2w[:3, :4], want=want, **render_kwargs)
                out.pop('rays')
            else:
                rgb, disp, acc, _ = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)    # RN:233
                out = {'rgb_map': rgb, 'disp_map': disp}
                if savedir is not None:
                    out['rgb8'] = to8b_device(rgb)
            writer.submit(i, out, on_ready)
        writer.drain()
    return np.stack(rgbs, 0), np.stack(disps, 0)


def render_path_grad(categorical_prob, render_poses, hwf, K, chunk, grad_E, render_kwargs, gt_imgs=None, savedir=None,
                     object_id=2, render_factor=0):
    """RN:126-210 with the image rendered and back-propagated in ONE pass instead of ceil(H*W/chunk) passes.

    The reference returns one dL/dpsi per `chunk` rays and its caller averages them all (MAIN:191), i.e. the estimator is
    sum_over_chunks(g_chunk) / (n_images * n_chunks).  Gradients add over rays, so the whole image's gradient equals
    sum_over_chunks(g_chunk); one entry per image, divided by n_chunks = ceil(H*W/chunk), leaves that mean unchanged.

    Per image: render_image_grad (no autograd tape over rays: dL/d(ray_batch) is folded to dL/dc2w [3,4] on the device in
    closed form), then ONE autograd.grad of the 12 pose entries w.r.t. categorical_prob through the pose sampler's graph
    (RN:179-181).  Render arguments outside the image path's envelope take the general autograd route (same numbers)."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    n_img = min(len(render_poses), len(grad_E))                                                 # RN:142
    rgbs, dLdpsis = [None] * n_img, []
    n_chunks = max(1, math.ceil(H * W / chunk))
    fast = _image_path_ok(render_kwargs) and PRECISION != 'fp16'
    if savedir is not None:
        os.makedirs(os.path.join(savedir, str(object_id), 'withgrad'), exist_ok=True)

    def on_ready(i, host):
        rgbs[i] = host['rgb_map']
        if savedir is not None:
            _imwrite(os.path.join(savedir, str(object_id), 'withgrad', '{:03d}.png'.format(i)), host['rgb8'])   # RN:200-206

    writer = _AsyncImageWriter()
    for i_pose in range(n_img):
        pose = render_poses[i_pose][:3, :4]
        g = grad_E[i_pose]['grad_E'][0]
        g = (g if torch.is_tensor(g) else torch.as_tensor(g)).to(device).permute(1, 2, 0).reshape(-1, 3).float()   # RN:154-155
        if fast and pose.is_cuda:
            rgb, d_c2w = render_image_grad(H, W, K, pose, g, **render_kwargs)                  # RN:148-178
            dLdpsi = torch.autograd.grad(pose, categorical_prob, grad_outputs=d_c2w.to(pose.dtype), retain_graph=True)   # RN:179-181
        else:
            rays_o, rays_d = get_rays(H, W, K, pose)                                            # RN:148 (graph-attached to psi)
            batch_rays = torch.stack([rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)], 0)         # RN:163, all rays at once
            rgb_p, _, _, _ = render(H, W, K, chunk=max(chunk, H * W), rays=batch_rays, retraw=True, **render_kwargs)   # RN:168-170
            dLdray = torch.autograd.grad(rgb_p, batch_rays, grad_outputs=g, retain_graph=True)  # RN:177-178
            dLdpsi = torch.autograd.grad(batch_rays, categorical_prob, grad_outputs=dLdray, retain_graph=True)       # RN:179-181
            rgb = rgb_p.detach().reshape(H, W, 3)
        dLdpsis.append(dLdpsi[0] / n_chunks)
        out = {'rgb_map': rgb}
        if savedir is not None:
            out['rgb8'] = to8b_device(rgb)
        writer.submit(i_pose, out, on_ready)
    writer.drain()
    return np.stack(rgbs, 0), [d.cpu().detach() for d in dLdpsis]


def create_nerf(args):
    """RN:257-340: build the coarse (+ fine) network, the Adam optimiser and the two render-kwargs dictionaries from the
    reference's argparse namespace, and resume from the newest `<basedir>/<expname>/*.tar` (or args.ft_path) unless
    args.no_reload.  Checkpoint keys as the reference writes them (RN:725-731): global_step, network_fn_state_dict,
    network_fine_state_dict, optimizer_state_dict.  Returns (render_kwargs_train, render_kwargs_test, start, grad_vars,
    optimizer).  Host-side plumbing only; geometries other than the one the kernels are built for are refused here."""
    _, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views = 0                                                                            # RN:263
    if args.use_viewdirs:
        _, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    geometry = dict(input_ch=input_ch, input_ch_views=input_ch_views, skips=[4], use_viewdirs=bool(args.use_viewdirs),
                    output_ch=5 if args.N_importance > 0 else 4)
    nets = [NeRF(D=args.netdepth, W=args.netwidth, **geometry).to(device)]
    if args.N_importance > 0:
        nets.append(NeRF(D=args.netdepth_fine, W=args.netwidth_fine, **geometry).to(device))
    if torch.cuda.is_available():
        for m in nets:                            # refuse unsupported depths / widths / multires up front
            _viewless_operands(m) if _is_viewless(m) else _net_tensors(m)
    grad_vars = [p for m in nets for p in m.parameters()]
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    start = 0
    ft = getattr(args, 'ft_path', None)
    if ft is not None and ft != 'None':
        found = [ft]
    else:
        run_dir = os.path.join(args.basedir, args.expname)
        found = [os.path.join(run_dir, f) for f in sorted(os.listdir(run_dir)) if 'tar' in f]
    print('Found ckpts', found)
    if found and not args.no_reload:
        print('Reloading from', found[-1])
        ckpt = torch.load(found[-1], map_location=device, weights_only=False)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        nets[0].load_state_dict(ckpt['network_fn_state_dict'])
        if len(nets) > 1:
            nets[1].load_state_dict(ckpt['network_fine_state_dict'])
    query = lambda inputs, viewdirs, network_fn: run_network(inputs, viewdirs, network_fn, netchunk=args.netchunk)   # RN:281-284
    train = dict(network_query_fn=query, perturb=args.perturb, N_importance=args.N_importance,
                 network_fine=nets[1] if len(nets) > 1 else None, N_samples=args.N_samples, network_fn=nets[0],
                 use_viewdirs=args.use_viewdirs, white_bkgd=args.white_bkgd, raw_noise_std=args.raw_noise_std)
    if args.dataset_type != 'llff' or args.no_ndc:                                               # RN:328-331
        print('Not ndc!')
        train['ndc'] = False
        train['lindisp'] = args.lindisp
    test = dict(train, perturb=False, raw_noise_std=0.)                                          # RN:333-335
    return train, test, start, grad_vars, optimizer


# ----------------------------------------------------------------------------- RN:691-707
_train_cache = weakref.WeakKeyDictionary()      # optimizer -> cached pointer tables


def _adam_state(optimizer, p):
    st = optimizer.state[p]
    if len(st) == 0:                                        # as torch.optim.Adam initialises it lazily
        st['step'] = torch.tensor(0.0, dtype=torch.float32, device='cpu')
        st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
    return st


def train_step(batch_rays, target_s, optimizer, near=0., far=1., seed=None, **kw):
    """The body of the reference's training iteration, RN:691-707, as ONE call of nsr_train_step:

        rgb, disp, acc, extras = render(H, W, K, chunk, rays=batch_rays, retraw=True, **render_kwargs_train)
        optimizer.zero_grad(); loss = img2mse(rgb, target_s) + img2mse(extras['rgb0'], target_s); loss.backward(); optimizer.step()

    batch_rays [2,n,3] = (rays_o, rays_d) and target_s [n,3] as RN:679-689 build them; `kw` = render_kwargs_train (+ near / far).
    `optimizer` is the torch.optim.Adam create_nerf returned: its exp_avg / exp_avg_sq / step state is what the kernel updates,
    so optimizer.state_dict() checkpoints stay in the reference's format (RN:725-731).  Random draws (perturb, raw_noise_std)
    come from the device Philox generator: `seed` (default torch.initial_seed()) and the Adam step count address them.
    Returns {'loss', 'img_loss', 'img_loss0', 'psnr', 'rgb'} as device tensors (no host synchronisation).  The learning-rate
    decay (RN:710-715) is the caller's, through optimizer.param_groups as in the reference."""
    if not isinstance(optimizer, torch.optim.Adam):
        raise NotImplementedError('train_step drives torch.optim.Adam (RN:287)')
    net_c, net_f = kw['network_fn'], kw.get('network_fine')
    S, Ni = int(kw['N_samples']), int(kw.get('N_importance', 0))
    if Ni == 0:
        net_f = None
    if not kw.get('use_viewdirs', False) or kw.get('ndc', True):
        raise NotImplementedError('train_step covers use_viewdirs=True, ndc=False')
    if len(optimizer.param_groups) != 1:
        raise NotImplementedError('train_step expects the single parameter group of RN:287')
    grp = optimizer.param_groups[0]
    if grp.get('weight_decay', 0) != 0 or grp.get('amsgrad', False) or grp.get('maximize', False):
        raise NotImplementedError('train_step implements plain Adam (no weight decay / amsgrad / maximize), as RN:287 configures it')
    nets = [net_c] + ([net_f] if net_f is not None else [])
    params = [p for m in nets for p in _params_of(m)]
    known = {id(p) for p in grp['params']}
    if any(id(p) not in known for p in params):
        raise ValueError('the optimizer does not own the parameters of network_fn / network_fine')
    L = lib()
    blobs = [packed_weights(m) for m in nets]
    key = tuple((p.data_ptr(), id(optimizer.state[p].get('exp_avg'))) for p in params) + tuple(b.data_ptr() for b in blobs)
    hit = _train_cache.get(optimizer)
    if hit is None or hit['key'] != key:
        tables, structs = [], []
        for m, blob in zip(nets, blobs):
            ps = _params_of(m)
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.NsrError('network parameters must be contiguous fp32 CUDA tensors (no CPU fallback)')
            sts = [_adam_state(optimizer, p) for p in ps]
            arr = lambda ts: (ctypes.c_void_p * 24)(*[t.data_ptr() for t in ts])
            t3 = (arr(ps), arr([s_['exp_avg'] for s_ in sts]), arr([s_['exp_avg_sq'] for s_ in sts]))
            tables.append(t3)
            structs.append(_lib.TrainNet(ctypes.cast(t3[0], ctypes.c_void_p), ctypes.cast(t3[1], ctypes.c_void_p),
                                         ctypes.cast(t3[2], ctypes.c_void_p), ctypes.c_void_p(blob.data_ptr())))
        key = tuple((p.data_ptr(), id(optimizer.state[p].get('exp_avg'))) for p in params) + tuple(b.data_ptr() for b in blobs)
        hit = {'key': key, 'tables': tables, 'structs': structs, 'ws': None}
        _train_cache[optimizer] = hit
    rays_o, rays_d = batch_rays[0], batch_rays[1]
    with torch.no_grad():
        rays_d = _f32c(rays_d, 'batch_rays').reshape(-1, 3)
        rays_o = _f32c(rays_o, 'batch_rays').reshape(-1, 3)
        n = rays_d.shape[0]
        rays = torch.empty(n, 11, dtype=torch.float32, device=rays_d.device)                     # RN:97 + RN:106-112, one kernel
        check(L.nsr_pack_rays(ptr(rays_o), ptr(rays_d), n, float(near), float(far), ptr(rays), _stream()), 'nsr_pack_rays')
        target = _f32c(target_s, 'target_s').reshape(-1, 3)
    if target.shape[0] != n:
        raise ValueError(f'target_s {tuple(target_s.shape)} does not match batch_rays {tuple(batch_rays.shape)}')
    step = int(optimizer.state[params[0]]['step']) + 1
    base = torch.initial_seed() if seed is None else int(seed)
    call_seed = (base * 0x9E3779B97F4A7C15 + step) & 0xFFFFFFFFFFFFFFFF
    ws_bytes = L.nsr_train_workspace_bytes(n, S, Ni)
    if hit['ws'] is None or hit['ws'].numel() < ws_bytes:
        hit['ws'] = torch.empty(ws_bytes, dtype=torch.uint8, device=rays.device)
    losses = torch.empty(2, dtype=torch.float32, device=rays.device)
    rgb = torch.empty(n, 3, dtype=torch.float32, device=rays.device)
    flags = (FLAG_LINDISP if kw.get('lindisp', False) else 0) | (FLAG_WHITE_BKGD if kw.get('white_bkgd', False) else 0)
    b1, b2 = grp['betas']
    structs = hit['structs']
    check(L.nsr_train_step(ptr(rays), ptr(target), n, ctypes.byref(structs[0]), ctypes.byref(structs[1]) if len(structs) > 1 else None,
                           S, Ni, flags, 1 if kw.get('perturb', 0.) else 0, float(kw.get('raw_noise_std', 0.)), call_seed,
                           float(grp['lr']), float(b1), float(b2), float(grp['eps']), step, ptr(losses), ptr(rgb),
                           ptr(hit['ws']), ws_bytes, _stream()), 'nsr_train_step')
    for p in params:                                            # what optimizer.step() does to the bookkeeping
        optimizer.state[p]['step'] += 1
        torch.autograd.graph.increment_version(p)
    for m, blob in zip(nets, blobs):                            # the blobs were re-packed in place: keep the cache in step
        ps = _params_of(m)
        _pack_cache[m] = (tuple((p.data_ptr(), p._version, p.device.index) for p in ps), blob)
    return {'loss': losses[0] + losses[1], 'img_loss': losses[0], 'img_loss0': losses[1], 'psnr': mse2psnr(losses[0:1])[0], 'rgb': rgb}


def install(reference_module, loops=False):
    """Monkey-patch a loaded reference `utils.run_nerf_noscale` module so its callers
    (render_path RN:233, render_path_grad RN:168, MAIN:128/184) run on this renderer.  With loops=True the two image
    loops themselves are replaced too (one launch sequence per image; same return values / same mean gradient)."""
    names = ['render', 'batchify_rays', 'render_rays', 'run_network', 'raw2outputs', 'sample_pdf', 'get_rays']
    if loops:
        names += ['render_path', 'render_path_grad']
    for name in names:
        setattr(reference_module, name, globals()[name])
    return reference_module
