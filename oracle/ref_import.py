"""Import the UNMODIFIED reference renderer modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing
that runs there may call this; it is used by oracle/make_golden.py and
tests/test_oracle_pinned.py (skipped when the tree is absent).

Recipe:
  * sys.path gets /root/reference/optimization so `utils.run_nerf_noscale` resolves;
  * `imageio`, `matplotlib`, `matplotlib.pyplot` are absent from this image and only
    used for I/O / plotting -> empty stub modules;
  * the hard-coded `.cuda()` calls (RH:158,159,208; RN:359,363,366,376,439) are made
    no-ops on a CPU-only box by shimming torch.Tensor.cuda.
Nothing in the reference files is edited or copied.
"""
import os
import sys
import types

REF_ROOT = '/root/reference/optimization'


def available():
    return os.path.isdir(REF_ROOT)


def load():
    """Returns (RN, RH): the reference's run_nerf_noscale and run_nerf_helpers modules."""
    import torch
    if not available():
        raise RuntimeError('reference tree not present (expected in the build container only)')
    for name in ('imageio', 'matplotlib', 'matplotlib.pyplot', 'cv2'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import utils.run_nerf_noscale as RN
    import utils.run_nerf_helpers as RH
    # RH:2 switches autograd anomaly mode on globally at import; leave the choice to the caller.
    return RN, RH


def build_models(sd_coarse, sd_fine):
    """Instantiate the reference's own NeRF modules (RH:70-97) and its
    network_query_fn closure (RN:281-284) around the given state-dicts."""
    RN, RH = load()
    embed_fn, input_ch = RH.get_embedder(10, 0)
    embeddirs_fn, input_ch_views = RH.get_embedder(4, 0)
    nets = []
    for sd in (sd_coarse, sd_fine):
        m = RH.NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4],
                    input_ch_views=input_ch_views, use_viewdirs=True)
        m.load_state_dict(sd)
        nets.append(m)
    query = lambda inputs, viewdirs, network_fn: RN.run_network(
        inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
    return nets[0], nets[1], query


def render_kwargs(sd_coarse, sd_fine, near, far, N_samples=64, N_importance=128):
    """The render_kwargs_test dict create_nerf would hand to render() (RN:318-338) plus near/far (MAIN:109-114)."""
    coarse, fine, query = build_models(sd_coarse, sd_fine)
    return dict(network_query_fn=query, perturb=False, N_importance=N_importance, network_fine=fine,
                N_samples=N_samples, network_fn=coarse, use_viewdirs=True, white_bkgd=False,
                raw_noise_std=0., ndc=False, lindisp=False, near=near, far=far)
