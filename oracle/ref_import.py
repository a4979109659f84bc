"""Import the UNMODIFIED reference renderer modules: from /root/reference in the build container, else from the bytecode
build of the same files in oracle/_ref/ (oracle/build_ref.py; it travels to the GPU box, /root/reference does not).

TEST INFRASTRUCTURE: used by oracle/make_golden.py, tests/test_oracle_pinned.py (live tree only) and bench.py's CPU baseline legs
(`cpu_baseline.kind: "reference"`).  The product package never imports it.

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
BUILT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def available():
    """The live reference tree (build container)."""
    return os.path.isdir(REF_ROOT)


def built():
    """The bytecode build of the reference modules (oracle/build_ref.py), importable sourceless."""
    return os.path.isfile(os.path.join(BUILT_ROOT, 'utils', 'run_nerf_noscale.pyc'))


def usable():
    return available() or built()


def load():
    """Returns (RN, RH): the reference's run_nerf_noscale and run_nerf_helpers modules."""
    import torch
    if not usable():
        raise RuntimeError('neither /root/reference nor oracle/_ref (python oracle/build_ref.py) is present')
    root = REF_ROOT if available() else BUILT_ROOT
    for name in ('imageio', 'matplotlib', 'matplotlib.pyplot', 'cv2', 'tqdm'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    # (with a GPU present, CPU runs of the reference go inside `with cpu_shim():`)
    if root not in sys.path:
        sys.path.insert(0, root)
    import utils.run_nerf_noscale as RN
    import utils.run_nerf_helpers as RH
    # RH:2 switches autograd anomaly mode on globally at import; leave the choice to the caller.
    return RN, RH


class cpu_shim:
    """Context manager: the reference's hard-coded `.cuda()` calls (RH:158,159,208; RN:359,363,366,376,439) become no-ops, so that
    its CPU path can be timed on a box that has a GPU.  Restores torch.Tensor.cuda on exit."""

    def __enter__(self):
        import torch
        self.saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self.saved
        return False


def build_models(sd_coarse, sd_fine):
    """Instantiate the reference's own NeRF modules (RH:70-97) and its
    network_query_fn closure (RN:281-284) around the given state-dicts."""
    RN, RH = load()
    embed_fn, input_ch = RH.get_embedder(10, 0)
    embeddirs_fn, input_ch_views = RH.get_embedder(4, 0)
    nets = []
    viewless = 'output_linear.weight' in sd_coarse        # use_viewdirs=False: RN:263 leaves input_ch_views = 0, embeddirs_fn = None
    if viewless:
        input_ch_views, embeddirs_fn = 0, None
    for sd in (sd_coarse, sd_fine):
        m = RH.NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4],
                    input_ch_views=input_ch_views, use_viewdirs=not viewless)
        m.load_state_dict(sd)
        nets.append(m)
    query = lambda inputs, viewdirs, network_fn: RN.run_network(
        inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)
    return nets[0], nets[1], query


def render_kwargs(sd_coarse, sd_fine, near, far, N_samples=64, N_importance=128):
    """The render_kwargs_test dict create_nerf would hand to render() (RN:318-338) plus near/far (MAIN:109-114)."""
    coarse, fine, query = build_models(sd_coarse, sd_fine)
    return dict(network_query_fn=query, perturb=False, N_importance=N_importance, network_fine=fine,
                N_samples=N_samples, network_fn=coarse, use_viewdirs='output_linear.weight' not in sd_coarse, white_bkgd=False,
                raw_noise_std=0., ndc=False, lindisp=False, near=near, far=far)
