"""ctypes binding of libnsr_b200.so (C ABI in include/nsr_b200.h).

The library is built in-tree by build.py (nvcc, sm_100a).  There is no CPU or
PyTorch fallback: if the library is missing or the device is not sm_100 the
calls raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NSR_LIB_PATH') or os.path.join(_HERE, 'libnsr_b200.so')   # NSR_LIB_PATH: A/B runs of two builds (tools/)

_lib = None

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_u32 = ctypes.c_uint32
c_vp = ctypes.c_void_p
c_size = ctypes.c_size_t

_SIGNATURES = {
    'nsr_version': (c_int, []),
    'nsr_last_error': (ctypes.c_char_p, []),
    'nsr_launch_count': (ctypes.c_uint64, []),
    'nsr_chunk_issue_order': (c_int, [c_int, c_vp, c_vp, c_int]),
    'nsr_packed_net_bytes': (c_size, []),
    'nsr_pack_net': (c_int, [c_vp, c_vp, c_vp, c_vp]),
    'nsr_mlp_forward': (c_int, [c_f32p, c_f32p, c_i64, c_int, c_vp, c_u32, c_f32p, c_vp]),
    'nsr_raw2outputs': (c_int, [c_f32p, c_f32p, c_f32p, c_int, c_i64, c_int, c_u32,
                                c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_vp]),
    'nsr_sample_pdf': (c_int, [c_f32p, c_f32p, c_i64, c_int, c_int, c_f32p, c_f32p, c_vp]),
    'nsr_resample_merge': (c_int, [c_f32p, c_f32p, c_i64, c_int, c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_vp]),
    'nsr_render_workspace_bytes': (c_size, [c_i64, c_int, c_int]),
    'nsr_render_rays_forward': (c_int, [c_f32p, c_i64, c_vp, c_vp, c_int, c_int, c_u32, c_f32p, c_f32p,
                                        c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                        c_vp, c_size, c_vp]),
    'nsr_relu_mask_bytes': (c_size, [c_i64, c_int]),
    'nsr_render_rays_forward_ex': (c_int, [c_f32p, c_i64, c_vp, c_vp, c_int, c_int, c_u32, c_f32p, c_f32p,
                                           c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                           c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    'nsr_render_rays_backward_ex': (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_int, c_vp, c_u32, c_f32p, c_f32p, c_vp, c_vp, c_vp, c_vp, c_vp,
                                            c_vp, c_size, c_vp]),
    'nsr_active_set_bytes': (c_size, [c_i64, c_int]),
    'nsr_render_workspace_layout': (c_int, [c_i64, c_int, c_int, c_vp, c_int]),
    'nsr_mlp_two_tier': (c_int, [c_f32p, c_f32p, c_i64, c_int, c_vp, c_f32p, c_vp, c_vp, c_int, c_vp]),
    'nsr_mlp_backward': (c_int, [c_f32p, c_f32p, c_i64, c_int, c_vp, c_f32p, c_f32p, c_vp, c_vp, c_vp]),
    'nsr_set_two_tier': (c_int, [c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float]),
    'nsr_get_two_tier': (c_int, [c_vp, c_vp, c_vp, c_vp]),
    'nsr_set_tier1_pair': (c_int, [c_int]),
    'nsr_set_coarse_refine': (c_int, [c_int]),
    'nsr_set_coarse_refine_limit': (c_int, [ctypes.c_float]),
    'nsr_set_coarse_refine_sigma': (c_int, [ctypes.c_float]),
    'nsr_coarse_refine_workspace_bytes': (c_size, [c_i64]),
    'nsr_coarse_refine': (c_int, [c_f32p, c_f32p, c_i64, c_int, c_vp, c_f32p, c_vp, c_size, c_vp]),
    'nsr_render_backward_workspace_bytes': (c_size, [c_i64, c_int]),
    'nsr_mlp_dump_bytes': (c_size, [c_i64, c_int]),
    'nsr_render_rays_backward': (c_int, [c_f32p, c_f32p, c_f32p, c_i64, c_int, c_vp, c_u32, c_f32p, c_f32p, c_vp, c_vp, c_vp, c_vp, c_size, c_vp]),
    'nsr_make_rays': (c_int, [c_int, c_int, c_vp, c_vp, ctypes.c_float, ctypes.c_float, c_f32p, c_vp]),
    'nsr_pack_rays': (c_int, [c_f32p, c_f32p, c_i64, ctypes.c_float, ctypes.c_float, c_f32p, c_vp]),
    'nsr_make_rays_dev': (c_int, [c_int, c_int, c_vp, c_f32p, c_int, ctypes.c_float, ctypes.c_float, c_f32p, c_vp]),
    'nsr_to8b': (c_int, [c_f32p, c_i64, c_vp, c_vp]),
    'nsr_c2w_grad_workspace_bytes': (c_size, []),
    'nsr_rays_grad_to_c2w': (c_int, [c_int, c_int, c_vp, c_f32p, c_f32p, c_vp, c_i64, c_f32p, c_int, c_vp, c_vp]),
    'nsr_render_image_workspace_bytes': (c_size, [c_int, c_int, c_int, c_int]),
    'nsr_render_image_forward': (c_int, [c_int, c_int, c_vp, c_vp, c_f32p, c_int, ctypes.c_float, ctypes.c_float, c_vp, c_vp,
                                         c_int, c_int, c_u32, c_vp, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                         c_vp, c_size, c_vp]),
    'nsr_render_image_grad_workspace_bytes': (c_size, [c_int, c_int, c_int, c_int]),
    'nsr_render_image_grad': (c_int, [c_int, c_int, c_vp, c_f32p, c_int, ctypes.c_float, ctypes.c_float, c_vp, c_vp, c_int, c_int, c_u32,
                                      c_f32p, c_f32p, c_f32p, c_int, c_vp, c_size, c_vp]),
    'nsr_random_uniform': (c_int, [ctypes.c_uint64, c_u32, c_f32p, c_i64, c_vp]),
    'nsr_add_sigma_noise': (c_int, [ctypes.c_uint64, c_u32, c_f32p, c_i64, ctypes.c_float, c_vp]),
    'nsr_train_workspace_bytes': (c_size, [c_i64, c_int, c_int]),
    'nsr_train_step': (c_int, [c_f32p, c_f32p, c_i64, c_vp, c_vp, c_int, c_int, c_u32, c_int, ctypes.c_float, ctypes.c_uint64,
                               ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_i64, c_f32p, c_f32p, c_vp, c_size, c_vp]),
}


class TrainNet(ctypes.Structure):
    """struct nsr_train_net (include/nsr_b200.h)."""
    _fields_ = [('params', c_vp), ('exp_avg', c_vp), ('exp_avg_sq', c_vp), ('packed', c_vp)]


EXPORTED_SYMBOLS = tuple(_SIGNATURES)

FLAG_LINDISP = 1
FLAG_WHITE_BKGD = 2
FLAG_PTS_INPUT = 4
FLAG_EMBEDDED_INPUT = 64
FLAG_FAST_FP16 = 8
FLAG_MIXED_F8 = 16
FLAG_DENSE = 32


class NsrError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NsrError(f'{LIB_PATH} not found: build it with `python -m neural_sim_nerf_b200.build` '
                           '(or __graft_entry__.build()); there is no CPU fallback')
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().nsr_last_error().decode('utf-8', 'replace')
        raise NsrError(f'{what} failed (code {rc}): {msg}')


def ptr(t):
    """Device pointer of a (contiguous fp32 CUDA) tensor, or None."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
