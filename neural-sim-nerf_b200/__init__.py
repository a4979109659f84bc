"""B200-native (sm_100a) NeRF per-ray renderer: drop-in for the render()/render_rays()/run_network()
surface of gyhandy/Neural-Sim-NeRF (optimization/utils/run_nerf_noscale.py, run_nerf_helpers.py).

The directory is named `neural-sim-nerf_b200`; import it as `neural_sim_nerf_b200`
(the repo-root shim neural_sim_nerf_b200.py registers it under that name).
"""
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, NsrError, lib  # noqa: F401
from .run_nerf import (NeRF, Embedder, batchify, batchify_rays, create_nerf, get_embedder, get_rays, img2mse, install,  # noqa: F401
                       make_rays, mse2psnr, ndc_rays, packed_weights, raw2outputs, rays_grad_to_c2w, render, render_image, render_image_grad, render_path, render_path_grad, render_rays,
                       run_network, sample_pdf, set_precision, to8b, to8b_device, train_step)
from .pose_sampler import pose_spherical, sample_pose, sample_pose_nograd  # noqa: F401

__all__ = ['NeRF', 'Embedder', 'batchify', 'batchify_rays', 'create_nerf', 'get_embedder', 'get_rays', 'install', 'make_rays',
           'ndc_rays', 'packed_weights', 'pose_spherical', 'sample_pose', 'sample_pose_nograd', 'raw2outputs', 'rays_grad_to_c2w', 'render', 'render_image', 'render_image_grad', 'render_path', 'render_path_grad', 'render_rays', 'run_network',
           'sample_pdf', 'set_precision', 'to8b', 'to8b_device', 'train_step']
