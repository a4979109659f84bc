"""Worker of tests/test_gpu_nccl.py: run as `python -m torch.distributed.run --nproc-per-node 2 tests/nccl_worker.py` on a box with
>= 2 GPUs.  Exercises neural-sim-nerf_b200/dist.py on the NCCL backend with the REAL renderer:
  1. render_rays_sharded(gather=True): the gathered maps are bit-equal to one rank rendering all rays (results do not depend on
     how rays are batched);
  2. all_reduce_grads_: dL/dMLP of both networks (2 x 595 844 floats, one 4.77 MB bucket) from a batch sharded over the ranks
     equals the gradient of the whole batch on one rank;
  3. reduce_psi_grad with CPU inputs under NCCL (the device of the collective follows the backend) and with band counts;
  4. one image as two row bands (render_image_grad(rows=...)): the bands' dL/dc2w, all-reduced, equal the whole image's.
Prints 'NCCL_WORKER_OK rank r' per rank."""
import functools
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O  # noqa: E402
import neural_sim_nerf_b200 as nsr  # noqa: E402
from neural_sim_nerf_b200 import dist as D  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    nets = []
    for pre in ('coarse/', 'fine/'):
        m = nsr.NeRF()
        m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)})
        nets.append(m.to(dev))
    H = W = 400
    c2w = O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4]
    rays = nsr.make_rays(H, W, O.YCBV_K_400, c2w, O.YCBV_NEAR, O.YCBV_FAR)[::37][:4001].contiguous()
    fn = functools.partial(nsr.render_rays, network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1])
    # 1. sharded render + all_gather
    with torch.no_grad():
        full = fn(rays)
        got = D.render_rays_sharded(rays, lambda r: {k: v for k, v in fn(r).items() if k in ('rgb_map', 'acc_map', 'rgb0')}, gather=True)
    for k, v in got.items():
        assert v.shape == full[k].shape and torch.equal(v, full[k]), k
    with torch.no_grad():
        dealt = D.render_rays_sharded(rays, lambda r: {k: v for k, v in fn(r).items() if k in ('rgb_map', 'acc_map')}, gather=True, interleave=True)
    for k, v in dealt.items():
        assert torch.equal(v, full[k]), k
    # 2. dL/dMLP: sharded batch + one all-reduce == whole batch
    n = 1024
    batch = rays[:n]
    target = torch.rand(n, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    def grads_of(r, t, scale):
        for m in nets:
            m.zero_grad(set_to_none=True)
        out = fn(r)
        loss = (((out['rgb_map'] - t) ** 2).sum() + ((out['rgb0'] - t) ** 2).sum()) * scale
        loss.backward()
        return [p.grad.clone() for m in nets for p in m.parameters()]
    ref = grads_of(batch, target, 1.0 / (3 * n))
    lo, hi = D.shard_bounds(n, rank, world)
    mine = grads_of(batch[lo:hi], target[lo:hi], 1.0 / (3 * n))
    assert sum(g.numel() for g in mine) == 2 * 595844
    D.all_reduce_grads_(mine)
    for a, b in zip(mine, ref):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-12, (float((a - b).abs().max()), float(b.abs().max()))
    # 3. psi gradient: CPU inputs, NCCL backend, band counts
    g = torch.Generator().manual_seed(1)
    whole = [torch.randn(8, generator=g) for _ in range(world)]
    img = torch.randn(8, generator=g)
    red = D.reduce_psi_grad([whole[rank], img / world], counts=[1.0, 1.0 / world])
    assert red.is_cuda and torch.allclose(red.cpu(), (torch.stack(whole).sum(0) + img) / (world + 1), atol=1e-6)
    # 4. one image over the ranks as row bands
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True, ndc=False,
              near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
    Hs = Ws = 64
    Ks = [[O.YCBV_K_400[0][0] * 0.16, 0, O.YCBV_K_400[0][2] * 0.16], [0, O.YCBV_K_400[1][1] * 0.16, O.YCBV_K_400[1][2] * 0.16], [0, 0, 1]]
    g_rgb = torch.randn(Hs * Ws, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
    pose = c2w.to(dev)
    rgb_all, d_all = nsr.run_nerf.render_image_grad(Hs, Ws, Ks, pose, g_rgb, **kw)
    r0, r1 = D.row_band(Hs, rank, world)
    rgb_band, d_band = nsr.run_nerf.render_image_grad(Hs, Ws, Ks, pose, g_rgb, rows=(r0, r1), **kw)
    assert torch.equal(rgb_band, rgb_all[r0:r1])
    dist.all_reduce(d_band)
    assert float((d_band - d_all).abs().max()) <= 1e-5 * float(d_all.abs().max()), (d_band, d_all)
    dist.barrier()
    print(f'NCCL_WORKER_OK rank {rank}', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
