"""BASELINE config 5 with a stub detector: the outer loop of optimization/neural_sim_main.py (MAIN:1144-1212) around this renderer.

detectron2 is not in this image, so the detector side -- create_dataset / train / inference / compute_inverse_hvp / compute_grad_E
(MAIN:1183-1196) -- is replaced by a stub that reads the PNGs the renderer wrote and returns grad_E ~ N(0, 1e-3) [1,3,H,W] per image
(SURVEY.md §8d C5).  Everything on the NeRF side is the real thing:

  epoch:  psi -> softmax(psi / 0.25)                                     MAIN:86-87
          sample_pose_nograd (K poses, logged noise)                     MAIN:91      (device sampler)
          render_path -> K PNGs                                          MAIN:128     (one C call per image, async PNG writer)
          [stub detector -> grad_E]
          sample_pose (replay, graph-attached to psi)                    MAIN:147
          render_path_grad -> dL/dpsi per image, mean                    MAIN:184-191 (saved-sign-bit backward, closed-form dL/dc2w)
          Momentum update of psi, learning-rate schedule                 MAIN:1203-1212

  python tools/bilevel_stub.py [--epochs 2] [--K 8] [--hw 400]
  torchrun --nproc-per-node N tools/bilevel_stub.py ...      poses sharded over the ranks, one all-reduce of dL/dpsi per epoch;
      the K % N remainder images are cut into row bands over groups of ranks (dist.plan_images), so K = 50 on 8 GPUs costs 6.25
      image-times per rank instead of 7

Prints one JSON line with the wall-clock split per epoch (rank 0).
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import neural_sim_nerf_b200 as nsr
from neural_sim_nerf_b200 import dist as nd

YCBV_NEAR, YCBV_FAR = 0.8103964843749999 - 0.5, 1.4297681884765627 + 0.5                      # LL:197-198, object 2
K400 = [[1333.3333740234375, 0.0, 195.43], [0.0, 1334.22, 200.63], [0.0, 0.0, 1.0]]                      # nerf_traindata_info.json K (object 2)


class MomentumPsi:
    """MAIN:1095-1110."""

    def __init__(self, lr, momentum=0.9):
        self.lr, self.momentum, self.v = lr, momentum, None

    def update(self, params, grads):
        params, grads = params.detach().cpu().numpy().astype(np.float32), grads.detach().cpu().numpy().astype(np.float32)
        self.v = (np.zeros_like(params) if self.v is None else self.momentum * self.v) - self.lr * grads
        return torch.tensor(params + self.v)


def lr_schedule(epoch, base_lr, max_epoch):                                                     # MAIN:1137-1141
    return base_lr * epoch / 5 if epoch <= 5 else base_lr * (1 - epoch / max_epoch)


def stub_detector(savedir, object_id, indices, H, W, seed):
    """Stands in for MAIN:1183-1196: loads every rendered PNG (the detector's input) and returns a fixed-seed grad_E per image."""
    from PIL import Image
    out = []
    for i in indices:
        img = np.asarray(Image.open(os.path.join(savedir, str(object_id), '{:03d}.png'.format(i))))
        assert img.shape == (H, W, 3) and img.dtype == np.uint8
        g = torch.Generator().manual_seed(seed * 100003 + i)
        out.append({'image_index': i, 'grad_E': torch.randn(1, 3, H, W, generator=g) * 1e-3})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--K', type=int, default=8, help='images per epoch (n_samples_K, MAIN:1342 uses 50)')
    ap.add_argument('--hw', type=int, default=400)
    ap.add_argument('--opt_lr', type=float, default=5e-5)
    ap.add_argument('--gumble_T', type=float, default=0.1)
    ap.add_argument('--psi_pose_cats_mode', type=int, default=5)
    args = ap.parse_args()

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    nets = []
    for pre in ('coarse/', 'fine/'):
        m = nsr.NeRF()
        m.load_state_dict({k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)})
        nets.append(m.to(dev).requires_grad_(False))
    H = W = args.hw
    s = H / 400.0
    K = [[K400[0][0] * s, 0.0, K400[0][2] * s], [0.0, K400[1][1] * s, K400[1][2] * s], [0.0, 0.0, 1.0]]
    hwf = [H, W, K[0][0]]
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=64, N_importance=128, network_fine=nets[1], use_viewdirs=True,
              ndc=False, near=YCBV_NEAR, far=YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False, lindisp=False)
    chunk = 512                                                                                  # CFG:25: sets the reference's mean scaling only

    psi = torch.full((8,), 0.02)
    psi[args.psi_pose_cats_mode - 1] = 0.86                                                      # MAIN:1164-1165
    opt = MomentumPsi(args.opt_lr)
    workdir = tempfile.mkdtemp(prefix='nsr_bilevel_')
    object_id = 2
    split = []
    for epoch in range(args.epochs):
        t = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # ---- 1. D_train: K images from the current psi (MAIN:1179-1180), every rank renders its share of the poses
        prob = torch.softmax(psi / 0.25, 0)
        poses, log = nsr.sample_pose_nograd(prob, args.K, args.gumble_T, seed=epoch, device=dev)
        (lo, hi), shared = nd.plan_images(args.K, rank, world)
        n_rem = args.K - (args.K // world) * world
        group = world // n_rem if n_rem else 1
        cap = -(-H // group)                                                                     # rows of the largest band
        savedir = os.path.join(workdir, 'epoch{:02d}_rank{}'.format(epoch, rank))
        os.makedirs(os.path.join(savedir, str(object_id)), exist_ok=True)
        if hi > lo:
            nsr.render_path(None, poses[lo:hi], hwf, K, chunk, kw, savedir=savedir, object_id=object_id)
        n_local = hi - lo                                                                        # PNGs this rank's detector reads
        if n_rem and world > 1:
            # remainder images: every rank of a group renders one row band; the bands meet on the owner rank, which writes the PNG
            band = torch.zeros(cap, W, 3, dtype=torch.uint8, device=dev)
            if shared:
                img, part, g_, owner = shared[0]
                r0, r1 = nd.row_band(H, part, g_)
                rr = nsr.make_rays(H, W, K, poses[img][:3, :4], YCBV_NEAR, YCBV_FAR)[r0 * W:r1 * W]
                with torch.no_grad():
                    rgb_b = nsr.render(H, W, K, chunk=1 << 20, rays=torch.stack([rr[:, 0:3], rr[:, 3:6]], 0), **kw)[0]
                band[:r1 - r0] = nsr.run_nerf.to8b_device(rgb_b).view(r1 - r0, W, 3)
            bands = [torch.empty_like(band) for _ in range(world)]
            dist.all_gather(bands, band)
            if shared and shared[0][1] == 0:
                img, _, g_, owner = shared[0]
                rows_ = [bands[owner + q][:nd.row_band(H, q, g_)[1] - nd.row_band(H, q, g_)[0]] for q in range(g_)]
                nsr.run_nerf._imwrite(os.path.join(savedir, str(object_id), '{:03d}.png'.format(n_local)), torch.cat(rows_, 0).cpu().numpy())
                n_local += 1
        elif n_rem:                                                                              # one rank: the tail images go out whole
            nsr.render_path(None, poses[hi:], hwf, K, chunk, kw, savedir=os.path.join(savedir, 'tail'), object_id=object_id)
            for q in range(args.K - hi):
                os.replace(os.path.join(savedir, 'tail', str(object_id), '{:03d}.png'.format(q)),
                           os.path.join(savedir, str(object_id), '{:03d}.png'.format(n_local + q)))
            n_local += args.K - hi
        torch.cuda.synchronize()
        t['render_images_s'] = time.perf_counter() - t0
        # ---- 2. detector (stub)
        t1 = time.perf_counter()
        grad_E = stub_detector(savedir, object_id, range(n_local), H, W, epoch * world + rank) if n_local else []
        shared_g = None
        if n_rem and world > 1:
            # the owner's detector saw the shared image: its grad_E goes to the ranks that hold the other bands
            for j in range(n_rem):
                buf = torch.zeros(3, H, W, device=dev)
                if rank == j * group:
                    buf.copy_(grad_E[-1]['grad_E'][0])
                dist.broadcast(buf, src=j * group)
                if shared and shared[0][0] == (args.K // world) * world + j:
                    shared_g = buf
        t['stub_detector_s'] = time.perf_counter() - t1
        # ---- 3. dL_val/dpsi = dI/dpsi . grad_E (MAIN:1199-1200), poses replayed with gradient
        t2 = time.perf_counter()
        prob_g = torch.softmax(psi.to(dev) / 0.25, 0).requires_grad_()
        poses_g = nsr.sample_pose(prob_g, args.K, args.gumble_T, log)
        dLdpsis, counts = [], []
        if hi > lo:
            _, dLdpsis = nsr.render_path_grad(prob_g, poses_g[lo:hi], hwf, K, chunk, grad_E[:hi - lo], kw)
            counts = [1.0] * len(dLdpsis)
        if n_rem and world > 1:
            if shared:
                img, part, g_, owner = shared[0]
                pose = poses_g[img][:3, :4]
                _, d_c2w = nsr.run_nerf.render_image_grad(H, W, K, pose, shared_g.permute(1, 2, 0).reshape(-1, 3), rows=nd.row_band(H, part, g_), **kw)
                d = torch.autograd.grad(pose, prob_g, grad_outputs=d_c2w.to(pose.dtype), retain_graph=True)[0]
                dLdpsis.append((d / max(1, -(-H * W // chunk))).cpu().detach())
                counts.append(1.0 / g_)
        elif n_rem:
            _, tail = nsr.render_path_grad(prob_g, poses_g[hi:], hwf, K, chunk, grad_E[hi - lo:], kw)
            dLdpsis += tail
            counts += [1.0] * len(tail)
        torch.cuda.synchronize()
        t['render_images_grad_s'] = time.perf_counter() - t2
        t3 = time.perf_counter()
        grad_psi = nd.reduce_psi_grad([g.to(dev) for g in dLdpsis], n_psi=8, counts=counts).cpu()   # MAIN:191 over all ranks
        t['reduce_s'] = time.perf_counter() - t3
        # ---- 4. update psi (MAIN:1203, 1212)
        psi = opt.update(psi, grad_psi)
        opt.lr = lr_schedule(epoch, args.opt_lr, args.epochs)
        t['epoch_s'] = time.perf_counter() - t0
        t['grad_psi_norm'] = float(grad_psi.norm())
        split.append(t)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        rays = args.K * H * W
        last = split[-1]
        print(json.dumps({'workload': f'bilevel outer loop with a stub detector: {args.epochs} epochs x K={args.K} images of {H}x{W}, 64+128 samples, '
                                      f'{world} GPU(s), poses sharded by rank, remainder images in row bands over rank groups',
                          'epochs': split,
                          'render_images_rays_per_s': rays / last['render_images_s'],
                          'render_images_grad_rays_per_s': rays / last['render_images_grad_s'],
                          'psi_final': [float(v) for v in psi]}))


if __name__ == '__main__':
    main()
