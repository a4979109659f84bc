"""Host-side multi-GPU logic (neural-sim-nerf_b200/dist.py) on CPU: world_size 2, gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def fake_render(rays):
    """Stands in for render_rays: a per-ray function, so sharding must not change it."""
    return {'rgb_map': torch.stack([rays[:, 0] * 2, rays[:, 1] + 1, rays[:, 2] ** 2], -1), 'acc_map': rays[:, 3].clone()}


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import neural_sim_nerf_b200.dist as D
    torch.manual_seed(0)
    rays = torch.randn(1001, 11)
    full = D.render_rays_sharded(rays, fake_render, gather=True)
    ok_render = all(torch.equal(full[k], fake_render(rays)[k]) for k in full)
    dealt = D.render_rays_sharded(rays, fake_render, gather=True, interleave=True)
    ok_render = ok_render and all(torch.equal(dealt[k], fake_render(rays)[k]) for k in dealt)
    mine = D.render_rays_sharded(rays, fake_render, gather=False)
    lo, hi = D.shard_bounds(1001, rank, world)
    ok_local = torch.equal(mine['acc_map'], rays[lo:hi, 3])
    # psi gradient: 7 chunk-gradients in total, split 4 / 3 across the ranks
    g = torch.Generator().manual_seed(1)
    chunks = [torch.randn(8, generator=g) for _ in range(7)]
    my_chunks = chunks[:4] if rank == 0 else chunks[4:]
    red = D.reduce_psi_grad(my_chunks)
    ok_psi = torch.allclose(red, torch.stack(chunks).mean(0), atol=1e-6)
    # fewer images than ranks: rank 1 renders nothing, the mean is still over the chunks that exist
    lone = D.reduce_psi_grad(chunks[:2] if rank == 0 else [])
    ok_psi = ok_psi and torch.allclose(lone, torch.stack(chunks[:2]).mean(0), atol=1e-6)
    # an image shared by both ranks as row bands (dist.plan_images): each band carries its share of the gradient and counts 1/2
    whole = [torch.randn(8, generator=g) for _ in range(2)]           # two whole-image entries, one per rank
    img = torch.randn(8, generator=g)                                 # the shared image's entry, split 30 / 70 between the bands
    band = img * (0.3 if rank == 0 else 0.7)
    mix = D.reduce_psi_grad([whole[rank], band], counts=[1.0, 0.5])
    ok_psi = ok_psi and torch.allclose(mix, (whole[0] + whole[1] + img) / 3.0, atol=1e-6)
    # device of the collective follows the backend, not the inputs (render_path_grad hands back CPU tensors, RN:190)
    ok_psi = ok_psi and D._collective_device().type == 'cpu' and red.device.type == 'cpu'
    poses, (plo, phi) = D.shard_poses(list(range(50)))
    grads = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))]
    D.all_reduce_grads_(grads)
    ok_grads = torch.equal(grads[0], torch.full((3,), 3.0)) and torch.equal(grads[1], torch.full((2, 2), 30.0))
    q.put((rank, ok_render, ok_local, ok_psi, ok_grads, len(poses)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    import neural_sim_nerf_b200.dist as D
    for n in (0, 1, 7, 160000, 160001):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_images_covers_every_image_once():
    import neural_sim_nerf_b200.dist as D
    for K in (1, 7, 8, 9, 50, 64):
        for w in (1, 2, 4, 8):
            plans = [D.plan_images(K, r, w) for r in range(w)]
            whole = (K // w) * w
            spans = [p[0] for p in plans]
            assert spans[0][0] == 0 and spans[-1][1] == whole and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert len({b - a for a, b in spans}) == 1                                   # no ragged tail among the whole images
            shared = [e for p in plans for e in p[1]]
            for img in range(whole, K):
                parts = sorted(e[1] for e in shared if e[0] == img)
                n_parts = {e[2] for e in shared if e[0] == img}
                assert len(n_parts) == 1 and parts == list(range(n_parts.pop()))          # every band exactly once
                assert {e[3] for e in shared if e[0] == img} == {min(r for r in range(w) if plans[r][1] and plans[r][1][0][0] == img)}
            assert all(whole <= e[0] < K for e in shared)
            for H in (400, 7):
                for n_parts in (1, 3, 4):
                    bands = [D.row_band(H, q, n_parts) for q in range(n_parts)]
                    assert bands[0][0] == 0 and bands[-1][1] == H and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    # 50 images on 8 GPUs: 6 whole images per rank + a quarter of one of the two remainder images
    (lo, hi), shared = D.plan_images(50, 5, 8)
    assert (lo, hi) == (30, 36) and shared == [(49, 1, 4, 4)]


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert all(r[1:5]), r
    assert sum(r[5] for r in res) == 50


def test_reduce_psi_grad_single_process_edge_cases():
    import neural_sim_nerf_b200.dist as D
    g = [torch.arange(8.), torch.ones(8)]
    assert torch.allclose(D.reduce_psi_grad(g), (g[0] + g[1]) / 2)
    with pytest.raises(ValueError):
        D.reduce_psi_grad([])
    assert torch.equal(D.reduce_psi_grad([], n_psi=8), torch.zeros(8))
