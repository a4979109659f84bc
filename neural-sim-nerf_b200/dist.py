"""Multi-GPU plumbing for the render path: one process per GPU, rays or whole images
sharded with NO forward collective; the only exchange is one all-reduce(SUM) per outer step on the
tiny dL/dpsi (8 floats) -- and dL/dMLP when the NeRF weights train -- plus an optional gather of the
rendered pixels to the rank that writes the PNGs (RN:245-250).

The reference is single-process (MAIN:1363-1383); the scaling rule it implies is kept: the psi
gradient is the MEAN over all per-chunk gradients of all images (MAIN:191), so ranks all-reduce
(sum of chunk gradients, number of chunks) and divide once.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous, balanced [lo, hi) of `n` items for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_poses(render_poses, rank=None, world_size=None):
    """The K poses of one epoch (MAIN:1342) are independent images: rank r renders poses[lo:hi]."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(len(render_poses), rank, world_size)
    return render_poses[lo:hi], (lo, hi)


def plan_images(n_images, rank=None, world_size=None):
    """K images of one epoch over the ranks without a ragged tail: the first (K // ws) * ws images go out whole, contiguous and
    balanced ([lo, hi) for this rank); each of the K % ws remainder images is cut into row bands over a group of ws // (K % ws)
    ranks, so that 50 images on 8 GPUs cost 6.25 image-times per rank instead of 7.  Returns ((lo, hi), shared) with
    shared = [(image_index, part, n_parts, owner_rank), ...] for this rank (at most one entry); owner_rank holds part 0."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    whole = (n_images // world_size) * world_size
    lo, hi = shard_bounds(whole, rank, world_size)
    rem = n_images - whole
    shared = []
    if rem:
        group = world_size // rem            # ranks per remainder image (ranks beyond rem * group stay idle for it)
        j, part = divmod(rank, group)
        if j < rem:
            shared.append((whole + j, part, group, j * group))
    return (lo, hi), shared


def row_band(H, part, n_parts):
    """Rows [r0, r1) of an H-row image for band `part` of `n_parts` (balanced, contiguous)."""
    return shard_bounds(H, part, n_parts)


def render_rays_sharded(rays_flat, render_fn, gather=True, interleave=False):
    """Render a [N,11] ray batch with every rank taking a contiguous slice.  `render_fn(rays) -> dict`
    (e.g. functools.partial(render_rays, **render_kwargs)).  With gather=True every rank returns the
    full-size maps (all_gather of the per-rank slices); otherwise only its own slice.
    interleave=True deals the rays out round-robin instead (ray i -> rank i % world_size): the renderer's work per ray depends on
    what the ray hits (two-tier evaluation), so contiguous slices of an image are unevenly loaded and a strided deal is not."""
    rank, ws = world()
    n = rays_flat.shape[0]
    if interleave:
        local = render_fn(rays_flat[rank::ws].contiguous())
        if ws == 1 or not gather:
            return local
        out = {}
        cap = (n + ws - 1) // ws
        for k, v in local.items():
            padded = torch.zeros((cap,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            padded[:v.shape[0]] = v
            parts = [torch.empty_like(padded) for _ in range(ws)]
            dist.all_gather(parts, padded)
            full = torch.empty((n,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            for r, p in enumerate(parts):
                full[r::ws] = p[:len(range(r, n, ws))]
            out[k] = full
        return out
    lo, hi = shard_bounds(n, rank, ws)
    local = render_fn(rays_flat[lo:hi])
    if ws == 1 or not gather:
        return local
    out = {}
    sizes = [shard_bounds(n, r, ws) for r in range(ws)]
    cap = max(b - a for a, b in sizes)
    for k, v in local.items():
        # all_gather wants equal shapes: pad every slice to the largest one, trim after
        padded = torch.zeros((cap,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        padded[:v.shape[0]] = v
        parts = [torch.empty_like(padded) for _ in range(ws)]
        dist.all_gather(parts, padded)
        out[k] = torch.cat([p[:b - a] for p, (a, b) in zip(parts, sizes)], 0)
    return out


def _collective_device():
    """Where tensors must live for the active backend's collectives: NCCL only moves CUDA tensors, gloo takes CPU ones."""
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def reduce_psi_grad(chunk_grads, n_psi=None, counts=None):
    """chunk_grads: this rank's list of per-chunk dL/dpsi tensors (what render_path_grad returns as
    `dLdpsis`, RN:190 -- CPU tensors, `.cpu().detach()`; any device is accepted).  Returns the reference's estimator over ALL
    ranks: mean over every chunk of every image (MAIN:191) -- one all_reduce(SUM) of [sum, count] on the device the backend
    needs (CUDA under NCCL, CPU under gloo), result on that device.  A rank with no chunks contributes zeros; `n_psi` (length
    of psi) is only needed when NO rank may have any chunk.  `counts` (one number per entry, default 1): how many entries of the
    mean an entry stands for -- a row band of an image shared by g ranks (plan_images) carries that band's share of the image's
    gradient and counts 1/g, so that the g bands together weigh like one image."""
    if counts is not None and len(counts) != len(chunk_grads):
        raise ValueError('reduce_psi_grad: counts must match chunk_grads')
    total = float(sum(counts)) if counts is not None else float(len(chunk_grads))
    if len(chunk_grads):
        s = torch.stack([g.detach().reshape(-1).to(torch.float64) for g in chunk_grads], 0).sum(0)
    else:
        s = None
    rank, ws = world()
    if ws == 1:
        if s is None:
            if n_psi is None:
                raise ValueError('reduce_psi_grad: no chunk gradients (render_path_grad returned an empty list) and n_psi not given')
            return torch.zeros(n_psi, dtype=torch.float32)
        return (s / total).to(torch.float32)
    dev = _collective_device()
    width = torch.tensor([s.numel() if s is not None else (n_psi or 0)], device=dev)
    dist.all_reduce(width, op=dist.ReduceOp.MAX)
    buf = torch.zeros(int(width.item()) + 1, dtype=torch.float64, device=dev)
    if s is not None:
        buf[:-1] = s.to(dev)
        buf[-1] = total
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return (buf[:-1] / buf[-1].clamp(min=1)).to(torch.float32)


def all_reduce_grads_(tensors):
    """In-place SUM all-reduce of a list of gradient tensors as one flat bucket (dL/dMLP: 2 x 595 844
    fp32 = 4.77 MB -- latency-bound on NVLink 5, so one bucket, no compute/communication fusion)."""
    rank, ws = world()
    if ws == 1 or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors]).to(_collective_device())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return tensors
