"""bench.py -- rays/sec of the NeRF render hot path (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            this repo's sm_100a renderer
  python bench.py --impl reference --gpus N ...            the reference algorithm on the host CPU (oracle port)

Workload (config.workload): BASELINE config 2 -- one 400x400 image = 160 000 rays per step and per GPU,
64 coarse + 128 fine samples, deterministic resampling, weights tests/golden/wfit.npz (the pretrained
YCB-V checkpoints are not available offline; throughput does not depend on the weights: the reference
has no early termination).  A step = one forward render of all rays of one image.

One JSON line on stdout (rank 0).  `value` = whole-job rays/s with the rays resident in HBM;
`e2e` = the same through the public render(rays=...) call with pinned-host rays copied in and
rgb/disp/acc copied out every step; `roofline` = the dominant MLP kernel against the measured
bf16 tensor peak (`roofline_kernels`: every MLP kernel of the forward and backward passes, `roofline_step`:
the whole step in reference-algorithm FLOPs); `cpu_baseline` = the oracle port on the host cores on a bounded
ray sample; `parity` = the timed image's pixels against those CPU-rendered rays (the run FAILS above 1e-3);
`baselines` = the other BASELINE configs on the host cores and the eager-PyTorch path on this GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

H = W = 400
N_SAMPLES, N_IMPORTANCE = 64, 128
RAYS_PER_IMAGE = H * W
FLOP_PER_POINT = 2 * 593408                      # SURVEY.md §8(d) / BASELINE.md §3
FLOP_PER_POINT_TIER1 = 2 * (63 * 256 + 4 * 256 * 256 + 319 * 256 + 2 * 256 * 256 + 256)   # pts_linears.0-7 + alpha head (RH:99-109)
K_200 = [[666.6666870117188, 0.0, 97.715], [0.0, 667.11, 100.315], [0.0, 0.0, 1.0]]          # BASELINE config 1: the 400x400 camera / 2
FLOP_PER_RAY = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_PER_POINT
METRIC = 'rays/sec (64c+128f samples, 400x400)'


def load_weights():
    import numpy as np
    import torch
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sdc = {k[len('coarse/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('coarse/')}
    sdf = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
    return sdc, sdf


def pose_for(step, rank):
    import nerf_oracle as O
    phi = 22.5 + 45.0 * ((step + 3 * rank) % 8)      # the 8 azimuth bins of LL:269
    return O.pose_spherical(90., phi - 180., 1.01)[:3, :4]


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples if len(s) > 2 + i)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.samples)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_kind():
    """'reference': the UNMODIFIED reference modules (RN.render and its own NeRF nn.Modules), live from /root/reference or from their
    bytecode build in oracle/_ref (oracle/build_ref.py -- what the GPU box has); 'port': the oracle restatement, when neither exists."""
    import ref_import
    return 'reference' if ref_import.usable() else 'port'


def _cpu_renderer(N_importance=N_IMPORTANCE):
    """run(rays [2,n,3], **extra) -> [rgb, disp, acc, extras] on the host cores through the reference's own render() (RN:58-123) with
    chunk = 512 (CFG:25) and netchunk = 65536, or through the oracle port of it."""
    import nerf_oracle as O
    sdc, sdf = load_weights()
    if cpu_kind() == 'reference':
        import ref_import
        RN, _ = ref_import.load()
        kw = ref_import.render_kwargs(sdc, sdf, O.YCBV_NEAR, O.YCBV_FAR, N_samples=N_SAMPLES, N_importance=N_importance)
        return lambda r, **extra: RN.render(H, W, O.YCBV_K_400, chunk=512, rays=r, **dict(kw, **extra))
    kw = dict(near=O.YCBV_NEAR, far=O.YCBV_FAR, N_samples=N_SAMPLES, N_importance=N_importance)
    return lambda r, **extra: O.render(H, W, O.YCBV_K_400, sdc, sdf if N_importance > 0 else None, chunk=512, rays=r, **dict(kw, **extra))


def _cpu_rays(n_rays, pose_index=0, rank=0, HW=(H, W), K=None):
    import torch
    import nerf_oracle as O
    ro, rd = O.get_rays(HW[0], HW[1], K if K is not None else O.YCBV_K_400, pose_for(pose_index, rank))
    sel = torch.linspace(0, HW[0] * HW[1] - 1, n_rays).long()
    return torch.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0), sel


def _cpu_setup(n_rays, pose_index=0, rank=0):
    rays, _ = _cpu_rays(n_rays, pose_index, rank)
    return rays, _cpu_renderer()


def pick_cpu_threads(run, rays):
    """torch's intra-op pool with every hardware thread is far from the fastest setting for the
    reference's 512-ray chunks (a 128-thread box measured 34 rays/s vs >1000 with 16 threads): time one
    chunk per candidate and keep the best, so the CPU baseline is the reference at its best."""
    import torch
    ncpu = os.cpu_count() or 8
    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = cands[-1], float('inf')
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            run(rays[:, :512])
            t0 = time.perf_counter()
            run(rays[:, :512])
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_render_rate(n_rays, repeats, pose_index=0, rank=0):
    """The reference algorithm (oracle restatement, same chunk=512 / netchunk=65536 structure as
    RN:43-55 / RN:14-23) on the host cores; returns rays/s over `repeats` passes of `n_rays` rays, the thread count, the
    pass times and the rendered maps of the last pass (what `parity` compares the GPU image with)."""
    import torch
    import ref_import
    with ref_import.cpu_shim():
        rays, run = _cpu_setup(n_rays, pose_index, rank)
        threads = pick_cpu_threads(run, rays)
        times = []
        out = None
        with torch.no_grad():
            for _ in range(repeats):
                t0 = time.perf_counter()
                out = run(rays)
                times.append(time.perf_counter() - t0)
    return n_rays / (sum(times) / len(times)), threads, times, out


def cpu_config1_rate(n_rays):
    """BASELINE config 1 on the host cores: 200x200 camera, 64 coarse samples only (RN:58 with N_importance = 0), chunk 512."""
    import torch
    rays, _ = _cpu_rays(n_rays, HW=(200, 200), K=K_200)
    run = _cpu_renderer(N_importance=0)
    with torch.no_grad():
        run(rays[:, :512])
        t0 = time.perf_counter()
        run(rays)
        dt = time.perf_counter() - t0
    return n_rays / dt, dt


def cpu_config3_rate(n_rays):
    """BASELINE config 3 on the host cores, the pattern of RN:168-181: per 512-ray chunk render (retraw) with autograd, then
    autograd.grad(rgb, batch_rays, grad_outputs=grad_E chunk)."""
    import torch
    rays, _ = _cpu_rays(n_rays)
    run = _cpu_renderer()
    g = torch.randn(n_rays, 3, generator=torch.Generator().manual_seed(0))
    t0 = time.perf_counter()
    for i in range(0, n_rays, 512):
        batch = rays[:, i:i + 512].clone().requires_grad_(True)
        rgb = run(batch, retraw=True)[0]
        torch.autograd.grad(rgb, batch, grad_outputs=g[i:i + 512])
    dt = time.perf_counter() - t0
    return n_rays / dt, dt


def eager_gpu_rates(dev):
    """The reference's own eager PyTorch path on THIS GPU -- the baseline SURVEY.md 2.1 names ("beat eager PyTorch"): RN.render with
    its NeRF modules on cuda (the kernels nn.Linear / sin / cumprod / searchsorted / sort dispatch to, fp32), or the oracle port
    with every tensor on cuda when the reference modules are not there.  TF32 off (what the reference gets by default) and on; the
    reference's chunk = 512 (CFG:25) on a 16 384-ray sample, and chunk = 32768 (nerf-pytorch's default) on the whole image."""
    import torch
    import nerf_oracle as O
    sdc, sdf = load_weights()
    kind = cpu_kind()
    out = {'kind': kind}
    full_image = None
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.device(dev), torch.no_grad():
            ro, rd = O.get_rays(H, W, O.YCBV_K_400, pose_for(0, 0).to(dev))
            if kind == 'reference':
                import ref_import
                RN, _ = ref_import.load()
                kw = ref_import.render_kwargs(sdc, sdf, O.YCBV_NEAR, O.YCBV_FAR, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE)
                kw['network_fn'].to(dev)
                kw['network_fine'].to(dev)
                render = lambda rays, chunk: RN.render(H, W, O.YCBV_K_400, chunk=chunk, rays=rays, **kw)
            else:
                sdc = {k: v.to(dev) for k, v in sdc.items()}
                sdf = {k: v.to(dev) for k, v in sdf.items()}
                render = lambda rays, chunk: O.render(H, W, O.YCBV_K_400, sdc, sdf, chunk=chunk, rays=rays, near=O.YCBV_NEAR, far=O.YCBV_FAR,
                                                      N_samples=N_SAMPLES, N_importance=N_IMPORTANCE)
            for tf32 in (False, True):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                for chunk, n_rays in ((512, 16384), (32768, RAYS_PER_IMAGE)):
                    sel = torch.linspace(0, RAYS_PER_IMAGE - 1, n_rays).long()
                    rays = torch.stack([ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel]], 0)
                    render(rays, chunk)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    res = render(rays, chunk)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    out[f'fp32_tf32_{"on" if tf32 else "off"}_chunk{chunk}'] = {'rays_per_s': n_rays / dt, 'rays': n_rays, 'seconds': dt}
                    if not tf32 and n_rays == RAYS_PER_IMAGE:
                        full_image = (rays.clone(), res[0].reshape(-1, 3).clone())      # the whole image in true fp32: the parity check's other arm
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out, full_image


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each MLP kernel, from the committed ncu --set full captures
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep files).  bench.py refuses to run without it:
    a traffic regression must show up as a changed file, not as a stale literal."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    with open(path) as f:
        return json.load(f), os.path.relpath(path, ROOT)


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    n = 2048
    import torch
    import ref_import
    kind = cpu_kind()
    with ref_import.cpu_shim():
        rays, run = _cpu_setup(n)
        threads = pick_cpu_threads(run, rays)
        with torch.no_grad():
            for _ in range(args.warmup):
                run(rays)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                run(rays)
            dt = time.perf_counter() - t0
    value = n * args.steps / dt
    what = 'the unmodified reference render() (RN:58-123, oracle/_ref bytecode build)' if kind == 'reference' else 'the oracle port of RN:58-123'
    sample = (f'{what}: {n} rays (uniform subsample of one 400x400 image) per step, chunk=512, netchunk=65536, fp32, torch CPU, '
              f'{threads} of {os.cpu_count()} threads (fastest of the candidates tried)')
    print(json.dumps({
        'impl': 'reference', 'device': 'cpu', 'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE config 2: 400x400 rays, 64 coarse + 128 fine, forward render (bounded sample per step)',
                   'rays_per_step': n, 'N_samples': N_SAMPLES, 'N_importance': N_IMPORTANCE},
        'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'gpus_used': 0,
    }))


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    import nerf_oracle as O
    import neural_sim_nerf_b200 as nsr

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    sdc, sdf = load_weights()
    nets = []
    for sd in (sdc, sdf):
        m = nsr.NeRF()
        m.load_state_dict(sd)
        nets.append(m.to(dev))
    L = nsr.lib()
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = RAYS_PER_IMAGE
    T = N_SAMPLES + N_IMPORTANCE
    n_steps_total = args.warmup + args.steps
    # per-step rays, resident in HBM before the timed region
    rays_dev = [nsr.make_rays(H, W, O.YCBV_K_400, pose_for(s, rank), O.YCBV_NEAR, O.YCBV_FAR) for s in range(min(n_steps_total, 8))]
    new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    outs = dict(rgb=new(n, 3), disp=new(n), acc=new(n), rgb0=new(n, 3), disp0=new(n), acc0=new(n), zstd=new(n))
    ws_bytes = L.nsr_render_workspace_bytes(n, N_SAMPLES, N_IMPORTANCE)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def step_device(s, flags=0):
        r = rays_dev[s % len(rays_dev)]
        rc = L.nsr_render_rays_forward(P(r), n, P(pc), P(pf), N_SAMPLES, N_IMPORTANCE, flags, None, None, P(outs['rgb']), P(outs['disp']),
                                       P(outs['acc']), P(outs['rgb0']), P(outs['disp0']), P(outs['acc0']), P(outs['zstd']), None,
                                       None, None, P(ws), ws_bytes, stream)
        if rc != 0:
            raise RuntimeError(L.nsr_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ------------------------------------------------------------- device-resident throughput
    for s in range(args.warmup):
        step_device(s)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.nsr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step_device(args.warmup + s)
    e1.record()
    barrier()
    launches = L.nsr_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    # two-tier evaluation: control words of the last step's coarse / fine pass (left in the workspace)
    off = (ctypes.c_size_t * 8)()
    L.nsr_render_workspace_layout(n, N_SAMPLES, N_IMPORTANCE, off, 8)
    import numpy as np
    c0, c1 = (ws[o:o + 16].view(torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF for o in (off[5], off[6]))
    as_float = lambda u: float(np.array([u], dtype=np.uint32).view(np.float32)[0])
    two_tier = {'coarse_active_fraction': float(c0[0]) / (n * N_SAMPLES), 'fine_active_fraction': float(c1[0]) / (n * T),
                'coarse_max_dsigma_on_active': as_float(c0[1]), 'fine_max_dsigma_on_active': as_float(c1[1]),
                'fine_pass_forced_dense': int(c1[2]), 'dense_re_evaluations': int(c0[3]) + int(c1[3]),
                'note': 'tier 1 (1 fp16 MMA per product, steps 0-7 + alpha head, every point) certifies sigma <= -4 => weight exactly 0; tier 2 = fp16x3 on the '
                        'active points only; maps bit-identical to the dense evaluation (tests/test_gpu_two_tier.py)'}
    value = world * n * args.steps / (ms_total * 1e-3)
    if args.profile:      # under ncu: device-resident steps only
        if world > 1:
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({'profile_run': True, 'value': value, 'ms_per_step': ms_total / args.steps, 'gpu_launches': int(launches)}))
        return

    # ------------------------------------------------------------- end to end through render(rays=...), host buffers
    kw = dict(network_fn=nets[0], network_query_fn=None, N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, network_fine=nets[1],
              use_viewdirs=True, ndc=False, near=O.YCBV_NEAR, far=O.YCBV_FAR, white_bkgd=False, raw_noise_std=0., perturb=False)
    host_rays = []
    for r in rays_dev:
        hr = torch.stack([r[:, 0:3], r[:, 3:6]], 0).cpu().pin_memory()            # [2, N, 3] as render_path_grad passes them (RN:163)
        host_rays.append(hr)
    host_out = [torch.empty(n, 3).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n).pin_memory()]
    h2d = host_rays[0].numel() * 4
    d2h = sum(t.numel() * 4 for t in host_out)

    def step_e2e(s):
        with torch.no_grad():
            r = host_rays[s % len(host_rays)].to(dev, non_blocking=True)
            rgb, disp, acc, _ = nsr.render(H, W, O.YCBV_K_400, chunk=1 << 20, rays=r, **kw)
            host_out[0].copy_(rgb, non_blocking=True)
            host_out[1].copy_(disp, non_blocking=True)
            host_out[2].copy_(acc, non_blocking=True)

    for s in range(args.warmup):
        step_e2e(s)
    barrier()
    e0.record()
    for s in range(args.steps):
        step_e2e(args.warmup + s)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * n * args.steps / (ms_e2e * 1e-3)

    # ------------------------------------------------------------- BASELINE configs 3 / 4: forward + backward to the pose, all ranks
    # per step and rank: one image forward (z / raw kept), nsr_render_rays_backward from a fixed dL/drgb, closed-form fold to dL/dc2w
    # [3,4], then the path's only collective: an all-reduce of the 12 pose-gradient floats (dist.py; MAIN:191's mean over images)
    T = N_SAMPLES + N_IMPORTANCE
    bws_bytes = L.nsr_render_backward_workspace_bytes(n, T)
    bws = torch.empty(bws_bytes, dtype=torch.uint8, device=dev)
    g_rgb = torch.randn(n, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    d_rays = new(n, 11)
    zsave, rawsave = new(n, T), new(n, T, 4)
    d_c2w = new(12)
    relu_mask = torch.empty(L.nsr_relu_mask_bytes(n, T), dtype=torch.uint8, device=dev)    # 272 B per ACTIVE sample point (room for all: 8.4 GB)
    aset = torch.empty(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device=dev)         # the last pass's active list: forward -> backward
    cws = torch.empty(L.nsr_c2w_grad_workspace_bytes(), dtype=torch.uint8, device=dev)
    Kf = (ctypes.c_float * 9)(*[float(v) for row in O.YCBV_K_400 for v in row])

    def step_pose_grad(s):
        r = rays_dev[s % len(rays_dev)]
        rc = L.nsr_render_rays_forward_ex(P(r), n, P(pc), P(pf), N_SAMPLES, N_IMPORTANCE, 0, None, None, P(outs['rgb']), P(outs['disp']),
                                          P(outs['acc']), P(outs['rgb0']), P(outs['disp0']), P(outs['acc0']), P(outs['zstd']), P(rawsave),
                                          P(zsave), None, P(relu_mask), None, P(aset), P(ws), ws_bytes, stream)
        rc = rc or L.nsr_render_rays_backward_ex(P(r), P(zsave), P(rawsave), n, T, P(pf), 0, P(g_rgb), P(d_rays), None, None, None,
                                                 P(relu_mask), P(aset), P(bws), bws_bytes, stream)
        rc = rc or L.nsr_rays_grad_to_c2w(H, W, Kf, P(r), P(d_rays), None, n, P(d_c2w), 0, P(cws), stream)
        if rc != 0:
            raise RuntimeError(L.nsr_last_error().decode())
        if world > 1:
            dist.all_reduce(d_c2w, op=dist.ReduceOp.SUM)

    for s in range(2):
        step_pose_grad(s)
    barrier()
    e0.record()
    for s in range(args.steps):
        step_pose_grad(args.warmup + s)
    e1.record()
    barrier()
    ms_pg = max_over_ranks(e0.elapsed_time(e1))
    pose_grad = {'workload': 'BASELINE config 3/4: per rank and step one 400x400 image forward (two-tier; one bit per ReLU of the active points saved, '
                             '272 B/point) + backward dL/d(rays) over the active set without recompute -> dL/dc2w (closed form) + all-reduce of the 12 '
                             'pose-gradient floats over the ranks',
                 'rays_per_s': world * n * args.steps / (ms_pg * 1e-3), 'ms_per_step': ms_pg / args.steps,
                 'collective': 'ncclAllReduce(SUM) of 48 bytes per step' if world > 1 else 'none (1 rank)',
                 'algorithmic_tflops': world * n * args.steps * (64 + 192 + 192) * FLOP_PER_POINT / (ms_pg * 1e-3) / 1e12}
    # ------------------------------------------------------------- BASELINE config 4 as written: 8 objects, each with its OWN pair of networks
    # (8 + 8 packed blobs resident on every GPU), the rays of every object's image sharded over the ranks, forward + backward to the
    # pose, ONE all-reduce of the [8, 12] pose gradients per step.  (No YCB-V checkpoints offline: the objects are the fitted scene with
    # per-object colour heads -- distinct weights, same geometry, so the active fractions stay those of the headline workload.)
    import copy
    n_obj = 8
    obj_blobs = []
    for o in range(n_obj):
        pair = []
        for m in nets:
            mo = copy.deepcopy(m)
            with torch.no_grad():
                gen_o = torch.Generator(device=dev).manual_seed(1000 + o)
                mo.rgb_linear.weight.add_(0.05 * torch.randn(mo.rgb_linear.weight.shape, device=dev, generator=gen_o))
                mo.rgb_linear.bias.add_(0.3 * torch.randn(mo.rgb_linear.bias.shape, device=dev, generator=gen_o))
            pair.append(nsr.packed_weights(mo).clone())
        obj_blobs.append(pair)
    assert len({b[1].data_ptr() for b in obj_blobs}) == n_obj and not torch.equal(obj_blobs[0][1], obj_blobs[1][1])
    # rays are dealt out round-robin (ray i -> rank i % world): a contiguous band through the object would carry twice the active
    # points of a band through the background, and the step ends with the slowest rank
    o_pix = torch.arange(rank, n, world, device=dev, dtype=torch.int32)
    ns = int(o_pix.numel())
    o_rays = [r_[o_pix.long()].contiguous() for r_ in rays_dev]
    o_grgb = g_rgb[o_pix.long()].contiguous()
    o_ws = torch.empty(L.nsr_render_workspace_bytes(ns, N_SAMPLES, N_IMPORTANCE), dtype=torch.uint8, device=dev)
    o_bws = torch.empty(L.nsr_render_backward_workspace_bytes(ns, T), dtype=torch.uint8, device=dev)
    o_mask = torch.empty(L.nsr_relu_mask_bytes(ns, T), dtype=torch.uint8, device=dev)
    o_aset = torch.empty(L.nsr_active_set_bytes(ns, T), dtype=torch.uint8, device=dev)
    o_z, o_raw, o_drays = new(ns, T), new(ns, T, 4), new(ns, 11)
    o_rgb = new(ns, 3)
    o_dc2w = torch.zeros(n_obj, 12, device=dev)

    def step_objects(s):
        for o in range(n_obj):
            r = o_rays[(s + o) % len(o_rays)]
            bc, bf = obj_blobs[o]
            rc = L.nsr_render_rays_forward_ex(P(r), ns, P(bc), P(bf), N_SAMPLES, N_IMPORTANCE, 0, None, None, P(o_rgb), None, None, None, None, None,
                                              None, P(o_raw), P(o_z), None, P(o_mask), None, P(o_aset), P(o_ws), o_ws.numel(), stream)
            rc = rc or L.nsr_render_rays_backward_ex(P(r), P(o_z), P(o_raw), ns, T, P(bf), 0, P(o_grgb), P(o_drays), None, None, None,
                                                     P(o_mask), P(o_aset), P(o_bws), o_bws.numel(), stream)
            rc = rc or L.nsr_rays_grad_to_c2w(H, W, Kf, P(r), P(o_drays), P(o_pix), ns, P(o_dc2w[o]), 0, P(cws), stream)
            if rc != 0:
                raise RuntimeError(L.nsr_last_error().decode())
        if world > 1:
            dist.all_reduce(o_dc2w, op=dist.ReduceOp.SUM)

    step_objects(0)
    barrier()
    o_steps = max(2, args.steps // 2)
    e0.record()
    for s in range(o_steps):
        step_objects(1 + s)
    e1.record()
    barrier()
    ms_obj = max_over_ranks(e0.elapsed_time(e1))
    objects8 = {'workload': f'BASELINE config 4: {n_obj} objects with their own coarse + fine networks ({2 * n_obj} packed blobs resident per GPU), one 400x400 image '
                            f'each per step, every image\'s rays dealt round-robin over the {world} rank(s); forward (two-tier, sign bits saved) + backward over the '
                            'active set -> dL/dc2w per object; one all-reduce of the [8,12] pose gradients per step',
                'rays_per_s': n_obj * n * o_steps / (ms_obj * 1e-3), 'ms_per_step': ms_obj / o_steps, 'rays_per_rank_per_step': n_obj * ns,
                'collective': 'ncclAllReduce(SUM) of 384 bytes per step' if world > 1 else 'none (1 rank)',
                'finite': bool(torch.isfinite(o_dc2w).all()), 'distinct_gradients': bool(len({round(float(v), 9) for v in o_dc2w[:, 0]}) == n_obj)}
    del o_ws, o_bws, o_mask, o_aset, o_z, o_raw, o_drays, obj_blobs, o_rays
    mlp_bwd_ms = None
    if rank == 0:       # the MLP stage of that backward pass by itself (for roofline_kernels): d_raw is in the backward workspace
        d_raw_view = bws[:n * T * 16].view(torch.float32)
        d_pts_tmp = new(n, T, 8)
        for _ in range(2):
            L.nsr_mlp_backward(P(rays_dev[(args.warmup + args.steps - 1) % len(rays_dev)]), P(zsave), n, T, P(pf), P(d_raw_view), P(d_pts_tmp), P(relu_mask), P(aset), stream)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            L.nsr_mlp_backward(P(rays_dev[(args.warmup + args.steps - 1) % len(rays_dev)]), P(zsave), n, T, P(pf), P(d_raw_view), P(d_pts_tmp), P(relu_mask), P(aset), stream)
        e1.record()
        torch.cuda.synchronize()
        mlp_bwd_ms = e0.elapsed_time(e1) / 3
        bwd_active = int(aset[:4].view(torch.int32).item())
        del d_pts_tmp
    del bws, zsave, rawsave, d_rays, relu_mask, aset

    # ------------------------------------------------------------- rooflines of the MLP kernels, each timed back to back
    roofline = cpu_base = fast = mixed = fwd_bwd = stages = train = parity = baselines = None
    roofline_kernels = roofline_step = None
    if rank == 0:
        zf = new(n, T)
        rawf = new(n, T, 4)
        # the fine pass's real depths (z_vals_out) of pose 0 and its real active list
        rc = L.nsr_render_rays_forward(P(rays_dev[0]), n, P(pc), P(pf), N_SAMPLES, N_IMPORTANCE, 0, None, None, P(outs['rgb']), P(outs['disp']),
                                       P(outs['acc']), P(outs['rgb0']), P(outs['disp0']), P(outs['acc0']), P(outs['zstd']), None, P(zf), None,
                                       P(ws), ws_bytes, stream)
        assert rc == 0, L.nsr_last_error()
        as1 = torch.zeros(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device=dev)
        z0, w0 = new(n, N_SAMPLES), new(n, N_SAMPLES)
        t = torch.linspace(0, 1, N_SAMPLES, device=dev)
        z0.copy_(O.YCBV_NEAR * (1 - t) + O.YCBV_FAR * t)
        w0.uniform_(0, 1)
        reps = max(3, args.steps)

        def time_kernel(fn):
            for _ in range(2):
                assert fn() == 0, L.nsr_last_error()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        peaks, how = measured_peaks()
        peak = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops'])
        traffic, traffic_file = ncu_traffic()
        t1_ms = time_kernel(lambda: L.nsr_mlp_two_tier(P(rays_dev[0]), P(zf), n, T, P(pf), P(rawf), P(as1), None, 1, stream))
        n_act = int(as1[:4].view(torch.int32).item())
        t2_ms = time_kernel(lambda: L.nsr_mlp_two_tier(P(rays_dev[0]), P(zf), n, T, P(pf), P(rawf), P(as1), None, 2, stream))
        k_ms = time_kernel(lambda: L.nsr_mlp_forward(P(rays_dev[0]), P(zf), n, T, P(pf), 0, P(rawf), stream))
        flops = n * T * FLOP_PER_POINT

        def entry(key, kernel, ms, alg_flop, mma_per_product, note):
            ach = alg_flop / (ms * 1e-3) / 1e12
            t_ = traffic.get(key)
            return {'bound': 'tensor', 'kernel': kernel, 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                    'traffic': None if t_ is None else t_['dram_bytes'], 'traffic_source': None if t_ is None else f"{traffic_file}: {t_['source']}",
                    'peak_source': f'{how} bf16_tflops_sustained (kernel timed back to back)', 'ms_per_launch': ms,
                    'algorithmic_flop_per_launch': alg_flop, 'tensor_flop_issued_per_algorithmic_flop': mma_per_product, 'note': note}

        roofline_kernels = [
            entry('fine_tier1', 'nerf_mlp_kernel<1,false,true> (fine pass, tier 1: pts_linears.0-7 + alpha head of all 160000 x 192 points, fp16)',
                  t1_ms, n * T * FLOP_PER_POINT_TIER1, 1, 'algorithmic = the FLOPs of the layers this kernel evaluates (RH:99-109), every point'),
            entry('fine_tier2', f'nerf_mlp_kernel<3,false,false> (fine pass, tier 2: the whole network on the {n_act} active points, fp16x3)',
                  t2_ms, n_act * FLOP_PER_POINT, 3, 'algorithmic = 1 186 816 FLOP x the active points only'),
            entry('fine_dense', 'nerf_mlp_kernel<3,false,false> (fine pass evaluated densely, fp16x3: NSR_FLAG_DENSE / round 1)', k_ms, flops, 3,
                  'every point through the error-compensated arithmetic'),
        ]
        if mlp_bwd_ms is not None:
            roofline_kernels.append(entry('bwd_masked_active', f'nerf_mlp_bwd_kernel<true> (data gradient from saved ReLU bits, {bwd_active} active points, fp16x3)',
                                          mlp_bwd_ms, bwd_active * FLOP_PER_POINT, 3, 'algorithmic = the transposed network, 1 186 816 FLOP per active point'))
        roofline = dict(max(roofline_kernels[:2], key=lambda r: r['ms_per_launch']))      # the longer of the two launches that make the fine pass
        step_ach = n * FLOP_PER_RAY / (ms_total / args.steps * 1e-3) / 1e12
        roofline_step = {'what': 'the whole forward step in FLOPs of the reference algorithm (64 + 64 + 128... = 256 point evaluations per ray x 1 186 816), '
                                 'including the views branch the empty points never run here', 'achieved': step_ach, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': step_ach / peak}
        achieved = roofline['achieved']
        if world == 1:          # side legs only at N=1 (the scaling runs stay short; cpu_baseline is an N=1 figure)
            # every point through fp16x3 (NSR_FLAG_DENSE): what round 1 measured as the default
            DENSE = 32
            for _ in range(2):
                step_device(0, DENSE)
            torch.cuda.synchronize()
            e0.record()
            for s in range(args.steps):
                step_device(s, DENSE)
            e1.record()
            torch.cuda.synchronize()
            dense_ms = e0.elapsed_time(e1) / args.steps
            two_tier['dense_fp16x3'] = {'rays_per_s_device_resident': n / (dense_ms * 1e-3), 'ms_per_step': dense_ms}
            # informational: the opt-in single-pass fp16 mode (NOT parity-valid, see DESIGN.md "precision")
            FAST = 8
            for _ in range(2):
                step_device(0, FAST)
            torch.cuda.synchronize()
            e0.record()
            for s in range(args.steps):
                step_device(s, FAST)
            e1.record()
            torch.cuda.synchronize()
            fast_ms = e0.elapsed_time(e1) / args.steps
            L.nsr_mlp_forward(P(rays_dev[0]), P(zf), n, T, P(pf), FAST, P(rawf), stream)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                L.nsr_mlp_forward(P(rays_dev[0]), P(zf), n, T, P(pf), FAST, P(rawf), stream)
            e1.record()
            torch.cuda.synchronize()
            fk_ms = e0.elapsed_time(e1) / reps
            fast = {'note': 'NSR_FLAG_FAST_FP16: one fp16 MMA per product; misses the 1e-3 parity bar on silhouette rays -- informational only',
                    'rays_per_s_device_resident': n / (fast_ms * 1e-3), 'ms_per_step': fast_ms,
                    'fine_mlp_ms_per_launch': fk_ms, 'fine_mlp_tflops': flops / (fk_ms * 1e-3) / 1e12,
                    'fine_mlp_frac_of_peak': flops / (fk_ms * 1e-3) / 1e12 / peak}
            # informational: the opt-in mixed mode (fp16 + e4m3 residual products; holds 1e-3 on the fitted scene, not on every network)
            MIXED = 16
            for _ in range(2):
                step_device(0, MIXED)
            torch.cuda.synchronize()
            e0.record()
            for s in range(args.steps):
                step_device(s, MIXED)
            e1.record()
            torch.cuda.synchronize()
            mixed_ms = e0.elapsed_time(e1) / args.steps
            mixed = {'note': 'NSR_FLAG_MIXED_F8: 2.25 tensor passes per product (fp16 main term + two e4m3 residual products in layers 3-9); opt-in, informational only',
                     'rays_per_s_device_resident': n / (mixed_ms * 1e-3), 'ms_per_step': mixed_ms}
            # secondary (BASELINE config 3): forward + backward dL/d(rays) for the pose path, same 160 000 rays
            bws_bytes = L.nsr_render_backward_workspace_bytes(n, T)
            bws = torch.empty(bws_bytes, dtype=torch.uint8, device=dev)
            g_rgb = torch.randn(n, 3, device=dev)
            d_rays = new(n, 11)
            zsave, rawsave = new(n, T), new(n, T, 4)
            aset_fb = torch.empty(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device=dev)

            def step_fwd_bwd(s):
                r = rays_dev[s % len(rays_dev)]
                rc = L.nsr_render_rays_forward_ex(P(r), n, P(pc), P(pf), N_SAMPLES, N_IMPORTANCE, 0, None, None, P(outs['rgb']), P(outs['disp']),
                                                  P(outs['acc']), P(outs['rgb0']), P(outs['disp0']), P(outs['acc0']), P(outs['zstd']), P(rawsave),
                                                  P(zsave), None, None, None, P(aset_fb), P(ws), ws_bytes, stream)
                rc = rc or L.nsr_render_rays_backward_ex(P(r), P(zsave), P(rawsave), n, T, P(pf), 0, P(g_rgb), P(d_rays), None, None, None,
                                                         None, P(aset_fb), P(bws), bws_bytes, stream)
                if rc != 0:
                    raise RuntimeError(L.nsr_last_error().decode())

            for s in range(2):
                step_fwd_bwd(s)
            torch.cuda.synchronize()
            e0.record()
            for s in range(args.steps):
                step_fwd_bwd(s)
            e1.record()
            torch.cuda.synchronize()
            fb_ms = e0.elapsed_time(e1) / args.steps
            fwd_bwd = {'workload': 'BASELINE config 3 on the recompute route (4 B per sample point of extra memory, the active list): forward (two-tier) + backward dL/d(rays) from dL/d(rgb_map), 160000 rays, 64+128 samples; the backward kernel recomputes the fine pass on the active points',
                       'rays_per_s': n / (fb_ms * 1e-3), 'ms_per_step': fb_ms, 'algorithmic_flop_per_ray': (64 + 192 + 192) * FLOP_PER_POINT,
                       'algorithmic_tflops': n * (64 + 192 + 192) * FLOP_PER_POINT / (fb_ms * 1e-3) / 1e12}
            del bws, zsave, rawsave, aset_fb
            # secondary (SURVEY a-12 / RN:643-716): one optimisation step of both networks on N_rand = 1024 random rays through the
            # public API: render(rays=...) -> img2mse(rgb) + img2mse(rgb0) -> backward (dL/dMLP of both nets) -> Adam -> re-pack
            import copy
            tnets = [copy.deepcopy(m) for m in nets]
            opt = torch.optim.Adam([p_ for m in tnets for p_ in m.parameters()], lr=5e-4, betas=(0.9, 0.999))
            tkw = dict(kw, network_fn=tnets[0], network_fine=tnets[1], perturb=1.0)      # stratified sampling as in training (RN:447-461)
            n_rand = 1024
            gen = torch.Generator(device=dev).manual_seed(0)
            target = torch.rand(n_rand, 3, device=dev, generator=gen)

            def train_step(s):
                sel = torch.randint(0, n, (n_rand,), device=dev, generator=gen)
                r = rays_dev[s % len(rays_dev)][sel]
                batch = torch.stack([r[:, 0:3], r[:, 3:6]], 0)
                rgb, _, _, extras = nsr.render(H, W, O.YCBV_K_400, chunk=1 << 15, rays=batch, retraw=True, **tkw)
                opt.zero_grad()
                loss = nsr.img2mse(rgb, target) + nsr.img2mse(extras['rgb0'], target)
                loss.backward()
                opt.step()

            launches_t0 = L.nsr_launch_count()
            for s in range(3):
                train_step(s)
            torch.cuda.synchronize()
            t_steps = max(10, args.steps)
            e0.record()
            for s in range(t_steps):
                train_step(s)
            e1.record()
            torch.cuda.synchronize()
            tr_ms = e0.elapsed_time(e1) / t_steps
            train = {'workload': 'one Adam step of both networks on N_rand=1024 rays, 64+128 samples, loss = mse(rgb) + mse(rgb0) (RN:643-716) through render() + autograd + torch.optim.Adam',
                     'ms_per_step': tr_ms, 'rays_per_s': n_rand / (tr_ms * 1e-3), 'steps_per_s': 1e3 / tr_ms,
                     'kernels_per_step': (L.nsr_launch_count() - launches_t0) / (t_steps + 3)}
            # the same iteration as ONE C call (nsr_train_step: forward, loss, both backward passes, Adam, re-pack; device Philox draws)
            fnets = [copy.deepcopy(m) for m in nets]
            fopt = torch.optim.Adam([p_ for m in fnets for p_ in m.parameters()], lr=5e-4, betas=(0.9, 0.999))
            fkw = dict(tkw, network_fn=fnets[0], network_fine=fnets[1])

            def fused_step(s):
                sel = torch.randint(0, n, (n_rand,), device=dev, generator=gen)
                r = rays_dev[s % len(rays_dev)][sel]
                return nsr.train_step(torch.stack([r[:, 0:3], r[:, 3:6]], 0), target, fopt, **fkw)

            launches_f0 = L.nsr_launch_count()
            for s in range(3):
                fused_step(s)
            torch.cuda.synchronize()
            e0.record()
            for s in range(t_steps):
                fused_step(s)
            e1.record()
            torch.cuda.synchronize()
            fu_ms = e0.elapsed_time(e1) / t_steps
            train['fused'] = {'api': 'train_step(batch_rays, target_s, optimizer, **render_kwargs_train) -> nsr_train_step', 'ms_per_step': fu_ms,
                              'rays_per_s': n_rand / (fu_ms * 1e-3), 'steps_per_s': 1e3 / fu_ms,
                              'kernels_per_step': (L.nsr_launch_count() - launches_f0) / (t_steps + 3)}
            del tnets, opt, fnets, fopt
            # per-stage table: the HBM-bound ray-stage kernels, each timed alone (CUDA events, 20 launches back to back)
            hbm = peaks.get('hbm_gbs', 6650.0)

            def time_stage(fn, reps=20):
                fn()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            S_, T_ = N_SAMPLES, T
            raw0 = new(n, S_, 4).normal_()
            wts = new(n, T_)
            rgb8 = torch.empty(n * 3, dtype=torch.uint8, device=dev)
            d3 = rays_dev[0][:, 3:6].contiguous()
            Kh = (ctypes.c_float * 9)(*[float(v) for row in O.YCBV_K_400 for v in row])
            c2wh = (ctypes.c_float * 12)(*[float(v) for v in pose_for(0, 0).reshape(-1).tolist()])
            rays_tmp = new(n, 11)
            stage_defs = [
                ('make_rays_kernel (RH:156-165 + RN:91-112)', n * 44,
                 lambda: L.nsr_make_rays(H, W, Kh, c2wh, float(O.YCBV_NEAR), float(O.YCBV_FAR), P(rays_tmp), stream)),
                ('raw2outputs_kernel S=64 (coarse composite, writes weights)', n * (S_ * 16 + S_ * 4 + 12 + S_ * 4 + 20),
                 lambda: L.nsr_raw2outputs(P(raw0), P(z0), P(d3), 3, n, S_, 0, P(outs['rgb0']), P(outs['disp0']), P(outs['acc0']), P(w0), None, stream)),
                ('resample_merge_kernel 64+128 (sample_pdf + sort + z_std)', n * (S_ * 4 + S_ * 4 + T_ * 4 + 4),
                 lambda: L.nsr_resample_merge(P(z0), P(w0), n, S_, N_IMPORTANCE, None, P(zf), None, P(outs['zstd']), stream)),
                ('raw2outputs_kernel S=192 (fine composite)', n * (T_ * 16 + T_ * 4 + 12 + 20),
                 lambda: L.nsr_raw2outputs(P(rawf), P(zf), P(d3), 3, n, T_, 0, P(outs['rgb']), P(outs['disp']), P(outs['acc']), None, None, stream)),
                ('to8b_kernel (RH:14)', n * 15,
                 lambda: L.nsr_to8b(P(outs['rgb']), n * 3, P(rgb8), stream)),
            ]
            stages = []
            ncu_key = {'raw2outputs_kernel S=64': 'composite_coarse', 'resample_merge_kernel': 'resample_merge', 'raw2outputs_kernel S=192': 'composite_fine'}
            for name, nbytes, fn in stage_defs:
                ms = time_stage(fn)
                st_ = {'kernel': name, 'ms': ms, 'algorithmic_bytes': nbytes, 'gb_per_s': nbytes / (ms * 1e-3) / 1e9,
                       'frac_of_hbm_peak': nbytes / (ms * 1e-3) / 1e9 / hbm}
                for prefix, key in ncu_key.items():          # ncu-reported DRAM bytes of the same kernel in a real render (profiles/ncu_traffic.json)
                    if name.startswith(prefix) and key in traffic:
                        t_ = traffic[key]
                        st_.update({'dram_bytes_ncu': t_['dram_bytes'], 'ncu_duration': t_['duration_under_ncu'],
                                    'gb_per_s_ncu_bytes': t_['dram_bytes'] / (ms * 1e-3) / 1e9, 'frac_of_hbm_peak_ncu_bytes': t_['dram_bytes'] / (ms * 1e-3) / 1e9 / hbm})
                stages.append(st_)
            stages.append({'kernel': 'nerf_mlp_kernel fine, dense fp16x3 (192 samples/ray)', 'ms': k_ms, 'algorithmic_tflops': flops / (k_ms * 1e-3) / 1e12,
                           'frac_of_tensor_peak': flops / (k_ms * 1e-3) / 1e12 / peak})
            # CPU baseline: the oracle port on this box's host cores, bounded sample
            kind = cpu_kind()
            rate, cores, times, cpu_out = cpu_render_rate(4096, 2)
            what = 'the unmodified reference render() (RN:58-123; oracle/_ref bytecode build of /root/reference)' if kind == 'reference' else 'the oracle port of RN:58-123'
            cpu_base = {'value': rate, 'unit': 'rays/s', 'cores': cores, 'kind': kind,
                        'sample': f'{what}: 2 passes over 4096 rays of the same image (chunk=512, netchunk=65536, fp32 torch CPU, {cores} of {os.cpu_count()} threads = fastest tried, {sum(times):.1f} s)'}
            # parity of the timed workload: the image of pose 0 as the GPU path renders it against those CPU-rendered rays
            step_device(0)
            torch.cuda.synchronize()
            sel = torch.linspace(0, RAYS_PER_IMAGE - 1, 4096).long()
            errs = {}
            for key, ref_t in (('rgb', cpu_out[0]), ('acc', cpu_out[2]), ('rgb0', cpu_out[3]['rgb0']), ('acc0', cpu_out[3]['acc0'])):
                got = outs[key][sel.to(dev)].cpu()
                errs[key] = float(((got - ref_t).abs() / ref_t.abs().clamp(min=1.0)).max())
            gd, rd_ = outs['disp'][sel.to(dev)].cpu(), cpu_out[1]
            nan_equal = bool((torch.isnan(gd) == torch.isnan(rd_)).all())
            ok_ = ~torch.isnan(rd_) & (cpu_out[2] > 1e-2)          # disparity = 1 / (depth / acc): ill-conditioned as acc -> 0
            errs['disp(acc>0.01)'] = float(((gd[ok_] - rd_[ok_]).abs() / rd_[ok_].abs().clamp(min=1.0)).max()) if bool(ok_.any()) else 0.0
            parity = {'max_rel_err': max(errs.values()), 'per_output': errs, 'n_rays': 4096, 'nan_equal': nan_equal, 'tolerance': 1e-3,
                      'against': kind, 'rel_err': '|a - b| / max(1, |b|)', 'ok': bool(max(errs.values()) <= 1e-3 and nan_equal)}
            # the other BASELINE configs on the host cores, and the reference's eager PyTorch path on this GPU
            import ref_import
            with ref_import.cpu_shim():
                torch.set_num_threads(cores)
                c1_rate, c1_s = cpu_config1_rate(4096)
                c3_rate, c3_s = cpu_config3_rate(1024)
            # config 1 on the GPU path: 200x200, 64 coarse samples only, through render(c2w=...)
            kw1 = dict(kw, N_importance=0, network_fine=None)
            c2w1 = pose_for(0, 0).to(dev)

            def gpu_config1():
                with torch.no_grad():
                    return nsr.render(200, 200, K_200, chunk=1 << 20, c2w=c2w1, **kw1)

            c1_gpu_ms = time_stage(gpu_config1, reps=10)
            baselines = {
                'kind': kind, 'cores': cores,
                'config1_cpu': {'workload': 'BASELINE config 1: 200x200 camera, 64 coarse samples only (N_importance = 0), forward, chunk 512; 4096-ray sample',
                                'rays_per_s': c1_rate, 'seconds': c1_s},
                'config1_gpu': {'workload': 'the same config through render(c2w=...) on this GPU, whole 200x200 image', 'rays_per_s': 40000 / (c1_gpu_ms * 1e-3),
                                'ms_per_image': c1_gpu_ms},
                'config3_cpu': {'workload': 'BASELINE config 3, the pattern of RN:168-181: per 512-ray chunk render + autograd.grad(rgb, batch_rays); 1024-ray sample',
                                'rays_per_s': c3_rate, 'seconds': c3_s, 'gpu_counterpart': 'pose_grad / fwd_bwd'},
            }
            baselines['eager_pytorch_on_this_gpu'], (eager_rays, eager_rgb) = eager_gpu_rates(dev)
            # every ray of that image: this path on the SAME rays against the reference's eager fp32 pixels.  Hierarchical sampling
            # is ill-conditioned on rays that graze the object: a handful of rays per image is decided by fp32 rounding
            # in the reference itself; the run fails if more than 8 of the 160 000 are beyond 1e-3.
            with torch.no_grad():
                mine = nsr.render(H, W, O.YCBV_K_400, chunk=1 << 20, rays=eager_rays, **kw)[0].reshape(-1, 3)
            dfull = (mine - eager_rgb).abs().max(-1).values
            parity['full_image'] = {'against': f'{kind} eager fp32 on this GPU, same rays', 'rays': int(dfull.numel()), 'beyond_1e-3': int((dfull > 1e-3).sum()),
                                    'beyond_1e-4': int((dfull > 1e-4).sum()), 'max_abs_err': float(dfull.max()), 'median_abs_err': float(dfull.median()),
                                    'allowed_beyond_1e-3': 8}
            parity['ok'] = bool(parity['ok'] and parity['full_image']['beyond_1e-3'] <= 8)
            del eager_rays, eager_rgb, mine

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f16 / f16x3: tier 1 (every point, density only) one fp16 tcgen05 MMA per product; tier 2 (active points) fp16 hi/lo-split operands '
                     '(x_hi.W_hi + x_lo.W_hi + x_hi.W_lo); f32 accumulate; f32 everywhere else',
            'data': 'synthetic',
            'config': {'workload': 'BASELINE config 2: one 400x400 image (160000 rays) per GPU per step, 64 coarse + 128 fine, forward render',
                       'rays_per_step_per_gpu': n, 'N_samples': N_SAMPLES, 'N_importance': N_IMPORTANCE, 'parallelism': f'dp{world} (rays sharded by image, no forward collective)',
                       'weights': 'tests/golden/wfit.npz (analytic-scene fit; no pretrained checkpoint offline)',
                       'l2': 'per-step intermediates 0.86 GB >> 126 MB L2; rays rotate over 8 poses'},
            'e2e': {'value': e2e_value, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps,
                    'api': 'render(H, W, K, chunk, rays=<pinned host [2,N,3] -> cuda>, **render_kwargs_test) + D2H of rgb/disp/acc'},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'roofline_kernels': roofline_kernels, 'roofline_step': roofline_step,
            'cpu_baseline': cpu_base, 'parity': parity, 'baselines': baselines,
            'flop_per_ray': FLOP_PER_RAY, 'tflops_device_resident': value * FLOP_PER_RAY / 1e12,
            'two_tier': two_tier, 'fast_fp16_mode': fast, 'mixed_f8_mode': mixed, 'fwd_bwd': fwd_bwd, 'pose_grad': pose_grad, 'objects8': objects8, 'train_step': train, 'stages': stages,
        }))
        if parity is not None and not parity['ok']:
            sys.stderr.write(f'bench.py: PARITY FAILED on the timed image: {parity}\n')
            sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--profile', action='store_true', help='device-resident steps only (for runs under ncu)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
