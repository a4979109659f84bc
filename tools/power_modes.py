"""Fine-pass MLP launch back to back for a few seconds per precision mode, with nvidia-smi sampling beside it:
ms / launch, median SM clock, mean power, throttle reasons.  GPU box only."""
import ctypes, os, subprocess, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_sim_nerf_b200 as nsr
from neural_sim_nerf_b200._lib import ptr
SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
sd = {k[len('fine/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
net = nsr.NeRF(); net.load_state_dict(sd); net = net.cuda()
blob = nsr.packed_weights(net)
n, T = 160000, 192
g = torch.Generator(device='cuda').manual_seed(0)
rays = torch.zeros(n, 11, device='cuda')
rays[:, 0:3] = torch.randn(n, 3, device='cuda', generator=g) * 0.3
d = torch.randn(n, 3, device='cuda', generator=g); d = d / d.norm(dim=-1, keepdim=True)
rays[:, 3:6] = d; rays[:, 8:11] = d
zv = torch.sort(torch.rand(n, T, device='cuda', generator=g) * 1.6 + 0.3, -1).values.contiguous()
raw = torch.empty(n, T, 4, device='cuda')
L = nsr.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, flag in (('fp16x3', 0), ('mixed', 16), ('fp16', 8), ('mixed', 16), ('fp16x3', 0)):
    for _ in range(2):
        L.nsr_mlp_forward(ptr(rays), ptr(zv), n, T, ptr(blob), flag, ptr(raw), st)
    torch.cuda.synchronize()
    smi = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active,temperature.gpu', '--format=csv,noheader,nounits', '-lms', '100'],
                           stdout=subprocess.PIPE, text=True)
    t0 = time.time(); launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < SECONDS:
        for _ in range(4):
            L.nsr_mlp_forward(ptr(rays), ptr(zv), n, T, ptr(blob), flag, ptr(raw), st)
        launches += 4
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    smi.terminate()
    rows = [r.split(',') for r in smi.stdout.read().strip().splitlines()]
    rows = [r for r in rows if len(r) >= 3][3:]
    clk = np.median([float(r[0]) for r in rows]); pw = np.mean([float(r[1]) for r in rows])
    reasons = sorted({r[2].strip() for r in rows}); temp = max(float(r[3]) for r in rows)
    ms = e0.elapsed_time(e1) / launches
    print(f'{name:8s} {ms:8.3f} ms/launch  {n * T * 1186816 / ms / 1e9:7.1f} alg TFLOP/s  clock {clk:6.0f} MHz  power {pw:6.0f} W  temp {temp:.0f}C  reasons {reasons}  ({len(rows)} samples)', flush=True)
