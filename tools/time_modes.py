"""Times the fine-pass MLP launch (160 000 rays x 192 samples) in each precision mode.  GPU box only."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_sim_nerf_b200 as nsr
from neural_sim_nerf_b200._lib import ptr
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
outs = {}
for name, flag in (('fp16x3', 0), ('mixed', 16), ('fp16', 8), ('mixed', 16), ('fp16x3', 0)):
    for _ in range(2):
        assert L.nsr_mlp_forward(ptr(rays), ptr(zv), n, T, ptr(blob), flag, ptr(raw), st) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 8
    e0.record()
    for _ in range(reps):
        L.nsr_mlp_forward(ptr(rays), ptr(zv), n, T, ptr(blob), flag, ptr(raw), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    outs[name] = raw.clone()
    print(f'{name:8s} {ms:8.3f} ms / launch   {n * T * 1186816 / ms / 1e9:8.1f} algorithmic TFLOP/s', flush=True)
ref = outs['fp16x3']
for k in ('mixed', 'fp16'):
    e = (outs[k] - ref).abs() / ref.abs().clamp(min=1.0)
    print(f'{k} vs fp16x3 raw: max rel {e.max().item():.3e}  mean {e.mean().item():.3e}  nan {int(torch.isnan(outs[k]).sum())}')
