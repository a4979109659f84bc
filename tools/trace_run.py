"""Debug: dump CTA 0's per-step timeline of the MLP kernel (NSR_TRACE_FILE) for the precision modes given as flags
on the command line (default: 0 = fp16x3, 16 = mixed), then print per-step intervals.
slots: MMA warp 0 step start | 1 a_ready[0] seen | 2 before / 3 after the a_ready[1] wait | 4 step issued
       epilogue 8 acc_ready[0] seen | 9 ACC0 converted | 11 first half stored | 10 acc_ready[1] seen | 12 second half stored"""
import ctypes, os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
OUT = 'gpurun_out/trace.txt'
os.environ['NSR_TRACE_FILE'] = OUT
if os.path.exists(OUT):
    os.remove(OUT)
import numpy as np, torch
import neural_sim_nerf_b200 as nsr, nerf_oracle as O
z = np.load('tests/golden/wfit.npz')
sd = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}
net = nsr.NeRF(); net.load_state_dict(sd); net.cuda()
pf = nsr.packed_weights(net)
L = nsr.lib(); n, T = 160000, 192
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
zf = torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, T, device='cuda').expand(n, T).contiguous()
raw = torch.empty(n, T, 4, device='cuda')
P = lambda t: ctypes.c_void_p(t.data_ptr())
modes = [int(a) for a in sys.argv[1:]] or [0, 16]
for flags in modes + modes:
    L.nsr_mlp_forward(P(rays), P(zf), n, T, P(pf), flags, P(raw), None)
torch.cuda.synchronize()
# ---- summarise the LAST launch of each mode: tile 2 (steady state)
blocks, cur = [], None
for line in open(OUT):
    if line.startswith('#'):
        cur = {'hdr': line.strip(), 'rows': []}
        blocks.append(cur)
    else:
        cur['rows'].append([int(x) for x in line.split()])
for b in blocks[len(modes):]:
    print(b['hdr'])
    print('step | mma: wait a0, issue k01, wait a1, issue rest | epi: acc0 after start, cvt0, wait acc1, store0, drain1 | step len | weight-wait cycles, waits')
    rows = [r for r in b['rows'] if r[0] == 2]
    for i, r in enumerate(rows):
        s = r[2:]
        nxt = rows[i + 1][2] if i + 1 < len(rows) else 0
        f = lambda a, b_: (s[b_] - s[a]) if (s[a] and s[b_]) else -1
        print(f'{r[1]:4d} | {f(0,1):6d} {f(1,2):6d} {f(2,3):6d} {f(3,4):6d} | {f(0,8):6d} {f(8,9):6d} {f(9,10):6d} {f(10,11):6d} {f(11,12):6d} | {nxt - s[0] if nxt else -1:6d} | {s[5]:6d} {s[6]:3d}')
    t0 = b['rows'][10][2]; t3 = b['rows'][30][2]
    print('cycles per tile (tiles 1..2 average):', (t3 - t0) / 2)
