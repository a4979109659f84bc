"""Debug-hook builds only (NSR_LIB_PATH=build/libnsr_dbg.so): CTA 0's timeline of the tier-1 kernel, steady-state tile 2.
MMA warp slots: 0 step start, 1 a_ready[0] seen, 5/6 before/after the acc_free1 wait, 2/3 before/after the a_ready[1] wait,
7 accumulator 0's last chunk issued, 4 step issued.  Epilogue group g (slots 8+4g ..): accumulator seen, in registers, converted, stored + arrived."""
import ctypes, os, sys
OUT = 'gpurun_out/trace_tier1.txt'
os.environ['NSR_TRACE_FILE'] = OUT
if os.path.exists(OUT):
    os.remove(OUT)
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import nerf_oracle as O
import neural_sim_nerf_b200 as nsr
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
m = nsr.NeRF(); m.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('fine/')}); m.cuda()
pf = nsr.packed_weights(m)
L = nsr.lib(); P = lambda t: ctypes.c_void_p(t.data_ptr())
n, T = 160000, 192
rays = nsr.make_rays(400, 400, O.YCBV_K_400, O.pose_spherical(90., 22.5 - 180., 1.01)[:3, :4], O.YCBV_NEAR, O.YCBV_FAR)
zf = torch.linspace(O.YCBV_NEAR, O.YCBV_FAR, T, device='cuda').expand(n, T).contiguous()
raw = torch.empty(n, T, 4, device='cuda')
aset = torch.zeros(L.nsr_active_set_bytes(n, T), dtype=torch.uint8, device='cuda')
for _ in range(2):
    assert L.nsr_mlp_two_tier(P(rays), P(zf), n, T, P(pf), P(raw), P(aset), None, 1, None) == 0
torch.cuda.synchronize()
blocks, cur = [], None
for line in open(OUT):
    if line.startswith('#'):
        cur = {'hdr': line.strip(), 'rows': []}
        blocks.append(cur)
    else:
        cur['rows'].append([int(x) for x in line.split()])
b = blocks[-1]
print(b['hdr'])
names = ['start', 'a0', 'b.a1', 'a.a1', 'issued', 'b.f1', 'a.f1', 'acc0 cm', 'g0 seen', 'g0 regs', 'g0 conv', 'g0 arr', 'g1 seen', 'g1 regs', 'g1 conv', 'g1 arr']
order = [0, 1, 5, 6, 2, 3, 7, 4, 8, 9, 10, 11, 12, 13, 14, 15]
print('step | ' + ' '.join(f'{names[i]:>8s}' for i in order) + ' | step len')
for tile in (1, 2):
    rows = [r for r in b['rows'] if r[0] == tile]
    for i, r in enumerate(rows[:8]):
        s = r[2:]
        nxt = rows[i + 1][2] if i + 1 < 8 else [q for q in b['rows'] if q[0] == tile + 1][0][2]
        print(f'{r[1]:4d} | ' + ' '.join(f'{(s[j] - s[0]) if s[j] else -1:8d}' for j in order) + f' | {nxt - s[0]:8d}')
    print()
t0 = [r for r in b['rows'] if r[0] == 1][0][2]
t3 = [r for r in b['rows'] if r[0] == 3][0][2]
print('cycles per tile (tiles 1..2 average):', (t3 - t0) / 2)
