"""Diagnostic: where does the dL/dMLP error of the fine pass come from?  Decodes the weight-gradient operand dump
(common.cuh "backward dump") and compares every array, and the gradients formed from it, with an fp64 torch evaluation of the
same network on the same sample positions.  GPU only; prints a table."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import nerf_oracle as O  # noqa: E402
import neural_sim_nerf_b200 as nsr  # noqa: E402
from test_gpu_backward import camera_rays  # noqa: E402


def decode(buf, off, P, W):
    """[P, W] fp16 array stored tile by tile in 8x8 blocks [8 points][8 features] (dump_blocked_off)."""
    a = buf[off:off + P * W * 2].view(torch.float16).view(P // 128, 16, W // 8, 8, 8)
    return a.permute(0, 1, 3, 2, 4).reshape(P, W).double()


def main():
    S, Ni = 64, 128
    n_side = int(os.environ.get('N_SIDE', 12))
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'wfit.npz'))
    sds = [{k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)} for p in ('coarse/', 'fine/')]
    nets = []
    for sd in sds:
        m = nsr.NeRF()
        m.load_state_dict(sd)
        nets.append(m.cuda())
    rays = camera_rays(n_side, 22.5).cuda()
    n, T = rays.shape[0], S + Ni
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(5)).cuda()
    L = nsr.lib()
    P_ = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    new = lambda *s: torch.empty(*s, device='cuda')
    pc, pf = nsr.packed_weights(nets[0]), nsr.packed_weights(nets[1])
    rgb, raw, zv = new(n, 3), new(n, T, 4), new(n, T)
    ws = torch.empty(L.nsr_render_workspace_bytes(n, S, Ni), dtype=torch.uint8, device='cuda')
    mask = torch.empty(L.nsr_relu_mask_bytes(n, T), dtype=torch.uint8, device='cuda')
    for route in ('saved', 'recompute'):
        dump = torch.zeros(L.nsr_mlp_dump_bytes(n, T), dtype=torch.uint8, device='cuda')
        rc = L.nsr_render_rays_forward_ex(P_(rays), n, P_(pc), P_(pf), S, Ni, 0, None, None, P_(rgb), None, None, None, None, None, None, P_(raw),
                                          P_(zv), None, P_(mask), P_(dump) if route == 'saved' else None, None, P_(ws), ws.numel(), None)
        assert rc == 0, L.nsr_last_error()
        g = (2.0 * (rgb - target) / (n * 3)).contiguous()
        bws = torch.zeros(L.nsr_render_backward_workspace_bytes(n, T), dtype=torch.uint8, device='cuda')
        gw = [torch.zeros(s_, device='cuda') for s_ in nsr.run_nerf._EXPECTED_SHAPES]
        gb = [torch.zeros(s_[0], device='cuda') for s_ in nsr.run_nerf._EXPECTED_SHAPES]
        dWp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in gw])
        dBp = (ctypes.c_void_p * 12)(*[t.data_ptr() for t in gb])
        d = new(n, 11)
        rc = L.nsr_render_rays_backward_ex(P_(rays), P_(zv), P_(raw), n, T, P_(pf), 0, P_(g), P_(d), P_(dump), dWp, dBp,
                                           P_(mask) if route == 'saved' else None, None, P_(bws), bws.numel(), None)
        assert rc == 0, L.nsr_last_error()
        torch.cuda.synchronize()
        gscale_raw = bws[-256:-252].view(torch.float32).item()
        gscale = float(np.frombuffer(np.uint32(np.float32(gscale_raw).view(np.uint32) & 0x7f800000).tobytes(), dtype=np.float32)[0])
        print(f'== route {route}: n={n} points={n * T} gmax {gscale_raw:.4e} gscale {gscale:.4e}')

        # ---- fp64 truth on the same depths, every intermediate kept
        for dt in (torch.float64, torch.float32):
            sd = {k: v.cuda().to(dt).requires_grad_(True) for k, v in sds[1].items()}
            r = rays.to(dt)
            pts = (r[:, None, 0:3] + r[:, None, 3:6] * zv.to(dt)[:, :, None]).reshape(-1, 3)
            ex = O.embed(pts, 10)
            ev = O.embed(r[:, None, 8:11].expand(n, T, 3).reshape(-1, 3), 4)
            lin = torch.nn.functional.linear
            pre, hs = [], []
            h = ex
            for i in range(8):
                a = lin(h, sd[f'pts_linears.{i}.weight'], sd[f'pts_linears.{i}.bias'])
                a.retain_grad()
                pre.append(a)
                h = torch.relu(a)
                hs.append(h)
                if i == 4:
                    h = torch.cat([ex, h], -1)
            sigma = lin(h, sd['alpha_linear.weight'], sd['alpha_linear.bias'])
            feat = lin(h, sd['feature_linear.weight'], sd['feature_linear.bias'])
            feat.retain_grad()
            av = lin(torch.cat([feat, ev], -1), sd['views_linears.0.weight'], sd['views_linears.0.bias'])
            av.retain_grad()
            hv = torch.relu(av)
            rgbr = lin(hv, sd['rgb_linear.weight'], sd['rgb_linear.bias'])
            raw_t = torch.cat([rgbr, sigma], -1).reshape(n, T, 4)
            rgb_t = O.raw2outputs(raw_t, zv.to(dt), r[:, 3:6])[0]
            loss = ((rgb_t - target.to(dt)) ** 2).mean()
            loss.backward()
            if dt == torch.float64:
                truth = dict(sd=sd, pre=pre, hs=hs, feat=feat, av=av, hv=hv, ex=ex, ev=ev, raw=raw_t)
            else:
                f32 = sd
        t = truth
        print(f'raw: ours vs fp64 max abs {float((raw.double() - t["raw"]).abs().max()):.3e}; rgb {float((rgb.double() - rgb_t.double()).abs().max()):.3e}')
        names = [f'pts_linears.{i}' for i in range(8)] + ['views_linears.0', 'feature_linear', 'alpha_linear', 'rgb_linear']
        for i, nm in enumerate(names):
            tw, tb = t['sd'][nm + '.weight'].grad, t['sd'][nm + '.bias'].grad
            ew = float((gw[i].double() - tw).abs().max() / tw.abs().max())
            eb = float((gb[i].double() - tb).abs().max() / tb.abs().max())
            fw = float((f32[nm + '.weight'].grad.double() - tw).abs().max() / tw.abs().max())
            print(f'  {nm:18s} dW ours-vs-fp64 {ew:.2e}  dB {eb:.2e}   (fp32 autograd vs fp64: {fw:.2e})')

        # ---- the dump, array by array
        P = ((n * T + 127) // 128) * 128
        lo = 9728 * P // 2 if False else None
        off_h = lambda l: P * 192 + l * P * 512
        off_hv = off_h(9)
        off_gv = off_hv + P * 256
        off_gf = off_gv + P * 256
        off_g = lambda l: off_gf + P * 512 + l * P * 512
        lo = off_g(8)
        assert 2 * lo == dump.numel(), (2 * lo, dump.numel())
        NP = n * T

        def arr(off, W):
            return (decode(dump, off, P, W) + decode(dump, off + lo, P, W))[:NP], decode(dump, off, P, W)[:NP]

        def cmp(what, off, W, ref, scale=1.0):
            full, hi = arr(off, W)
            ref = ref.detach().double()
            ref = torch.nn.functional.pad(ref, (0, W - ref.shape[1]))
            m = float(ref.abs().max())
            e = (full * scale - ref).abs()
            eh = (hi * scale - ref).abs()
            # error weighted the way a column sum sees it
            cs_ref = ref.sum(0)
            cs = (full * scale).sum(0)
            print(f'  {what:4s} max|ref| {m:.3e}  hi+lo err/max {float(e.max()) / m:.2e}  hi-only err/max {float(eh.max()) / m:.2e}  '
                  f'colsum err/max {float((cs - cs_ref).abs().max() / cs_ref.abs().max()):.2e}  rows-with-err>1e-3max {int((e.max(1).values > 1e-3 * m).sum())}')
            return full * scale

        D = {}
        D['EX'] = cmp('EX', 0, 64, t['ex'])
        D['EV'] = cmp('EV', P * 128, 32, t['ev'])
        for l in range(8):
            D[f'H{l}'] = cmp(f'H{l}', off_h(l), 256, t['hs'][l])
        D['F'] = cmp('F', off_h(8), 256, t['feat'])
        D['HV'] = cmp('HV', off_hv, 128, t['hv'])
        D['GV'] = cmp('GV', off_gv, 128, t['av'].grad, gscale)
        D['GF'] = cmp('GF', off_gf, 256, t['feat'].grad, gscale)
        for l in range(8):
            D[f'G{l}'] = cmp(f'G{l}', off_g(l), 256, t['pre'][l].grad, gscale)
        # which rows carry G6's error, and what do they look like
        l = int(os.environ.get('LAYER', 6))
        ref = t['pre'][l].grad.double()
        e = (D[f'G{l}'] - ref).abs().max(1).values
        top = torch.argsort(e, descending=True)[:12]
        for p in top.tolist():
            ray, s = divmod(p, T)
            zz = zv[ray]
            print(f'    point {p} (ray {ray}, sample {s}): err {float(e[p]):.3e} |Gref|max {float(ref[p].abs().max()):.3e}  sigma {float(raw[ray, s, 3]):.4f} '
                  f'dz {float(zz[min(s + 1, T - 1)] - zz[s]):.3e} d_raw {[float(x) for x in bws[:n * T * 16].view(torch.float32).view(n * T, 4)[p]]}')
        # dW from the decoded operands in fp64
        for l in range(8):
            Hin = D['EX'][:, :63] if l == 0 else (torch.cat([D['EX'][:, :63], D['H4']], 1) if l == 5 else D[f'H{l - 1}'])
            dW = D[f'G{l}'].T @ Hin
            tw = t['sd'][f'pts_linears.{l}.weight'].grad
            print(f'  dW{l} from decoded dump (fp64 GEMM) vs fp64 truth: {float((dW - tw).abs().max() / tw.abs().max()):.2e};  kernel dW vs decoded-dump dW: '
                  f'{float((gw[l].double() - dW).abs().max() / tw.abs().max()):.2e}')


if __name__ == '__main__':
    main()
