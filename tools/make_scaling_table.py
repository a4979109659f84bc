"""profiles/r02_scaling_and_stages.md from the committed bench / bilevel / ncu JSON files of the round:  python tools/make_scaling_table.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda f: os.path.join(ROOT, 'profiles', f)


def load(f):
    return [json.loads(l) for l in open(P(f)) if l.startswith('{')][0]


b1, b2, b4, b8 = load('r02_v4_bench.json'), load('r02_v4_bench_n2.json'), load('r02_v4_bench_n4.json'), load('r02_v3_bench_n8.json')
ref, bl = load('r02_v1_bench_reference_arm.json'), load('r02_v1_bilevel_stub_n8_k50.json')
tr = json.load(open(P('ncu_traffic.json')))
L = ['# Round 2: scaling, stages, baselines (B200, sm_100a)\n',
     'Sources: `r02_v4_bench.json` (N=1), `r02_v4_bench_n2.json`, `r02_v4_bench_n4.json`, `r02_v3_bench_n8.json` (final build; one box each),\n'
     '`r02_v1_bench_reference_arm.json`, `r02_v1_bilevel_stub_n8_k50.json`, `ncu_traffic.json` / `r02_ncu_kernels.csv` (ncu --set full of\n'
     '`tools/ncu_kernels.py`).  The driver\'s SCALE_r02.json is the authoritative N = 1, 2, 4, 8 series.\n',
     '## Throughput (whole job, rays/s; one 400x400 image per GPU per step, 64 + 128 samples)\n',
     '| N GPUs | forward, rays in HBM | forward e2e (host rays in, maps out) | ms / step | pose_grad (fwd + bwd + NCCL all-reduce) | objects8 (config 4) |',
     '|---|---|---|---|---|---|']
for n, b in ((1, b1), (2, b2), (4, b4), (8, b8)):
    o = b['objects8']
    L.append(f"| {n} | {b['value']/1e6:.3f} M | {b['e2e']['value']/1e6:.3f} M | {b['ms_per_step']:.1f} | {b['pose_grad']['rays_per_s']/1e6:.3f} M ({b['pose_grad']['ms_per_step']:.1f} ms) | "
             f"{o['rays_per_s']/1e6:.3f} M ({o['ms_per_step']:.0f} ms per 8 images) |")
L.append(f"\nScaling of the forward path: {b2['value']/b1['value']:.2f}x at 2 GPUs, {b4['value']/b1['value']:.2f}x at 4, {b8['value']/b1['value']:.2f}x at 8 (no forward collective; every GPU at its own 1 kW cap, SM clock ~1.63 of 1.97 GHz).\n")
L.append('## Baselines on the same box\n')
c, bs = b1['cpu_baseline'], b1['baselines']
e = bs['eager_pytorch_on_this_gpu']
L.append(f"* the UNMODIFIED reference `render()` on the host cores (`oracle/_ref` bytecode build, {c['cores']} torch threads = fastest tried): **{c['value']:.0f} rays/s** in the default run, "
         f"{ref['value']:.0f} rays/s in the `--impl reference` arm -> this path is **{b1['e2e']['value']/c['value']:.0f}x** end to end;")
L.append(f"* the reference's own eager PyTorch path on this B200 (RN.render, modules on cuda): {e['fp32_tf32_off_chunk512']['rays_per_s']/1e3:.0f} k rays/s at its chunk = 512 (fp32), "
         f"{e['fp32_tf32_off_chunk32768']['rays_per_s']/1e3:.0f} k at chunk = 32768, {e['fp32_tf32_on_chunk32768']['rays_per_s']/1e3:.0f} k with TF32 allowed -> "
         f"**{b1['value']/e['fp32_tf32_off_chunk512']['rays_per_s']:.0f}x / {b1['value']/e['fp32_tf32_off_chunk32768']['rays_per_s']:.0f}x / {b1['value']/e['fp32_tf32_on_chunk32768']['rays_per_s']:.0f}x**;")
L.append(f"* BASELINE config 1 (200x200, 64 coarse samples only): reference on the host {bs['config1_cpu']['rays_per_s']:.0f} rays/s, this path {bs['config1_gpu']['rays_per_s']/1e6:.1f} M rays/s "
         f"({bs['config1_gpu']['ms_per_image']:.2f} ms per image);")
L.append(f"* BASELINE config 3 (RN:168-181 pattern: render + autograd.grad per 512-ray chunk): reference on the host {bs['config3_cpu']['rays_per_s']:.0f} rays/s; this path "
         f"{b1['pose_grad']['rays_per_s']/1e6:.2f} M rays/s (`pose_grad`, saved sign bits) / {b1['fwd_bwd']['rays_per_s']/1e6:.2f} M (`fwd_bwd`, recompute over the active set).")
L.append(f"* parity of the timed image against the reference's pixels (4096 rays): max rel err {b1['parity']['max_rel_err']:.2e} (bar 1e-3), NaN disparity masks equal.\n")
L.append('## Kernels of one image (timed back to back in bench.py; DRAM bytes from ncu)\n')
L.append('| kernel | ms / launch | algorithmic TFLOP/s | frac of measured tensor peak (1375.5) | MMAs per product | DRAM bytes (ncu) |')
L.append('|---|---|---|---|---|---|')
for r in b1['roofline_kernels']:
    L.append(f"| {r['kernel']} | {r['ms_per_launch']:.2f} | {r['achieved']:.0f} | {r['frac']:.3f} | {r['tensor_flop_issued_per_algorithmic_flop']} | {r['traffic']/1e6:.0f} MB |")
rs = b1['roofline_step']
L.append(f"\nWhole forward step in FLOPs of the reference algorithm: {rs['achieved']:.0f} TFLOP/s = {rs['frac']:.2f} of the peak (it counts the work the certified-empty points never do here).\n")
L.append('| ray-stage kernel (timed alone, 20 launches) | ms | algorithmic bytes | GB/s | frac of HBM peak | DRAM bytes in a real render (ncu) | GB/s on those |')
L.append('|---|---|---|---|---|---|---|')
for s_ in b1['stages'][:-1]:
    L.append(f"| {s_['kernel']} | {s_['ms']:.4f} | {s_['algorithmic_bytes']/1e6:.1f} MB | {s_['gb_per_s']:.0f} | {s_['frac_of_hbm_peak']:.3f} | "
             f"{('%.1f MB' % (s_['dram_bytes_ncu']/1e6)) if 'dram_bytes_ncu' in s_ else '-'} | {('%.0f' % s_['gb_per_s_ncu_bytes']) if 'gb_per_s_ncu_bytes' in s_ else '-'} |")
fwd = [k for k in tr if not k.startswith('pg_') and k not in ('composite_bwd', 'bwd_masked_active', 'ray_grad_reduce', 'fine_dense')]
L.append(f"\nDRAM traffic of one forward image, all ten launches (ncu): {sum(tr[k]['dram_bytes'] for k in fwd)/1e9:.2f} GB (tier 1 writes (0,0,0,sigma~) for every point: "
         f"{tr['fine_tier1']['dram_write_bytes']/1e6:.0f} MB in the fine pass).\n")
L.append('## Other legs (N = 1)\n')
t = b1['train_step']
L.append(f"* `train_step`: one Adam iteration on 1024 rays {t['ms_per_step']:.2f} ms through render() + autograd + torch.optim.Adam, {t['fused']['ms_per_step']:.2f} ms as one `nsr_train_step` call;")
L.append(f"* two-tier: coarse / fine active fraction {b1['two_tier']['coarse_active_fraction']:.3f} / {b1['two_tier']['fine_active_fraction']:.3f}, max |sigma~ - sigma| on active points "
         f"{b1['two_tier']['fine_max_dsigma_on_active']:.2f}; the same image evaluated densely: {b1['two_tier']['dense_fp16x3']['ms_per_step']:.1f} ms;")
L.append(f"* opt-in precisions (informational, not parity-valid): fp16 {b1['fast_fp16_mode']['ms_per_step']:.1f} ms, mixed f8 {b1['mixed_f8_mode']['ms_per_step']:.1f} ms per image (dense);")
L.append(f"* bilevel outer loop with a stub detector, K = 50 poses on 8 GPUs: {bl['epochs'][-1]['epoch_s']:.2f} s per epoch ({bl['epochs'][-1]['render_images_s']:.2f} s rendering 50 PNGs, "
         f"{bl['epochs'][-1]['render_images_grad_s']:.2f} s back-propagating them): {bl['render_images_rays_per_s']/1e6:.1f} / {bl['render_images_grad_rays_per_s']/1e6:.1f} M rays/s.")
open(P('r02_scaling_and_stages.md'), 'w').write('\n'.join(L) + '\n')
print(P('r02_scaling_and_stages.md'))
