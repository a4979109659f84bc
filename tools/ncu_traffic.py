"""profiles/ncu_traffic.json from an ncu --set full capture of tools/ncu_kernels.py (run here, the .ncu-rep comes back in gpurun_out/):
  python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep profiles/r02_ncu_kernels.csv > profiles/ncu_traffic.json
(or give it the `ncu -i ... --page raw --csv` dump made on the GPU box, when the report itself is too big to bring back)
Per launch: dram__bytes_read.sum + dram__bytes_write.sum (what bench.py reports as roofline.traffic), duration, tensor-pipe activity."""
import csv, io, json, subprocess, sys
FWD = ['coarse_tier1', 'coarse_tier2', 'coarse_redo', 'composite_coarse', 'resample_merge', 'fine_tier1', 'fine_tier2', 'fine_redo', 'composite_fine']
ORDER = FWD + ['pg_' + k for k in FWD] + ['composite_bwd', 'bwd_masked_active', 'ray_grad_reduce', 'fine_dense']
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum']
rep = sys.argv[1]
out = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
header, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(header)}
num = lambda r, k: float(r[col[k]].replace(',', '')) if k in col and r[col[k]] not in ('', 'n/a') else None
to_bytes = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
res = {}
assert len(data) == len(ORDER), (len(data), [r[col['Kernel Name']][:40] for r in data])
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
        w = csv.writer(f)
        w.writerow(['launch'] + KEEP)
        w.writerow(['unit'] + [units[col[k]] if k in col else '' for k in KEEP])
        for key, r in zip(ORDER, data):
            w.writerow([key] + [r[col[k]] if k in col else '' for k in KEEP])
for key, r in zip(ORDER, data):
    rd = num(r, 'dram__bytes_read.sum') * to_bytes[units[col['dram__bytes_read.sum']]]
    wr = num(r, 'dram__bytes_write.sum') * to_bytes[units[col['dram__bytes_write.sum']]]
    res[key] = {'kernel': r[col['Kernel Name']].split('(')[0], 'dram_bytes': int(rd + wr), 'dram_read_bytes': int(rd), 'dram_write_bytes': int(wr),
                'duration_under_ncu': f"{r[col['gpu__time_duration.sum']]} {units[col['gpu__time_duration.sum']]}",
                'tensor_pipe_active_pct': num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                'source': f'{rep.split("/")[-1]} (ncu --set full --clock-control none of tools/ncu_kernels.py), launch "{key}"'}
print(json.dumps(res, indent=1))
