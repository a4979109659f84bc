#!/bin/bash
# First GPU contact: descriptor probe, parity tests, smoke, a short bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 300 ./tools/umma_probe > gpurun_out/probe.txt 2>&1
echo "probe exit $?" >> gpurun_out/probe.txt
SWAP=0
if grep -q "N=128 variant 0 .*FAIL" gpurun_out/probe.txt && grep -q "N=128 variant 1 .*PASS" gpurun_out/probe.txt; then SWAP=1; fi
echo "NSR_DESC_SWAP=$SWAP" >> gpurun_out/probe.txt
export NSR_DESC_SWAP=$SWAP
timeout 900 python -m pytest tests -m gpu -q -rA -s > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.txt 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/probe.txt; tail -30 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -5; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
