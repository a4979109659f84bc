#!/bin/bash
# Runs tools/fp8_probe with nvidia-smi sampling clocks / power beside it.  Output: gpurun_out/fp8_probe.{log,smi.csv}
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 250 > gpurun_out/fp8_probe.smi.csv &
SMI=$!
timeout 120 ./tools/fp8_probe ${1:-3.0} > gpurun_out/fp8_probe.log 2>&1
echo "exit $?" >> gpurun_out/fp8_probe.log
kill $SMI
cat gpurun_out/fp8_probe.log
