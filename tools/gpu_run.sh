#!/bin/bash
# usage: tools/gpu_run.sh [tests] [smoke] [bench] [launches] [ncu]
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 1200 python -m pytest tests -m gpu -q -rA -s > gpurun_out/pytest_gpu.txt 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
  grep -E "passed|failed|PASSED|FAILED|max err|err:" gpurun_out/pytest_gpu.txt | tail -40 ;;
tt)
  timeout 900 python -m pytest tests/test_gpu_two_tier.py -m gpu -q -rA -s -x > gpurun_out/pytest_tt.txt 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_tt.txt
  grep -E "passed|failed|PASSED|FAILED|active|Error|error|assert" gpurun_out/pytest_tt.txt | tail -40 ;;
smoke)
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1
  echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -3 gpurun_out/smoke.txt ;;
bench)
  timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.txt 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err; cat gpurun_out/bench.txt; tail -3 gpurun_out/bench.err ;;
benchref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.txt ;;
bench[248])
  N=${what#bench}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.txt 2> gpurun_out/bench_n$N.err
  echo "bench$N exit $?" >> gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.txt; tail -3 gpurun_out/bench_n$N.err ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --profile > gpurun_out/launches.log 2>&1
  echo "launches exit $?" >> gpurun_out/launches.log; tail -2 gpurun_out/launches.log ;;
ncu)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp -s 3 -c 1 -f -o gpurun_out/prof \
      python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu.log 2>&1
  echo "ncu exit $?" >> gpurun_out/ncu.log; tail -3 gpurun_out/ncu.log ;;
ncubwd)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_bwd -s 1 -c 1 -f -o gpurun_out/prof_bwd \
      python tools/trace_bwd_plain.py > gpurun_out/ncu_bwd.log 2>&1
  echo "ncubwd exit $?" >> gpurun_out/ncu_bwd.log; tail -3 gpurun_out/ncu_bwd.log ;;
bilevel)
  timeout 600 python tools/bilevel_stub.py --epochs 3 --K 8 > gpurun_out/bilevel_n1.txt 2> gpurun_out/bilevel_n1.err
  echo "bilevel exit $?" >> gpurun_out/bilevel_n1.err; cat gpurun_out/bilevel_n1.txt; tail -3 gpurun_out/bilevel_n1.err ;;
bilevel[248])
  N=${what#bilevel}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      tools/bilevel_stub.py --epochs 3 --K 8 > gpurun_out/bilevel_n$N.txt 2> gpurun_out/bilevel_n$N.err
  echo "bilevel$N exit $?" >> gpurun_out/bilevel_n$N.err; cat gpurun_out/bilevel_n$N.txt; tail -3 gpurun_out/bilevel_n$N.err ;;
ncubwdm)
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:nerf_mlp_bwd -s 1 -c 1 -f -o gpurun_out/prof_bwd_masked \
      python tools/trace_bwd_masked.py > gpurun_out/ncu_bwd_masked.log 2>&1
  echo "ncubwdm exit $?" >> gpurun_out/ncu_bwd_masked.log; tail -3 gpurun_out/ncu_bwd_masked.log ;;
esac
done
