#!/bin/bash
# idle power, then power while streaming weights at the MLP kernel's rate (gap 768) and at the maximum rate (gap 0)
q() { nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader -i 0; }
echo "idle: $(q)"; sleep 1; echo "idle: $(q)"
for gap in 768 384 0; do
  ./tools/stream_power $gap 5 &
  pid=$!
  sleep 2; a=$(q); sleep 1; b=$(q); sleep 1; c=$(q)
  wait $pid
  echo "gap $gap: $a | $b | $c"
done
