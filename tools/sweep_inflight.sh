#!/bin/bash
# bench.py over frames-in-flight settings:  tools/sweep_inflight.sh "1x378 2x189 3x126" [extra env]
mkdir -p gpurun_out
for cfg in $1; do
  ns=${cfg%x*}; bf=${cfg#*x}
  env VKN_STREAMS=$ns VKN_BATCH=$bf $2 python bench.py --steps 20 --warmup 5 2>gpurun_out/sweep_$cfg.err > gpurun_out/sweep_$cfg.json
  python tools/bench_summary.py gpurun_out/sweep_$cfg.json
done
