#!/bin/bash
# First GPU call of the next round (one gpurun call, ~4 GPU-minutes): the measurements profiles/r01_notes.md ends with.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/next_round_first_call.sh'
# 1. hi-only operands on wide rows (screen_terms=2: rows of 32+ floats) against the 3-term split: cfg3 through bench.py,
#    a 2e6-frame cfg4 slice through config_bench.py (Lloyd step, screen kernel time, candidate statistics)
# 2. CTA-pair MMAs on top of it (screen_cluster=3)
# Labels are exact in every mode (tests/test_gpu_screen.py runs both term counts and the cluster modes).
mkdir -p gpurun_out
P='import sys,json; j=json.loads(sys.stdin.read()); r=j["roofline"]; print(j["ms_per_step"], r["kernel_ms"], r["frac"], j["clocks"])'
for o in "screen_terms=0" "screen_terms=2" "screen_terms=2 --option screen_cluster=3"; do
  echo "== cfg3 (1.25e7 x 64, k=2000) --option $o"
  timeout 150 python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline --option $o 2>&1 | tail -1 | python -c "$P"
done > gpurun_out/next_terms_cfg3.log 2>&1
for o in "screen_terms=0" "screen_terms=2"; do
  echo "== cfg4 slice (2e6 x 256, k=5000) --option $o"
  timeout 200 python tools/config_bench.py --cfg 4 --cfg4-frames 2000000 --no-kmpp --option $o --out gpurun_out/next_cfg4_$o.json 2>&1 | tail -2 | cut -c1-1200
done > gpurun_out/next_terms_cfg4.log 2>&1
cat gpurun_out/next_terms_cfg3.log gpurun_out/next_terms_cfg4.log
