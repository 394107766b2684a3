#!/bin/bash
# operand-term sweep of the wide-row screen (forced 1 / 2 / 3 terms and the measured choice) on cfg3 and a cfg4 slice
mkdir -p gpurun_out
for t in 3 2 1 0; do
  python tools/config_bench.py --cfg 3 --option screen_terms=$t --out gpurun_out/terms_cfg3_t$t.json
done 2>&1 | tee gpurun_out/terms_cfg3.log
for t in 3 2 1 0; do
  python tools/config_bench.py --cfg 4 --cfg4-frames ${CFG4_FRAMES:-2000000} --no-kmpp --option screen_terms=$t --out gpurun_out/terms_cfg4_t$t.json
done 2>&1 | tee gpurun_out/terms_cfg4.log
