#!/bin/bash
# A/B of the wide-row listed verify kernels: verify_mode 0 (staged, 2 stages), 5 (staged, 3 stages), 6 (rows through L1)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_prune.py tests/test_gpu_screen.py -x -q -m gpu > gpurun_out/vs_tests.log 2>&1
tail -2 gpurun_out/vs_tests.log
for wl in cfg3 cfg4; do
  for m in 0 7; do
    st=60; [ $wl = cfg4 ] && st=12
    python bench.py --workload $wl --steps $st --warmup 5 --no-cpu-baseline --option verify_mode=$m > gpurun_out/vs_${wl}_m$m.json 2> gpurun_out/vs_${wl}_m$m.err
    python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/vs_${wl}_m$m.json") if l.startswith("{")][-1])
print("$wl verify_mode=$m ms/step", round(j["ms_per_step"],3), j.get("step_breakdown_ms"), "final_cost", j.get("final_cost"))
PY
  done
done
