#!/bin/bash
# incremental member sums on/off: tests, then the three Lloyd workloads
mkdir -p gpurun_out
python -m pytest tests/test_gpu_prune.py -x -q -m gpu > gpurun_out/delta_tests.log 2>&1
tail -2 gpurun_out/delta_tests.log
for wl in cfg2 cfg3 cfg4; do
  for m in 1 0; do
    st=300; [ $wl = cfg3 ] && st=100; [ $wl = cfg4 ] && st=20
    python bench.py --workload $wl --steps $st --warmup 5 --no-cpu-baseline --option delta_sums=$m > gpurun_out/delta_${wl}_m$m.json 2> gpurun_out/delta_${wl}_m$m.err
    python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/delta_${wl}_m$m.json") if l.startswith("{")][-1])
print("$wl delta_sums=$m ms/step", round(j["ms_per_step"],3), {k:round(v,3) for k,v in j.get("step_breakdown_ms").items()}, "final_cost", j.get("final_cost"), j["pruning"].get("incremental_sum_steps"), j["pruning"].get("labels_changed_last_step_frac"))
PY
  done
done
