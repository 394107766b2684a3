#!/bin/bash
# incremental member sums on/off: tests, then the three Lloyd workloads
mkdir -p gpurun_out


for wl in cfg2 cfg3; do
  for m in 5 15; do
    st=300; [ $wl = cfg3 ] && st=100; [ $wl = cfg4 ] && st=20
    python bench.py --workload $wl --steps $st --warmup 5 --no-cpu-baseline --option prune_list_margin=$m > gpurun_out/lm_${wl}_m$m.json 2> gpurun_out/lm_${wl}_m$m.err
    python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/lm_${wl}_m$m.json") if l.startswith("{")][-1])
print("$wl list_margin=$m ms/step", round(j["ms_per_step"],3), {k:round(v,3) for k,v in j.get("step_breakdown_ms").items()}, "final_cost", j.get("final_cost"), j["pruning"].get("incremental_sum_steps"), j["pruning"].get("list_reuse_steps"), j["pruning"].get("mean_centers_per_tile_list"))
PY
  done
done
