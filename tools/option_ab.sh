#!/bin/bash
# A/B of library options on a bench workload:  tools/option_ab.sh WORKLOAD STEPS "name=value[,name=value]" ...
wl=$1; st=$2; shift 2
mkdir -p gpurun_out
for opts in "$@"; do
  args=""; for o in ${opts//,/ }; do args="$args --option $o"; done
  python bench.py --workload $wl --steps $st --warmup 5 --no-cpu-baseline $args > gpurun_out/ab_tmp.json 2> gpurun_out/ab_tmp.err
  python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/ab_tmp.json") if l.startswith("{")][-1])
p=j.get("pruning",{})
print("$wl $opts ms/step", round(j["ms_per_step"],3), {k:round(v,3) for k,v in (j.get("step_breakdown_ms") or {}).items()}, "final_cost", j.get("final_cost"), "list", round(p.get("mean_centers_per_tile_list",0),1), "kept", p.get("list_reuse_steps"))
PY
done
