#!/bin/bash
# final evidence of the round: GPU tests, the five workloads, the reference arm, smoke
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r02f_gpu_tests.log 2>&1
tail -4 gpurun_out/r02f_gpu_tests.log
python __graft_entry__.py smoke > gpurun_out/r02f_smoke.log 2>&1; tail -1 gpurun_out/r02f_smoke.log
for wl in cfg1 cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $wl > gpurun_out/r02f_bench_$wl.json 2> gpurun_out/r02f_bench_$wl.err
  python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/r02f_bench_$wl.json") if l.startswith("{")][-1])
print("$wl ms/step", round(j["ms_per_step"],3), "e2e ms", round(j["e2e"]["ms_per_step"],3), "roofline", j["roofline"]["bound"], round(j["roofline"]["frac"],4), j.get("step_breakdown_ms"), "clocks", j.get("clocks"))
PY
done
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
tail -c 600 gpurun_out/r02f_bench_reference.json
