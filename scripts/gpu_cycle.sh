#!/bin/bash
# One GPU round trip: parity tests, a short bench, an ncu --set full capture of the attempt kernel.
# usage: scripts/gpu_cycle.sh TAG [notest] [noncu]
TAG=${1:-x}
mkdir -p gpurun_out
if [[ "$*" != *notest* ]]; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/tests_$TAG.log
  cat gpurun_out/tests_$TAG.log
fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
r=d["roofline"]
print("value %.4g ms_step %.4f e2e %.4g dyn_ms %.4f env_ms %.4f frac %.4f k %.3f lane_eff %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], r["frac"], r["mean_attempts_per_env_step"], r["warp_divergence"]["lane_efficiency"]))
P
if [[ "$*" != *noncu* ]]; then
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fw_attempt_kernel -s 12 -c 1 -o gpurun_out/prof_att_$TAG -f \
     python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_$TAG.log
fi
