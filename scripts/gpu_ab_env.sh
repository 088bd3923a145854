#!/bin/bash
# A/B an environment variable on the bench: scripts/gpu_ab_env.sh VAR v1 v2 ...
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 60 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err || tail -3 gpurun_out/bench_ab.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_ab.json")); r=d["roofline"]
print("$VAR=$v value %.4g ms_step %.4f p50 %.4f e2e %.4g dyn_ms %.4f env_ms %.4f serial %.4f" % (d["value"], d["ms_per_step"], d["step_ms"]["p50"], d["e2e"]["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], d["overlap"]["serial_ms_per_step"]))
P
done
