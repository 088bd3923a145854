#!/bin/bash
# usage: gpu_sweep.sh "ENV=val ENV2=val" "ENV=val" ...   (bench only, one line each)
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sw.json 2> gpurun_out/bench_sw.err || tail -3 gpurun_out/bench_sw.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_sw.json"))
r=d["roofline"]
print("[$cfg] value %.4g e2e %.4g dyn_ms %.4f env_ms %.4f frac %.4f lane_eff %.3f" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], r["frac"], r["warp_divergence"]["lane_efficiency"]))
P
done
