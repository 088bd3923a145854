#!/bin/bash
# ncu capture (with source) of the in-tree attempt kernel, then the pass time of base / comb at 1, 2, 4 attempt warps per SM
mkdir -p gpurun_out
TAG=${1:-r2comb}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fw_attempt_kernel -s 212 -c 1 -o gpurun_out/prof_attempt_$TAG -f \
   python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_attempt_$TAG.log 2>&1
tail -1 gpurun_out/ncu_attempt_$TAG.log | cut -c1-200
for w in 1 4; do
for name in base comb; do
  FWGYM_ATTEMPT_WARPS_PER_SM=$w FWGYM_LIB=build/variants/libfwgym_$name.so timeout 300 python bench.py --steps 10 --warmup 3 --burn-in 200 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_w${w}_$name.json 2> gpurun_out/bench_w${w}_$name.err || tail -3 gpurun_out/bench_w${w}_$name.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_w${w}_$name.json")); r=d["roofline"]; wd=r["warp_divergence"]
print("[$name warps/SM=$w] %.1f us/step dyn_ms %.4f lane_eff %.3f passes %.0f" % (d["ms_per_step"]*1e3, r["kernel_ms_per_launch"], wd["lane_efficiency"], wd["warp_passes"]))
P
done
done
