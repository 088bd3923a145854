#!/bin/bash
# attempt-kernel pass time vs resident warps per SM: scripts/gpu_warpsweep.sh 2 4 6 8
for w in "$@"; do
  FWGYM_ATTEMPT_WARPS_PER_SM=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err || tail -3 gpurun_out/bench_ws.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_ws.json")); r=d["roofline"]; wd=r["warp_divergence"]
warps=$w*148
passes_per_warp=wd["warp_passes"]/20/warps
print("[W=$w] value %.4g dyn_ms %.4f passes/warp %.2f -> T_pass %.2f us lane_eff %.3f" % (d["value"], r["kernel_ms_per_launch"], passes_per_warp, (r["kernel_ms_per_launch"]-0.018)*1e3/passes_per_warp, wd["lane_efficiency"]))
P
done
