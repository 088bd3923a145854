#!/bin/bash
# closed-loop e2e with the part handles' attempt kernels sized to co-reside: "PARTS:WARPS_PER_SM" pairs
mkdir -p gpurun_out
for pw in "$@"; do
  p=${pw%%:*}; w=${pw#*:}
  if [ "$w" = "0" ]; then unset FWGYM_ATTEMPT_WARPS_PER_SM; else export FWGYM_ATTEMPT_WARPS_PER_SM=$w; fi
  FWGYM_E2E_PARTS=$p timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_share_${p}_${w}.json 2> gpurun_out/bench_share_${p}_${w}.err || tail -3 gpurun_out/bench_share_${p}_${w}.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_share_${p}_${w}.json")); e=d["e2e"]
print("[parts=$p warps/SM=$w] value %.4g (%.1f us) e2e %.4g us/step %s host %s" % (d["value"], d["ms_per_step"]*1e3, e["value"], e["us_per_step_by_rank"], e["rank0_host_us_per_step"]))
P
done
