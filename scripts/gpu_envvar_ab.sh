#!/bin/bash
# A/B runtime switches of the shipped library on the bench workload: scripts/gpu_envvar_ab.sh "FWGYM_PREFETCH=0" "FWGYM_PREFETCH=1" ...
for rep in $(seq 1 ${REPS:-2}); do
for setting in "$@"; do
  env $setting timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ev.json 2> gpurun_out/bench_ev.err || tail -3 gpurun_out/bench_ev.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_ev.json")); r=d["roofline"]
print("[$setting #$rep] value %.4g (%.1f us/step) e2e %.4g dyn_ms %.4f env_ms %.4f serial %.4f watchdog %.0f" % (d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], d["overlap"]["serial_ms_per_step"], d["overlap"]["watchdog"]))
P
done
done
