#!/bin/bash
# Where an attempt-kernel warp spends its cycles (experiment builds, scripts/build_variant.sh):
#   seg     -DFW_TIME_SEGMENTS : the divergence counters carry cycle sums (refill / total / park)
#   nofence -DFW_PARK_NO_FENCE : parking without the release fence (FWGYM_OVERLAP=0 only)
V=build/variants
run() { # name lib overlap
  FWGYM_OVERLAP=$3 FWGYM_LIB=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err || tail -3 gpurun_out/bench_$1.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_$1.json")); r=d["roofline"]; wd=r["warp_divergence"]
print("[$1 overlap=$3] value %.4g dyn_ms %.4f env_ms %.4f  passes %.0f lane_att %.0f watchdog %.0f" % (d["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], wd["warp_passes"], wd["lane_attempts"], d["overlap"]["watchdog"]))
P
}
run base1 $V/libfwgym_base.so 1
run base0 $V/libfwgym_base.so 0
run nofence0 $V/libfwgym_nofence.so 0
run seg0 $V/libfwgym_seg.so 0
run seg1 $V/libfwgym_seg.so 1
