#!/bin/bash
# new edge-case tests, the driver's smoke(), then the default bench line (-> gpurun_out/bench_$1.json)
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "ragged or capi or on_success or watchdog" 2>&1 | tail -6
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -5 gpurun_out/bench_$TAG.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); r=d["roofline"]; e=d["e2e"]; c=d["controlled_flight"]
print("value %.4g ms %.4f p50 %.4f p99 %.4f frac %.4f dyn %.4f env %.4f | e2e %.4g d1 %.4g open %.4g | fp32 %.4g | cpu %.0f | pid %.4g %s" % (d["value"], d["ms_per_step"], d["step_ms"]["p50"], d["step_ms"]["p99"], r["frac"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], e["value"], e["closed_loop_depth1"]["value"], e["open_loop_depth2"]["value"], d["fp32_mode"]["value"], d["cpu_baseline"]["value"], c["value"], c["kernels"]))
P
