#!/bin/bash
# default bench line + reference arm + ncu launch list; usage: scripts/gpu_bench_only.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -5 gpurun_out/bench_$TAG.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); r=d["roofline"]; e=d["e2e"]; c=d["controlled_flight"]
print("value %.4g ms %.4f p50 %.4f p99 %.4f frac %.4f dyn %.4f env %.4f | e2e %.4g d1 %.4g open %.4g | fp32 %.4g | cpu %.0f" % (d["value"], d["ms_per_step"], d["step_ms"]["p50"], d["step_ms"]["p99"], r["frac"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], e["value"], e["closed_loop_depth1"]["value"], e["open_loop_depth2"]["value"], d["fp32_mode"]["value"], d["cpu_baseline"]["value"]))
print("controlled_flight", json.dumps(c)[:900])
P
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; head -c 300 gpurun_out/bench_ref_$TAG.json; echo
bash scripts/gpu_launchlist.sh $TAG | tail -7
