#!/bin/bash
# Round-end evidence on one B200: parity tests, sanitizer passes, bench line, ncu captures of the three kernels of a step,
# launch list, BASELINE.md §3 report (GPU arm).  usage: scripts/gpu_final_evidence.sh TAG
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/tests_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || tail -5 gpurun_out/bench_$TAG.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_$TAG.json")); r=d["roofline"]; e=d["e2e"]
print("value %.4g ms %.4f p50 %.4f p99 %.4f frac %.4f dyn %.4f env %.4f | e2e %.4g d1 %.4g open %.4g | fp32 %.4g | cpu %.0f" % (d["value"], d["ms_per_step"], d["step_ms"]["p50"], d["step_ms"]["p99"], r["frac"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], e["value"], e["closed_loop_depth1"]["value"], e["open_loop_depth2"]["value"], d["fp32_mode"]["value"], d["cpu_baseline"]["value"]))
P
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; head -c 300 gpurun_out/bench_ref_$TAG.json; echo
bash scripts/gpu_launchlist.sh $TAG | tail -7
for k in attempt env init; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fw_${k}_kernel -s 212 -c 1 -o gpurun_out/prof_${k}_$TAG -f \
     python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_${k}_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_${k}_$TAG.log
done
bash scripts/gpu_sanitize.sh $TAG
[ -n "$WITH_BASELINE_REPORT" ] && python scripts/baseline_report.py --reuse-cpu profiles/r2_baseline_report_cpu_arm.json 2>&1 | grep -E "^GPU|wrote"
