#!/bin/bash
# sweep FWGYM_STAGES on the GPU box (bench only)
for st in "${@}"; do
  FWGYM_STAGES="$st" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_st.json 2> gpurun_out/bench_st.err || tail -3 gpurun_out/bench_st.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_st.json"))
r=d["roofline"]
print("stages [$st] value %.4g e2e %.4g dyn_ms %.4f env_ms %.4f frac %.4f lane_eff %.3f" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"], r["frac"], r["warp_divergence"]["lane_efficiency"]))
P
done
