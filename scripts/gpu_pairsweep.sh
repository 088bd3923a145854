#!/bin/bash
# attempt kernels at 1 / 2 / 4 / 8 resident groups of 32 aircraft per SM: one-thread kernel vs two-warp kernel
for pair in 0 1; do for w in 1 2 4 8; do
  FWGYM_PAIR=$pair FWGYM_ATTEMPT_WARPS_PER_SM=$w timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err || tail -3 gpurun_out/bench_ws.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_ws.json")); r=d["roofline"]; wd=r["warp_divergence"]
groups=$w*148
ppg=wd["warp_passes"]/20/groups
print("[pair=$pair groups/SM=$w] value %.4g dyn_ms %.4f passes/group %.2f -> T_pass_eff %.2f us lane_eff %.3f" % (d["value"], r["kernel_ms_per_launch"], ppg, (r["kernel_ms_per_launch"]-0.016)*1e3/ppg, wd["lane_efficiency"]))
P
done; done
for mix in 0 1; do
FWGYM_PAIR_MIX=$mix timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mix.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/bench_mix.json')); print('mix=$mix value %.4g dyn_ms %.4f'%(d['value'], d['roofline']['kernel_ms_per_launch']))"
done
