#!/bin/bash
# bench prebuilt library variants: scripts/gpu_libsweep.sh path1.so path2.so ...   (FWGYM_LIB selects the library)
for lib in "$@"; do
  FWGYM_LIB=$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lib.json 2> gpurun_out/bench_lib.err || tail -3 gpurun_out/bench_lib.err
  python - <<P
import json
d=json.load(open("gpurun_out/bench_lib.json"))
r=d["roofline"]
print("[$lib] value %.4g e2e %.4g dyn_ms %.4f env_ms %.4f" % (d["value"], d["e2e"]["value"], r["kernel_ms_per_launch"], d["env_kernel"]["ms_per_launch"]))
P
done
