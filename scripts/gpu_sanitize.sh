#!/bin/bash
# compute-sanitizer passes over the step path: logs -> gpurun_out/sanitizer_<tool>_$TAG.log
TAG=${1:-x}
mkdir -p gpurun_out
for tool in memcheck initcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitizer_${tool}_$TAG.log \
    python scripts/sanitize_target.py > gpurun_out/sanitizer_${tool}_$TAG.out 2>&1
  echo "$tool rc=$? $(tail -1 gpurun_out/sanitizer_${tool}_$TAG.log)"
done
