#!/bin/bash
# GPU global-timer timeline of one env step, with and without the env kernel's pre-wait draws
for v in timeline timeline_nopredraw; do
  for ov in 1 0; do
    FWGYM_LIB=build/variants/libfwgym_$v.so FWGYM_OVERLAP=$ov timeout 300 python scripts/gpu_timeline.py 2>&1 | tail -1 | sed "s/^/$v: /"
  done
done
