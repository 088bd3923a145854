#!/bin/bash
# parity subset on a library variant ($1), then the interleaved A/B of the remaining arguments
mkdir -p gpurun_out
V=$1; shift
FWGYM_LIB=build/variants/libfwgym_$V.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
REPS=${REPS:-2} bash scripts/gpu_ab.sh "$@"
