#!/bin/bash
# parity subset on the in-tree library, then an interleaved A/B of library variants (scripts/gpu_ab.sh)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
REPS=${REPS:-2} bash scripts/gpu_ab.sh "$@"
