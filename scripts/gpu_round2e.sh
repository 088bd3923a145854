#!/bin/bash
# full GPU suite on the in-tree library, then the interleaved A/B of the given variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
REPS=${REPS:-2} bash scripts/gpu_ab.sh "$@"
