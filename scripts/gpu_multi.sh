#!/bin/bash
# run one copy of a script per GPU concurrently: scripts/gpu_multi.sh NGPU script.py [args]; prints the last lines of each
N=$1; shift
for i in $(seq 0 $((N-1))); do CUDA_VISIBLE_DEVICES=$i python "$@" > gpurun_out/multi_$i.log 2>&1 & done
wait
for i in $(seq 0 $((N-1))); do echo "== gpu $i"; tail -${TAILN:-3} gpurun_out/multi_$i.log; done
