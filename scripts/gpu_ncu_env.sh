#!/bin/bash
# full ncu capture (with source) of the env kernel of the in-tree or FWGYM_LIB library -> gpurun_out/prof_env_$1.ncu-rep
TAG=$1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fw_env_kernel -s 212 -c 1 -o gpurun_out/prof_env_$TAG -f \
   python bench.py --steps 10 --warmup 5 --no-cpu-baseline --e2e-steps 5 > gpurun_out/ncu_env_$TAG.log 2>&1
tail -1 gpurun_out/ncu_env_$TAG.log | cut -c1-200
