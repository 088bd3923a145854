#!/bin/bash
# full ncu capture of one launch of kernel matching regex $2 (skip $3 launches) -> gpurun_out/prof_$1.ncu-rep
TAG=$1; K=$2; SKIP=${3:-12}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$K -s $SKIP -c 1 -o gpurun_out/prof_$TAG -f \
   python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
