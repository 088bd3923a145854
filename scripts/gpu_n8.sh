#!/bin/bash
# N-GPU bench runs with different host-pipeline settings, one JSON summary line each into gpurun_out/n8_summary.txt
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 50 --warmup 5 2>gpurun_out/bench_n8_$2.err | tail -1 > gpurun_out/bench_n8_$2.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_n8_$2.json')); print('$2', '%.4g' % d['value'], '%.4g' % d['e2e']['value'], d['e2e']['us_per_step_by_rank'], d['e2e']['rank0_host_us_per_step'])" >> gpurun_out/n8_summary.txt; }
rm -f gpurun_out/n8_summary.txt
run 29541 base
FWGYM_HOST_BLOCKING=1 run 29542 blocking
FWGYM_HOST_DEPTH=3 run 29543 depth3
cat gpurun_out/n8_summary.txt
