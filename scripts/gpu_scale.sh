#!/bin/bash
# bench.py at N GPUs of one box (torchrun): scripts/gpu_scale.sh N TAG
N=$1; TAG=$2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err || tail -5 gpurun_out/bench_${TAG}_n$N.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_${TAG}_n$N.json")); e=d["e2e"]
print("N=$N value %.4g (%.1f us) e2e closed %.4g depth1 %.4g open %.4g by rank %s" % (d["value"], 1e3*d["ms_per_step"], e["value"], e["closed_loop_depth1"]["value"], e["open_loop_depth2"]["value"], e["us_per_step_by_rank"]))
P
