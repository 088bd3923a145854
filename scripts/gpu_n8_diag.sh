#!/bin/bash
# 8-GPU host-side diagnosis: topology, NUMA placement, concurrent D2H bandwidth, then the bench.
mkdir -p gpurun_out
{
echo "== nproc $(nproc)"; lscpu | grep -iE "socket|numa|model name|^cpu\(s\)" 
echo "== topo"; nvidia-smi topo -m
echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then echo "$(basename $d) numa=$(cat $d/numa_node)"; fi; done
echo "== nodes"; for n in /sys/devices/system/node/node*; do echo "$(basename $n): $(cat $n/cpulist)"; done
echo "== affinity"; taskset -p $$ 
} > gpurun_out/n8_topology.txt 2>&1
python scripts/gpu_d2h_concurrent.py > gpurun_out/n8_d2h.txt 2>&1
tail -30 gpurun_out/n8_d2h.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_n8_r2.json 2> gpurun_out/bench_n8_r2.err
python - <<P
import json
d=json.load(open("gpurun_out/bench_n8_r2.json"))
e=d["e2e"]
print("N=8 value %.4g e2e %.4g depth1 %.4g open %.4g by rank %s" % (d["value"], e["value"], e["closed_loop_depth1"]["value"], e["open_loop_depth2"]["value"], e["us_per_step_by_rank"]))
print(e["host_placement_by_rank"])
P
