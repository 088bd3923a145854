#!/bin/bash
# ncu launch list (gpu__time_duration + a few counters) for a short bench run; usage: gpu_launchlist.sh TAG
TAG=$1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:fw_ -s 40 -c 16 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 6 --warmup 5 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
python - <<P
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$TAG.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
by=collections.OrderedDict()
for r in rows[1:]:
    by.setdefault(r[ix["ID"]],{"name":r[ix["Kernel Name"]][:60]})[r[ix["Metric Name"]]]=r[ix["Metric Value"]]
for k,v in by.items():
    print(k, v["name"], " ".join("%s=%s"%(m.split(".")[0].replace("smsp__","").replace("sm__","")[:28],v[m]) for m in v if m!="name"))
P
