#!/bin/bash
# env-kernel instruction / local-memory counts of a short bench run (ncu metrics pass)
for pd in ${PREDRAWS:-1 0}; do
FWGYM_ENV_PREDRAW=$pd timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,launch__registers_per_thread \
  --clock-control none -k regex:fw_env_kernel -s 210 -c 2 --csv --log-file gpurun_out/envcounts_$pd.csv python bench.py --steps 6 --warmup 5 --no-cpu-baseline --e2e-steps 5 > /dev/null 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/envcounts_$pd.csv")) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
out={}
for r in rows[1:]:
    out.setdefault(r[ix["ID"]],{})[r[ix["Metric Name"]]]=r[ix["Metric Value"]]
for k,v in out.items(): print("predraw=$pd", k, v)
P
done
