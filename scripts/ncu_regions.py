#!/usr/bin/env python
"""Where a warp of one kernel spends its time, by CODE REGION: executed instructions and stall samples of an
`ncu --set full --import-source on` report summed over address ranges (offsets from the kernel's first instruction, as
`cuobjdump -sass` prints them).

    python scripts/ncu_regions.py gpurun_out/prof_X.ncu-rep "[('refill',0x380,0x1900),('rhs',0x4000,0x6480), ...]"

(the attempt kernel's refill / stage states / right-hand side / error estimate / parking split.)"""
import csv,io,subprocess,sys,collections
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True,check=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    data.append((int(r[ix["Address"]],16), r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0), r))
base=data[0][0]
regions=eval(sys.argv[2])
stall_cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s=sum(d[3] for d in data); tot_n=sum(d[2] for d in data)
print("total samples",tot_s,"instr",tot_n)
for name,lo,hi in regions:
    sel=[d for d in data if lo<=d[0]-base<hi]
    s=sum(d[3] for d in sel); n=sum(d[2] for d in sel)
    st=collections.Counter()
    for d in sel:
        for c in stall_cols: st[c]+=int(d[4][ix[c]] or 0)
    print("%-14s instr %9d (%4.1f%%) samples %5d (%4.1f%%)  samples/kinstr %.3f  top: %s"%(name,n,100*n/tot_n,s,100*s/tot_s,1000*s/max(n,1),", ".join("%s %d"%(k[6:],v) for k,v in st.most_common(5))))
