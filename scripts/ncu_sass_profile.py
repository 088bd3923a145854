#!/usr/bin/env python
"""Dynamic instruction mix + stall samples of one kernel from an `ncu --set full --import-source on` report.

    python scripts/ncu_sass_profile.py gpurun_out/prof_X.ncu-rep [top_n]

Prints executed warp-instructions per opcode, the stall-sample totals per reason, and the instructions with the most
samples (address, opcode text, samples, dominant stall)."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samples = collections.Counter(), collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stall_tot = collections.Counter()
    lines = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        src = r[ix["Source"]].strip()
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0].rstrip(";")
        n = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        ops[op] += n
        samples[op] += s
        st = {c: int(r[ix[c]] or 0) for c in stall_cols}
        for c, v in st.items():
            stall_tot[c] += v
        lines.append((s, n, r[ix["Address"]], src, max(st, key=st.get) if s else ""))
    tot = sum(ops.values())
    print("executed warp instructions: %d" % tot)
    for op, n in ops.most_common(28):
        print("  %-10s %10d  %5.1f %%   samples %6d" % (op, n, 100.0 * n / tot, samples[op]))
    ts = sum(stall_tot.values())
    print("stall samples: %d" % ts)
    for c, v in stall_tot.most_common(12):
        print("  %-24s %8d  %5.1f %%" % (c, v, 100.0 * v / max(1, ts)))
    print("hottest instructions:")
    for s, n, addr, src, why in sorted(lines, reverse=True)[:top]:
        print("  %6d samples  %9d exec  %s  %-60s %s" % (s, n, addr[-5:], src[:60], why))


if __name__ == "__main__":
    main()
