#!/usr/bin/env python
"""Split a kernel's executed instructions and stall samples (ncu --set full --import-source on) by HOW OFTEN each
instruction ran: the straight-line step path runs in (nearly) every warp, rare paths (auto-resets) in a few.

    python scripts/ncu_exec_split.py gpurun_out/prof_env_X.ncu-rep

(the env kernel's reset path was 49 % of its stall samples before the cooperative reset, 28 % after.)"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [(r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0))
            for r in rows[2:] if len(r) >= len(hdr)]
    full = max(d[1] for d in data)
    tot_s = sum(d[2] for d in data)
    print("instructions in the listing %d, most-executed instruction %d times, stall samples %d" % (len(data), full, tot_s))
    classes = [("step path (>= 25 % of the warps)", lambda n: n >= 0.25 * full),
               ("rare paths (< 25 % of the warps)", lambda n: 0 < n < 0.25 * full),
               ("never executed", lambda n: n == 0)]
    for name, f in classes:
        sel = [d for d in data if f(d[1])]
        print("  %-34s static %6d  executed %9d  samples %6d (%4.1f %%)"
              % (name, len(sel), sum(d[1] for d in sel), sum(d[2] for d in sel), 100.0 * sum(d[2] for d in sel) / max(1, tot_s)))
    rare = [d for d in data if 0 < d[1] < 0.25 * full]
    cnt = collections.Counter(d[1] for d in rare)
    (n_exec, n_static), = cnt.most_common(1)
    print("  dominant rare path: %d instructions executed by %d warps each" % (n_static, n_exec))
    ops = collections.Counter()
    for t, n, _ in rare:
        ops[(t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]] += n
    print("  its opcodes (executed / warp): " + ", ".join("%s %d" % (k, v // max(1, n_exec)) for k, v in ops.most_common(10)))


if __name__ == "__main__":
    main()
