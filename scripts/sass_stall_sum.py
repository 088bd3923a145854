#!/usr/bin/env python
"""Static issue-time estimate of a stretch of SASS: the sum of the stall counts ptxas put into the control words.

    python scripts/sass_stall_sum.py LIB.so KERNEL_SUBSTRING LO HI [-v]        # one address range (hex offsets)
    python scripts/sass_stall_sum.py LIB.so KERNEL_SUBSTRING --branches LO HI  # list the branches (to find the regions)
    python scripts/sass_stall_sum.py LIB.so KERNEL_SUBSTRING --path "[(mult, lo, hi), ...]"   # weighted sum over a path

Every sm_100a instruction is 128 bits; bits 105-108 of it hold the number of cycles the issuing warp must wait before its
NEXT instruction (fixed-latency dependencies: 8 cycles behind a dependent DFMA, 2 behind an independent FP64 instruction,
~4 between shared-memory instructions ...).  Variable-latency results (LDS / LDC / MUFU / LDG) are tracked by scoreboards
instead and are NOT in the sum, so it is a lower bound of a lone warp's time - but for the attempt kernel, whose two warps
per scheduler mostly wait on their own fixed latencies, differences of this sum predicted the measured A/B differences
(K pairs -690 cycles per pass -> -4 us per launch; output sink -90 cycles -> no change)."""
import collections
import re
import subprocess
import sys


def load(lib, kernel):
    names = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    cands = sorted(set(re.findall(r"\b(_Z\w*%s\w*)\b" % re.escape(kernel), names)), key=len)
    cands = [c for c in cands if "_param_" not in c]
    if not cands:
        sys.exit("no kernel matching %r" % kernel)
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", cands[0], lib], capture_output=True, text=True).stdout
    pat = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
    pat2 = re.compile(r"/\* (0x[0-9a-f]{16}) \*/")
    lines, ins, i = txt.split("\n"), [], 0
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            c = int(pat2.search(lines[i + 1]).group(1), 16) >> 41
            ins.append((int(m.group(1), 16), m.group(2).strip(), c & 0xf, (c >> 5) & 7, (c >> 11) & 0x3f))
            i += 2
        else:
            i += 1
    return cands[0], ins


def main():
    lib, kernel, args = sys.argv[1], sys.argv[2], sys.argv[3:]
    name, ins = load(lib, kernel)
    print("#", name)
    if args[0] == "--path":
        tot = 0
        for mult, lo, hi in eval(args[1]):
            tot += mult * sum(x[2] for x in ins if lo <= x[0] < hi)
        print("weighted stall sum:", tot)
        return
    if args[0] == "--branches":
        lo, hi = int(args[1], 16), int(args[2], 16)
        for a, t, st, _, _ in ins:
            if lo <= a < hi and re.search(r"\b(BRA|BRX|BSSY|BSYNC|BREAK|EXIT)\b", t):
                print("%05x st=%2d %s" % (a, st, t))
        return
    lo, hi = int(args[0], 16), int(args[1], 16)
    sel = [x for x in ins if lo <= x[0] < hi]
    print("instructions %d, sum of stall counts %d, instructions that wait on a scoreboard %d"
          % (len(sel), sum(x[2] for x in sel), sum(1 for x in sel if x[4])))
    by, cnt = collections.Counter(), collections.Counter()
    for a, t, st, wr, wm in sel:
        op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
        by[op] += st
        cnt[op] += 1
    for op, st in by.most_common(16):
        print("  %-8s n=%4d stall=%5d avg=%.2f" % (op, cnt[op], st, st / cnt[op]))
    if "-v" in args:
        for a, t, st, wr, wm in sel:
            print("%05x st=%2d wr=%d wait=%02x  %s" % (a, st, wr, wm, t))


if __name__ == "__main__":
    main()
