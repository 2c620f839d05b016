#!/usr/bin/env python
"""Opcode mix (warp-instructions executed) per barrier-delimited region of a kernel in an ncu --set full capture.
usage: ncu_opmix.py report.ncu-rep kernel-substring [region ...]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
want = sys.argv[2]
regs = [int(x) for x in sys.argv[3:]]
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]; hdr = rows[k + 1]; k += 2
        body = []
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            body.append(rows[k]); k += 1
        if want not in name:
            continue
        si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
        seg = 0
        mix = collections.defaultdict(lambda: collections.Counter())
        smp = collections.defaultdict(lambda: collections.Counter())
        for r in body:
            ops = r[1].strip().split()
            op = ops[1] if ops[0].startswith("@") else ops[0]
            op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL")) and "." in op else "")
            mix[seg][op] += int(r[ei] or 0)
            smp[seg][op] += int(r[si] or 0)
            if "BAR.SYNC" in r[1]:
                seg += 1
        for sg in sorted(mix):
            if regs and sg not in regs:
                continue
            tot = sum(mix[sg].values()); ts = sum(smp[sg].values())
            print(f"region {sg}: {tot} warp-instr, {ts} samples")
            for op, c in mix[sg].most_common(14):
                print(f"    {op:12s} {c:12d} {100.0 * c / max(tot, 1):5.1f}%   samples {100.0 * smp[sg][op] / max(ts, 1):5.1f}%")
        break
    k += 1
