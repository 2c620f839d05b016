#!/usr/bin/env python
"""Hot spots of an ncu --set full capture with --import-source on: stall samples per SASS instruction, grouped into the
regions between BAR.SYNC instructions (the kernel phases), plus the top instructions.
usage: ncu_hot.py report.ncu-rep [topN] [kernel-name-substring]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
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
        tot = sum(int(r[si] or 0) for r in body)
        print(f"## {name}: {len(body)} SASS instructions, {tot} samples")
        seg, acc, start, nexec = 0, 0, 0, 0
        for n_, r in enumerate(body):
            acc += int(r[si] or 0); nexec += int(r[ei] or 0)
            if "BAR.SYNC" in r[1] or n_ == len(body) - 1:
                print(f"  region {seg}: instr {start}-{n_}  samples {acc} ({100.0 * acc / max(tot, 1):.1f}%)  warp-instr executed {nexec}")
                seg += 1; acc = 0; start = n_ + 1; nexec = 0
        order = sorted(range(len(body)), key=lambda x: -int(body[x][si] or 0))[:top]
        for x in sorted(order):
            print(f"  #{x:5d} {int(body[x][si]):6d} ({100.0 * int(body[x][si]) / max(tot, 1):.1f}%)  {body[x][1].strip()}")
        break
    k += 1
