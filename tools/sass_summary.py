#!/usr/bin/env python
"""SASS evidence per kernel of the shipped objects (cuobjdump -sass galaexi_b200/csrc/build/dgx_inst_N<N>.o): instruction counts of
the mnemonics that tell how a kernel moves and multiplies data -- DMMA (FP64 tensor core, mma.sync.m8n8k4), DFMA/DMUL/DADD,
LDG/STG, LDS/STS, LDGSTS (cp.async), UBLKCP (cp.async.bulk = TMA 1-D), SYNCS (mbarrier), BAR. usage: sass_summary.py 7 5 4"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "BAR", "MUFU"]
WANT = ("k_lifting", "k_sideflux", "k_volsurf2", "k_volsurf", "k_timestep", "k_overint", "k_source_rk", "k_filter")


def main():
    for N in sys.argv[1:] or ["7", "5"]:
        obj = os.path.join(ROOT, "galaexi_b200", "csrc", "build", f"dgx_inst_N{N}.o")
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "--dump-resource-usage", obj], capture_output=True, text=True).stdout
        regs = {}
        fn = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                fn = m.group(1)
            m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
            if m and fn:
                regs[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
        print(f"### N={N} (`dgx_inst_N{N}.o`, sm_100a)\n")
        print("| kernel | instr | regs | stack | " + " | ".join(KEYS) + " |")
        print("|---|---:|---:|---:|" + "---:|" * len(KEYS))
        cur, cnt, tot = None, None, 0
        rows = []

        def flush():
            if cur and any(w in cur for w in WANT):
                dem = subprocess.run(["cu++filt", cur], capture_output=True, text=True).stdout.strip() or cur
                dem = re.sub(r"\(dgx::KParams.*", "", dem).replace("void dgx::", "")
                r = regs.get(cur, (0, 0, 0))
                rows.append((dem, tot, r[0], r[1], [cnt[k] for k in KEYS]))
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                flush()
                cur, cnt, tot = m.group(1), collections.Counter(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur:
                tot += 1
                op = m.group(1)
                for k in KEYS:
                    if op == k or (k in ("LDG", "STG", "LDS", "STS") and op == k):
                        cnt[k] += 1
                if op.startswith("BAR"):
                    pass
        flush()
        for dem, tot, r, st, c in sorted(rows):
            if "ILi1ELi" in dem:
                pass
            print(f"| `{dem}` | {tot} | {r} | {st} | " + " | ".join(str(x) for x in c) + " |")
        print()


if __name__ == "__main__":
    main()
