#!/usr/bin/env python
"""registers / spills / stack per kernel instance of one degree:  tools/ptxas_res.py <N> [filter-regex] [-- extra nvcc flags]"""
import re, subprocess, sys, os
N = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "--" else ""
extra = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else []
csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "galaexi_b200", "csrc")
out = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                      "--expt-relaxed-constexpr", "-Xptxas", "-v", *extra, f"-DDGX_N={N}", "-c", "dgx_inst.cu", "-o", f"/tmp/ptxas_N{N}.o"],
                     cwd=csrc, capture_output=True, text=True).stderr
cur, rows = None, {}
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\w+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(dgx::KParams.*", "", cur).replace("void dgx::", "")
        rows[cur] = []
    elif cur and ("Used" in line or "spill" in line):
        rows[cur].append(re.sub(r"ptxas info\s*:\s*", "", line).strip())
for k in sorted(rows):
    if re.search(flt, k):
        print(k, "::", " | ".join(rows[k]))
