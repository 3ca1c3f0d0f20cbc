"""SASS instruction histogram per kernel of a built library (cuobjdump -sass): code bytes, instruction count, and the counts of the
mnemonics that prove what a kernel runs on (UTCHMMA / UTCBAR / LDTM / STTM = tcgen05 + tensor memory, SYNCS = mbarrier, LDGSTS =
cp.async, UTMALDG = TMA, HMMA = legacy mma.sync, FFMA / FFMA2 / DFMA / MUFU = CUDA-core math, LDL / STL = local-memory spills).

usage: python tools/sass_histogram.py <lib.so> [kernel-name substring ...]   ->  stdout (profiles/rNN_sass_histogram.txt)"""
import re
import subprocess
import sys
from collections import Counter

lib, pats = sys.argv[1], sys.argv[2:]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "SYNCS", "LDGSTS", "UTMALDG", "HMMA", "FFMA", "FFMA2", "FMUL2", "FADD2", "DFMA", "MUFU", "SHFL",
       "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "F2FP"]
rows = []
for block in re.split(r"\n\s*Function : ", txt)[1:]:
    name = block.split("\n", 1)[0].strip()
    if pats and not any(p in name for p in pats):
        continue
    ops = Counter()
    n = 0
    for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", block):
        ops[m.group(1)] += 1
        n += 1
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    rows.append((n, demangled, ops))
rows.sort(reverse=True)
print(f"{lib}: {len(rows)} kernels; 16 bytes per SASS instruction")
for n, name, ops in rows:
    shown = "  ".join(f"{k} {ops[k]}" for k in KEY if ops[k])
    print(f"\n{name}\n    {n} instructions = {16 * n / 1024:.1f} KB    {shown}")
