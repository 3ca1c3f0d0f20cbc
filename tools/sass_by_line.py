"""SASS instructions of one kernel attributed to source lines (needs -lineinfo): which parts of the source the code bytes
come from, and where in the instruction stream (address ranges) each source region sits.

usage: python tools/sass_by_line.py <object or cubin> <kernel name substring> [bucket]"""
import re, subprocess, sys, os, tempfile
from collections import Counter, defaultdict

obj, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    obj = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "--print-line-info", obj], capture_output=True, text=True).stdout
secs = re.split(r"\n//-+ \.text\.", txt)
for s in secs[1:]:
    name = s.split(" ")[0]
    if pat not in name:
        continue
    cur, cnt, first, ops = None, Counter(), {}, defaultdict(Counter)
    n = 0
    for l in s.split("\n"):
        m = re.search(r'//## File ".*?([A-Za-z_0-9]+\.cuh?)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)) // bucket * bucket)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m and cur:
            cnt[cur] += 1
            ops[cur][m.group(3)] += 1
            first.setdefault(cur, int(m.group(1), 16))
            n += 1
    print(name[:70], "instructions", n, "bytes", n * 16)
    for k in sorted(cnt, key=lambda k: (k[0], k[1])):
        if cnt[k] >= 24:
            print(f"  {k[0]:18s} line {k[1]:5d}+  {cnt[k]:6d} instr  first at 0x{first[k]:05x}   {dict(ops[k].most_common(6))}")
    break
