"""Warp-stall samples and executed instructions of one profiled kernel, attributed to SOURCE LINES: joins the SASS page of an
ncu report (per-address samples) with nvdisasm's line table of the same object (needs -lineinfo).

usage: python tools/ncu_by_line.py <report.ncu-rep> <object or cubin> <kernel name substring> [bucket lines] [min share %]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict

rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 10
min_share = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
else:
    cubins = [obj]
line_of = {}
for cb in cubins:
    txt = subprocess.run(["nvdisasm", "--print-line-info", cb], capture_output=True, text=True).stdout
    for s in re.split(r"\n//-+ \.text\.", txt)[1:]:
        if pat not in s.split(" ")[0]:
            continue
        cur = None
        for l in s.split("\n"):
            m = re.search(r'//## File ".*?([A-Za-z_0-9]+\.cuh?)", line (\d+)', l)
            if m:
                cur = (m.group(1), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+", l)
            if m and cur:
                line_of[int(m.group(1), 16)] = cur
        break
    if line_of:
        break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {k: hdr.index(k) for k in hdr}
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
base = None
samples, execd, stalls = Counter(), Counter(), defaultdict(Counter)
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    key = line_of.get(a - base, ("?", 0))
    key = (key[0], key[1] // bucket * bucket)
    samples[key] += int(r[ci["# Samples"]])
    execd[key] += int(r[ci["Instructions Executed"]])
    for k in stall_cols:
        stalls[key][k[6:]] += int(r[ci[k]])
tot_s, tot_e = sum(samples.values()), sum(execd.values())
print(f"{rep}: {tot_s} samples, {tot_e} warp instructions executed; by source region ({bucket}-line buckets, >= {min_share} % of samples or instructions)")
for k in sorted(samples, key=lambda k: (k[0], k[1])):
    if 100.0 * samples[k] / tot_s >= min_share or 100.0 * execd[k] / tot_e >= min_share:
        top = ", ".join(f"{n} {100.0 * v / max(samples[k], 1):.0f}%" for n, v in stalls[k].most_common(3))
        print(f"  {k[0]:18s} {k[1]:5d}+  samples {100.0 * samples[k] / tot_s:5.1f} %   instr {100.0 * execd[k] / tot_e:5.1f} %   [{top}]")
