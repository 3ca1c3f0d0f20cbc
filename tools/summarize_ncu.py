"""Turn ncu output brought back from the GPU box into the text summaries kept under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> "<command line>"   -> per-kernel totals and shares
  python tools/summarize_ncu.py kernel <report.ncu-rep>                      -> key metrics + hottest SASS lines by stall samples"""
import csv, io, subprocess, sys
from collections import defaultdict


def launches(path, cmdline):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    kn, mv, un = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = {"ns": v * 1e-6, "us": v * 1e-3, "usecond": v * 1e-3, "nsecond": v * 1e-6, "ms": v, "msecond": v}.get(r[un], v * 1e-6)
        name = r[kn].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {cmdline}")
    print(f"(cold-cache, serialised launches: compare SHARES, not absolutes; {sum(cnt.values())} launches captured)")
    for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:20]:
        print(f"{name[:70]:70s} n={cnt[name]:5d} total={v:10.3f} ms share={v / total:6.3f} avg={v / cnt[name] * 1e3:9.1f} us")


def kernel(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit, r = rows[0], rows[1], rows[2]
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {r[i]} {unit[i]}")
    st = sorted(((float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                 if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")), reverse=True)
    print("warp stall reasons (warps per issue-active cycle):")
    for v, h in st[:7]:
        print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.2f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    si, sj, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    sec = [x for x in rows[2:] if len(x) > sj and x[0] != "Kernel Name"]
    total = sum(float(x[sj]) for x in sec)
    print(f"hottest SASS instructions by warp-stall samples ({int(total)} samples, {len(sec)} instructions):")
    for x in sorted(sec, key=lambda x: -float(x[sj]))[:10]:
        top = max(((float(x[hdr.index(h)]), h[6:]) for h in stalls))
        print(f"    {100 * float(x[sj]) / total:5.1f} %  executed {x[ie]:>9}  {x[si].strip()[:56]:56s} top stall: {top[1]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        kernel(sys.argv[2])
