"""Summarise an .ncu-rep: headline metrics per kernel and the hottest SASS instructions with context.
usage: python scripts/ncu_hot.py report.ncu-rep [kernel-regex] [threshold]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.006
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread"]
stall = [n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[h.index("Kernel Name")][:60]
    print("=====", name)
    for k in keys:
        if k in h:
            print(f"  {k} [{u[h.index(k)]}] = {r[h.index(k)]}")
    st = sorted(((float(r[h.index(n)]), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for n in stall), reverse=True)
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))
if kre:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    si = hdr.index("Warp Stall Sampling (All Samples)"); ii = hdr.index("Instructions Executed")
    d = []
    for r in rows[2:]:
        if len(r) <= max(si, ii) or not r[si].isdigit():
            continue
        if d and d[-1][0] == r[0]:
            continue
        d.append(r)
    tot = sum(int(r[si]) for r in d)
    print("---- hot SASS (share of stall samples), total", tot)
    hot = [i for i, r in enumerate(d) if int(r[si]) > thr * tot]
    last = -10
    for i in hot:
        lo = max(last + 1, i - 5); hi = min(len(d), i + 2)
        if lo > last + 1:
            print("....")
        for k in range(lo, hi):
            r = d[k]
            print(f"{100*int(r[si])/tot:5.2f}% ex={int(r[ii]):8d} {r[1].strip()[:100]}")
        last = hi - 1
