"""Per-kernel share of a timed step from an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python scripts/launch_shares.py launches.csv <first launch id> <last launch id> [title]"""
import collections, csv, sys
path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
title = sys.argv[4] if len(sys.argv) > 4 else path
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
agg = collections.defaultdict(lambda: [0.0, 0])
n = 0
for r in rows[1:]:
    if lo <= int(r[ii]) <= hi:
        agg[r[ki]][0] += float(r[vi].replace(",", "")) / 1e6
        agg[r[ki]][1] += 1
        n += 1
tot = sum(v[0] for v in agg.values())
print(f"ncu launch list (gpu__time_duration, cold-cache serialised): {title}, launches {lo}..{hi}")
for k, (ms, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{ms:10.3f} ms  {100*ms/tot:5.2f}%  x{c:4d}  {k[:100]}")
print(f"total {tot:.3f} ms over {n} launches")
