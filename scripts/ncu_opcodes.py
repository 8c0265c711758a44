"""Executed-instruction histogram by SASS opcode for one kernel of an .ncu-rep (needs --import-source on).
usage: python scripts/ncu_opcodes.py report.ncu-rep <ncu kernel regex> [top N]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ii, so, si = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
agg = collections.Counter(); st = collections.Counter(); seen = set()
for r in rows[2:]:
    if len(r) <= max(ii, so) or not r[ii].isdigit() or r[0] in seen:
        continue
    seen.add(r[0])
    toks = r[so].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDG", "STG", "LDS", "STS", "MUFU", "F2FP", "LDL", "STL", "BAR", "SYNCS")) else op.split(".")[0]
    agg[op] += int(r[ii]); st[op] += int(r[si])
tot = sum(agg.values()); ts = sum(st.values())
print(f"total warp instructions {tot}, stall samples {ts}")
for op, n in agg.most_common(topn):
    print(f"{100*n/tot:6.2f}% instr {100*st[op]/max(ts,1):6.2f}% stall  {op}")
