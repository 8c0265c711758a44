"""Attribute ncu stall samples / executed instructions to CUDA source lines.

ncu's CSV source page is SASS-only; this joins it (by instruction offset) with `nvdisasm -g` line info
taken from the cubin inside the CURRENT libmpinn_b200.so, so the report must come from the same build.

usage: python scripts/ncu_lines.py report.ncu-rep <ncu kernel regex> <mangled-name substring> [top N]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "multipinn_b200", "csrc", "libmpinn_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line = None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and cur_fn and mangled in cur_fn:
            line_of[int(m.group(1), 16)] = cur_line
if not line_of:
    sys.exit("no function matching " + mangled)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
si, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
d = []
for r in rows[2:]:
    if len(r) <= max(si, ii) or not r[si].isdigit():
        continue
    if d and d[-1][0] == r[0]:
        continue
    d.append(r)
base = int(d[0][0], 16)
agg = collections.defaultdict(lambda: [0, 0])
for r in d:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    agg[key][0] += int(r[si])
    agg[key][1] += int(r[ii])
tot_s = sum(v[0] for v in agg.values())
tot_i = sum(v[1] for v in agg.values())
srcs = {}
def text(fn, ln):
    if fn not in srcs:
        try:
            srcs[fn] = open(os.path.join(root, "multipinn_b200", "csrc", fn)).read().splitlines()
        except OSError:
            srcs[fn] = []
    L = srcs[fn]
    return L[ln - 1].strip()[:95] if 0 < ln <= len(L) else ""
print(f"total stall samples {tot_s}, warp instructions {tot_i}")
print(" stall%  instr%  file:line")
for (fn, ln), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*s/tot_s:6.2f}  {100*i/tot_i:6.2f}  {fn}:{ln}  {text(fn, ln)}")
