"""dram bytes per launch of the domain condition's forward / reverse kernels from an `ncu --set full` report
(scripts/gpu_ncu.sh), written as the table bench.py reads for roofline.traffic:
    python scripts/ncu_traffic.py gpurun_out/r2_prof.ncu-rep poisson_2D_1C:1000000 >> (merged into) profiles/r2_ncu_traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, key = sys.argv[1], sys.argv[2]
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ncu_traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]


def val(r, name):
    v = float(r[h.index(name)].replace(",", ""))
    unit = u[h.index(name)].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


tab = json.load(open(out_path)) if os.path.exists(out_path) else {}
ent = tab.setdefault(key, {})
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    kind = "bwd" if "jet_bwd" in name else ("fwd" if "jet_fwd" in name else None)
    if kind:
        ent[kind] = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        ent[kind + "_kernel"] = name.split("(")[0].replace("void ", "")
        ent[kind + "_duration_ms_under_ncu"] = float(r[h.index("gpu__time_duration.sum")].replace(",", "")) * \
            {"msecond": 1.0, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6}.get(u[h.index("gpu__time_duration.sum")].lower(), 1.0)
ent["source"] = os.path.basename(rep) + " (ncu --set full, one launch each, scripts/gpu_ncu.sh)"
json.dump(tab, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(ent, indent=1))
