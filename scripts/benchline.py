"""Condense a bench.py JSON line (stdin) to the few numbers watched while tuning."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    if not line.startswith("{"):
        print(line.rstrip()[:200])
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print(f"ms/step {d['ms_per_step']:.2f}  pts/s {d['value']:.3e}  e2e {d['e2e']['value']:.3e}  "
          f"bwd {r.get('avg_launch_ms')}  fwd {r.get('fwd_avg_launch_ms')}  frac_tf32 {r.get('frac_of_tf32_peak')}  "
          f"clocks {d.get('clocks')}  launches {d.get('gpu_launches')}")
    for k in ("cpu_baseline", "ref_torch_cuda"):
        if k in d:
            print(" ", k, d[k])
