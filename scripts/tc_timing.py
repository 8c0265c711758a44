"""GPU diagnostic (needs a build with EXTRA=-DMPINN_TC_TIMING=1): cycle breakdown of the tcgen05 reverse kernel."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipinn_b200 as mp
from multipinn_b200 import _lib, engine
torch.manual_seed(0)
m = mp.FNN(2, 1, [64] * 9).cuda()
e = m.jet_engine
n = 148 * 128 * 8
x = torch.rand(n, 2, device="cuda") * 0.5
ORDER = int(os.environ.get("TC_ORDER", "3"))
LAP = (1.0, 1.0) if ORDER == 3 else None
out, stash = e.forward(x, ORDER, True, lap=LAP)
dout = torch.randn_like(out)
lib = _lib.load()
buf = (ctypes.c_longlong * 32)()
lib.mpinn_debug_counters(buf, 32)
e.backward(x, ORDER, dout, stash, lap=LAP)
lib.mpinn_debug_counters(buf, 32)
names = {15: "tile start", 0: "output op", 1: "layer setup (coeffs, cache fill, W)", 2: "phase: loads + zbar chunk0", 3: "phase: wait prev GEMMs",
         4: "phase: rest of staging", 5: "phase: barrier", 6: "phase: MMA issue", 7: "layer: final GEMM wait", 8: "layer: dW flush", 9: "op0 + tail"}
for th, label in ((0, "thread 0 (issuer)"), (1, "thread 224")):
    vals = [buf[th * 16 + i] for i in range(16)]
    tot = sum(vals)
    print(label, "total cycles", tot, " per tile", tot / 8)
    for k, nm in names.items():
        print(f"   {nm:40s} {vals[k]:12d}  {100*vals[k]/max(tot,1):5.1f}%   per tile-layer {vals[k]/8/8:9.0f}")

# ---- forward
buf = (ctypes.c_longlong * 64)()
lib.mpinn_debug_counters(buf, 64)
out, stash = e.forward(x, ORDER, True, lap=LAP)
lib.mpinn_debug_counters(buf, 64)
names = {15: "tile start", 10: "layer start (W prefetch, bias)", 11: "channel: source + activation math + stash, chunk 0",
         12: "channel: wait A buffer", 13: "channel: rest of staging", 14: "channel: barrier", 6: "channel: MMA issue",
         7: "layer: end GEMM wait", 9: "output op"}
print("FORWARD")
for th, label in ((0, "thread 0 (issuer)"), (1, "thread 224")):
    vals = [buf[32 + th * 16 + i] for i in range(16)]
    tot = sum(vals)
    print(label, "total cycles", tot, " per tile", tot / 8)
    for k, nm in names.items():
        print(f"   {nm:52s} {vals[k]:12d}  {100*vals[k]/max(tot,1):5.1f}%   per tile-layer {vals[k]/8/8:9.0f}")
