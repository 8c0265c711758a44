"""Diagnostic (GPU box): error of GPU jets / reverse pass vs the fp64 oracle, next to the error of the
same algebra evaluated in numpy fp32, for a full-depth network."""
import sys, os, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipinn_b200 as mp
from multipinn_b200 import _lib
from oracle import autograd_ref as AR, jet_ref as JR

eng = {"ffma": _lib.ENGINE_FFMA, "auto": _lib.ENGINE_AUTO}[sys.argv[1] if len(sys.argv) > 1 else "auto"]
torch.manual_seed(3)
model = mp.FNN(2, 1, [64] * 9).cuda()
model.jet_engine.kernel_engine = eng
rng = np.random.default_rng(5)
x_np = (rng.random((2048, 2)) * 0.5).astype(np.float32)
prog, flat = AR.build_program(model)
o64, c64 = JR.forward(prog, flat, x_np.astype(np.float64), 2)
o32, c32 = JR.forward(prog, flat.astype(np.float32), x_np, 2)
x = torch.tensor(x_np).cuda()
out, stash = model.jet_engine.forward(x, 2, need_stash=True)
og = out.cpu().numpy().transpose(1, 0, 2)
rel = lambda a, b: np.abs(a.astype(np.float64) - b).max() / np.abs(b).max()
for j, nm in enumerate(["u", "u_x", "u_y", "u_xx", "u_yy"]):
    print(f"{nm:5s} gpu {rel(og[:, j], o64[:, j]):.2e}   numpy-fp32 {rel(o32[:, j], o64[:, j]):.2e}")
dout = np.zeros_like(o64)
r = o64[:, 3, 0] + o64[:, 4, 0] + (20 * np.pi**2) * np.cos(2 * np.pi * x_np[:, 0]) * np.sin(4 * np.pi * x_np[:, 1])
dout[:, 3, 0] = dout[:, 4, 0] = 2 * r / 4096
g64, _ = JR.backward(prog, flat, x_np.astype(np.float64), 2, c64, dout)
g32, _ = JR.backward(prog, flat.astype(np.float32), x_np, 2, c32, dout.astype(np.float32))
gg = model.jet_engine.backward(x, 2, torch.tensor(np.ascontiguousarray(dout.transpose(1, 0, 2)), dtype=torch.float32).cuda(), stash).cpu().numpy()
print(f"grad  gpu {rel(gg, g64):.2e}   numpy-fp32 {rel(g32, g64):.2e}")
offs = [0]
for p in model.parameters():
    offs.append(offs[-1] + p.numel())
for k in range(len(offs) - 1):
    a, b = offs[k], offs[k + 1]
    print(f"  param {k:2d} [{a}:{b}] gpu {np.abs(gg[a:b]-g64[a:b]).max()/np.abs(g64).max():.2e} numpy-fp32 {np.abs(g32[a:b]-g64[a:b]).max()/np.abs(g64).max():.2e}")
