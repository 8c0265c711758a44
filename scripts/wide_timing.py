"""One forward + reverse pass of the burgers network (4x128) on the layer-wise tcgen05 engine; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel durations, or plain for CUDA-event totals."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import multipinn_b200 as mp
from multipinn_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
m = mp.FNN(2, 1, [128, 128, 128, 128]).cuda()
e = m.jet_engine
e.kernel_engine = _lib.ENGINE_TCGEN05
x = torch.rand((n, 2), device="cuda")
out = torch.empty((4, n, 1), device="cuda")
stash = torch.empty(e.stash_floats(3, n), device="cuda")
dout = torch.randn((4, n, 1), device="cuda")
dflat = torch.zeros_like(e.flat_grad) if e.flat_grad is not None else None
for it in range(reps):
    t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0.record()
    e.forward(x, 3, True, lap=(1.0, 0.0), out=out, stash=stash)
    t1.record()
    if dflat is None:
        e.ensure_flat()
        dflat = torch.zeros_like(e.flat_grad)
    e.backward_into(x, 3, dout, stash, dflat, lap=(1.0, 0.0))
    t2.record()
    torch.cuda.synchronize()
    print(f"n={n} fwd {t0.elapsed_time(t1):.3f} ms  bwd {t1.elapsed_time(t2):.3f} ms")
