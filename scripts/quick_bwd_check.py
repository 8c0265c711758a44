"""GPU: quick correctness + timing of the tcgen05 forward / reverse kernels (poisson net) -- used while tuning.
Always run under `timeout -s KILL`: a kernel that dead-locks cannot be interrupted otherwise."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipinn_b200 as mp  # noqa: E402
from oracle import autograd_ref as AR  # noqa: E402
from oracle import jet_ref as JR  # noqa: E402

torch.manual_seed(0)
m = mp.FNN(2, 1, [64] * 9).cuda()
e = m.jet_engine
order = int(os.environ.get("TC_ORDER", "3"))
lap = (1.0, 1.0) if order == 3 else None
for n in (300, 148 * 128 * 2 + 77):
    rng = np.random.default_rng(n)
    x_np = (rng.random((n, 2)) * 0.5).astype(np.float32)
    x = torch.tensor(x_np).cuda()
    out, stash = e.forward(x, order, True, lap=lap)
    dout_np = rng.standard_normal((n, out.shape[0], 1)).astype(np.float32)
    dflat = e.backward(x, order, torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda(), stash, lap=lap)
    torch.cuda.synchronize()
    prog, flat = AR.build_program(m)
    ref, cache = JR.forward(prog, flat, x_np.astype(np.float64), order, lap)
    rdp, _ = JR.backward(prog, flat, x_np.astype(np.float64), order, cache, dout_np.astype(np.float64), lap)
    eo = np.abs(out.cpu().numpy().transpose(1, 0, 2) - ref).max() / np.abs(ref).max()
    eg = np.abs(dflat.cpu().numpy() - rdp).max() / np.abs(rdp).max()
    print(f"n={n}: jets rel err {eo:.2e}  param-grad rel err {eg:.2e}", flush=True)
n = 1_000_000
x = torch.rand(n, 2, device="cuda") * 0.5
out, stash = e.forward(x, order, True, lap=lap)
dout = torch.randn_like(out)
for name, fn in (("fwd", lambda: e.forward(x, order, True, lap=lap, out=out, stash=stash)),
                 ("bwd", lambda: e.backward(x, order, dout, stash, lap=lap))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} 1M points order {order}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
