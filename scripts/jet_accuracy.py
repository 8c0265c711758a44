"""GPU diagnostic: accuracy of the jets (u, grad u, Laplacian) of the poisson network against the fp64 oracle, at
random-init weights and at the trained golden trajectory's scale -- used to compare library variants
(MPINN_B200_LIB=...)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipinn_b200 as mp  # noqa: E402
from oracle import autograd_ref as AR  # noqa: E402
from oracle import jet_ref as JR  # noqa: E402

for seed in (0, 1, 2):
    torch.manual_seed(seed)
    m = mp.FNN(2, 1, [64] * 9).cuda()
    with torch.no_grad():
        for p in m.parameters():   # a trained network has larger weights than the default init
            p.mul_(1.0 + 0.5 * seed)
    e = m.jet_engine
    rng = np.random.default_rng(seed)
    x_np = (rng.random((4096, 2)) * 0.5).astype(np.float32)
    x = torch.tensor(x_np).cuda()
    prog, flat = AR.build_program(m)
    for order, lap in ((2, None), (3, (1.0, 1.0))):
        out, _ = e.forward(x, order, True, lap=lap)
        ref, _ = JR.forward(prog, flat, x_np.astype(np.float64), order, lap)
        o = out.cpu().numpy().transpose(1, 0, 2).astype(np.float64)
        errs = [np.abs(o[:, j] - ref[:, j]).max() / np.abs(ref[:, j]).max() for j in range(ref.shape[1])]
        print(f"seed {seed} order {order}: per-channel max-normalised rel err " + " ".join(f"{v:.2e}" for v in errs))
