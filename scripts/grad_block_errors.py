import os, sys, types
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import tests.test_gpu_parity as T
from tests.test_gpu_parity import *
from multipinn_b200 import _lib

def run(name, eng, reps):
    model, din, dout = T.full_model(name, 3, eng)
    kind = problems.CONFIGS[name][1]
    ref = AR.MODELS[kind](din, dout, **problems.CONFIGS[name][2]).double().cuda()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    rng = np.random.default_rng(5)
    spec_p, _, _ = problems.CONFIGS[name][0](product_ns())
    spec_r, _, _ = problems.CONFIGS[name][0](types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None))
    pts = [problems.sample_points(lo, hi, 2048 if i == 0 else 512, rng) for i, (_, (lo, hi)) in enumerate(spec_p)]
    conds = []
    for (fn, (lo, hi)), p in zip(spec_p, pts):
        c = mp.Condition(fn, mp.Hypercube(lo, hi))
        t = torch.tensor(p).cuda().requires_grad_()
        c.generator = types.SimpleNamespace(n_points=len(p), generate=lambda cond, model, t=t: t)
        conds.append(c)
    pinn = mp.PINN(model, conds)
    pinn.update_data()
    for c in conds:
        c.set_batching(1)
    pinn.select_batch(0)
    rtotal, _ = AR.mean_loss(ref, [fn for fn, _ in spec_r],
                             [torch.tensor(p, dtype=torch.float64).cuda().requires_grad_() for p in pts])
    rtotal.backward()
    rg = [(p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy().reshape(-1) for p in ref.parameters()]
    rgrad = np.concatenate(rg)
    scale = np.abs(rgrad).max()
    sizes = [len(g) for g in rg]
    offs = np.cumsum([0] + sizes)
    grads = []
    for r in range(reps):
        model.jet_engine.zero_grad()
        total, losses = pinn.calculate_mean_loss()
        total.backward()
        torch.cuda.synchronize()
        grads.append(model.jet_engine.flat_grad.cpu().numpy().astype(np.float64).copy())
    G = np.stack(grads)
    print(f"engine {eng}: overall err per rep", [f"{np.abs(g - rgrad).max() / scale:.2e}" for g in grads])
    for i in range(len(sizes)):
        sl = slice(offs[i], offs[i + 1])
        err = np.abs(G[:, sl] - rgrad[sl]).max(axis=1) / scale
        spread = (G[:, sl].max(axis=0) - G[:, sl].min(axis=0)).max() / scale
        print(f"  param {i:2d} n={sizes[i]:5d} |ref|max/scale {np.abs(rgrad[sl]).max() / scale:.2e} err min {err.min():.2e} max {err.max():.2e} run-to-run spread {spread:.2e}")

for eng in (_lib.ENGINE_TCGEN05, _lib.ENGINE_FFMA):
    run("poisson_2D_1C", eng, 6)
