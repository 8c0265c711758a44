"""Free-running loss-curve comparison over 1000 Adam steps (GPU box): this package vs the reference's algorithm
(nested torch.autograd, oracle/autograd_ref.py, fp32 and fp64) from the same weights and points.
poisson_2D_1C, 20 000 + 4 x 1 000 points, Adam(1e-3) + ExponentialLR(0.9999), calc_loss="mean"."""
import json, os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipinn_b200 as mp
from oracle import autograd_ref as AR
from tests import problems

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
torch.manual_seed(42)
rng = np.random.default_rng(42)
ns_p = types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack, Symbols=mp.Symbols)
ns_r = types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None)
spec_p, din, dout = problems.poisson_2D_1C(ns_p)
spec_r, _, _ = problems.poisson_2D_1C(ns_r)
pts = [problems.sample_points(lo, hi, 20000 if i == 0 else 1000, rng) for i, (_, (lo, hi)) in enumerate(spec_p)]
model = mp.FNN(din, dout, [64] * 9).cuda()
sd = {k: v.clone() for k, v in model.state_dict().items()}

def run_reference(dtype):
    ref = AR.FNN(din, dout, [64] * 9).to(dtype).cuda()
    ref.load_state_dict({k: v.to(dtype) for k, v in sd.items()})
    P = [torch.tensor(p, dtype=dtype).cuda().requires_grad_() for p in pts]
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    out = []
    for _ in range(STEPS):
        total, _ = AR.train_step(ref, opt, [f for f, _ in spec_r], P)
        sch.step()
        out.append(float(total))
    return np.array(out)

conds = []
for (fn, (lo, hi)), p in zip(spec_p, pts):
    c = mp.Condition(fn, mp.Hypercube(lo, hi))
    t = torch.tensor(p).cuda().requires_grad_()
    c.generator = types.SimpleNamespace(n_points=len(p), generate=lambda cond, model, t=t: t)
    conds.append(c)
pinn = mp.PINN(model, conds)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
sch = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
tr = mp.TrainerOneBatch(pinn, opt, sch, num_epochs=STEPS, update_grid_every=10**9)
pinn.update_data()
ours = []
for e in range(STEPS):
    tr.current_epoch = e + 1
    tr._train_epoch()
    sch.step()
    ours.append(float(tr.total_loss))
ours = np.array(ours)
r32, r64 = run_reference(torch.float32), run_reference(torch.float64)
rel = lambda a, b: np.abs(a - b) / np.abs(b)
rows = []
for k in (0, 9, 99, 299, 499, 749, STEPS - 1):
    if k < STEPS:
        rows.append({"step": k + 1, "loss_fp64_ref": r64[k], "ours_vs_fp64": rel(ours, r64)[k],
                     "ref_fp32_vs_fp64": rel(r32, r64)[k], "ours_vs_ref_fp32": rel(ours, r32)[k]})
summary = {"steps": STEPS, "max_ours_vs_fp64": float(rel(ours, r64).max()), "max_ref32_vs_fp64": float(rel(r32, r64).max()),
           "first_step_ours_vs_fp64": float(rel(ours, r64)[0]), "rows": rows,
           "steps_ours_within_1e-4_of_fp64": int(np.argmax(rel(ours, r64) > 1e-4)) if (rel(ours, r64) > 1e-4).any() else STEPS,
           "steps_ref32_within_1e-4_of_fp64": int(np.argmax(rel(r32, r64) > 1e-4)) if (rel(r32, r64) > 1e-4).any() else STEPS}
print(json.dumps(summary, indent=1, default=float))
