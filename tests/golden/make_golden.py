"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (``/root/reference``, read-only) is pure Python on PyTorch; its package
``__init__`` drags in plotting / Hydra / graph libraries that are not installed here and
that the hot path never touches, so permissive stand-in modules are registered for those
names before ``import multipinn``.  Everything numeric below is computed by reference
code: its ``neural_network`` classes, the residual closures of its ``examples/*/problem.py``,
its ``condition.diff.grad/unpack``, ``PINN.calculate_mean_loss`` and
``TrainerOneBatch._train_batches`` with ``torch.optim.Adam``.

The GPU box has no /root/reference: tests only read the .npz files written here.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Stub("stub_call")

    def __mro_entries__(self, bases):
        return (object,)


def install_stubs():
    for name in [
        "plotly", "plotly.graph_objects", "plotly.express", "plotly.subplots", "plotly.io",
        "hydra", "hydra.utils", "omegaconf", "skopt", "skopt.sampler",
        "torch_geometric", "torch_geometric.data", "torch_geometric.nn",
        "torch_geometric.nn.conv", "torch_geometric.nn.pool", "torch_geometric.nn.pool.glob",
        "torch_geometric.utils", "meshio", "matplotlib", "matplotlib.pyplot", "kaleido",
        "tqdm", "tqdm.auto",
    ]:
        if name in ("tqdm", "tqdm.auto"):
            continue  # installed
        sys.modules.setdefault(name, _Stub(name))


def import_reference():
    install_stubs()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import multipinn  # noqa: F401

    return multipinn


# golden cases: (name, reference example module, factory name, model class name, kwargs,
#                n_points per condition list).  Depth is reduced for all but the headline
#                config (poisson) so the fixtures stay small; the layer structure is unchanged.
CASES = [
    ("regression_1D", "examples.regression_1D.problem", "regression_1D_problem", "FNN",
     dict(hidden_layers=[64, 64, 64]), [64]),
    ("poisson_2D_1C", "examples.poisson_2D_1C.problem", "poisson_2D_1C", "FNN",
     dict(hidden_layers=[64] * 9), [96, 16, 16, 16, 16]),
    ("burgers_1D", "examples.burgers_1D.problem", "problem_burgers1D", "FNN",
     dict(hidden_layers=[128, 128]), [96, 32, 32]),
    ("allen_cahn_2D_1C", "examples.allen_cahn_2D_1C.problem", "problem_2D1C_Allen_Cahn", "FNN",
     dict(hidden_layers=[64, 64, 64]), [96, 32, 32, 32]),
    ("navier_stokes_2D", "examples.navier_stokes_2D_obstacle_multigpu.problem", "problem", "ResNet",
     dict(hidden_layers=[64, 64, 64]), [96, 32, 32, 32]),
    ("navier_stokes_3D", "examples.navier_stokes_3D_pipe.problem", "navier_stokes_3D_pipe", "FFNN",
     dict(hidden_layers=[64, 64, 64], encoding_dim=64, sigma=1.0), [96, 32, 32, 32]),
]


def _problems():
    """tests/problems.py by path (the reference has its own top-level ``tests`` package)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "b200_test_problems", os.path.join(ROOT, "tests", "problems.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def boxes_for(name):
    problems = _problems()

    class NS:  # only the boxes are taken from tests/problems.py
        grad = unpack = Symbols = None

    fn = problems.CONFIGS[name][0]
    conds, _, _ = fn(NS)
    return [box for _, box in conds]


class _FixedGenerator:
    """Stands in for multipinn.generation.Generator: hands back preset points."""

    def __init__(self, pts):
        self.pts = pts
        self.n_points = len(pts)

    def generate(self, condition, model):
        return self.pts


def run_case(mp, name, modname, factory, model_name, kwargs, counts):
    problems = _problems()

    torch.manual_seed(42)
    rng = np.random.default_rng(42)
    mod = importlib.import_module(modname)
    conds, din, dout = getattr(mod, factory)()
    Model = getattr(mp.neural_network, model_name)
    model = Model(input_dim=din, output_dim=dout, **kwargs)
    boxes = boxes_for(name)
    assert len(boxes) == len(conds), (name, len(boxes), len(conds))
    pts_np = [problems.sample_points(lo, hi, n, rng) for (lo, hi), n in zip(boxes, counts)]

    out = {"input_dim": din, "output_dim": dout}
    sd = model.state_dict()
    out["state_keys"] = np.array(list(sd.keys()))
    for k, v in sd.items():
        out["w::" + k] = v.detach().numpy().copy()
    for i, p in enumerate(pts_np):
        out[f"pts{i}"] = p

    from multipinn.condition.diff import grad, unpack

    def derivs(m, x):
        f = m(x)
        g = torch.stack([grad(f[:, k], x) for k in range(dout)], 1)  # [N,o,d]
        h = torch.stack(
            [torch.stack([grad(g[:, k, i], x)[:, i] for i in range(din)], 1) for k in range(dout)], 1
        )
        return f, g, h

    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        m = Model(input_dim=din, output_dim=dout, **kwargs).to(dtype)
        m.load_state_dict({k: v.to(dtype) for k, v in sd.items()})
        pts = [torch.tensor(p, dtype=dtype).requires_grad_() for p in pts_np]
        f, g, h = derivs(m, pts[0])
        out[f"u_{tag}"] = f.detach().numpy()
        out[f"du_{tag}"] = g.detach().numpy()
        out[f"d2u_{tag}"] = h.detach().numpy()
        # loss through the reference PINN (PINN.py:109-125)
        for c, p in zip(conds, pts):
            c.generator = _FixedGenerator(p)
            c.points = p
            c.batch_size = len(p)
        pinn = mp.PINN(model=m, conditions=conds) if dtype == torch.float32 else None
        if pinn is None:  # init_output_len builds an fp32 probe point; skip it for fp64
            pinn = mp.PINN.__new__(mp.PINN)
            pinn.model, pinn.conditions = m, conds
        pinn.select_batch(0)
        k = 0
        for ci, c in enumerate(conds):
            for r in c.get_residual(m):
                out[f"res{ci}_{k}_{tag}"] = r.detach().numpy()
                k += 1
        total, losses = pinn.calculate_mean_loss()
        legacy = torch.sum(torch.stack(pinn.calculate_loss()))
        m.zero_grad()
        total.backward()
        out[f"total_{tag}"] = total.detach().numpy()
        out[f"legacy_{tag}"] = legacy.detach().numpy()
        out[f"losses_{tag}"] = torch.stack(losses).detach().numpy()
        out[f"grad_{tag}"] = np.concatenate(
            [(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
             for p in m.parameters()]
        )
        if dtype == torch.float32:
            # three reference train steps (trainer.py:318-338) with Adam(1e-3)+ExponentialLR(.9999)
            opt = torch.optim.Adam(m.parameters(), lr=1e-3)
            sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
            tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=3, update_grid_every=10**9,
                                    calc_loss="mean",
                                    callbacks_organizer=mp.CallbacksOrganizer([]))
            traj = []
            for e in range(3):
                tr.current_epoch = e + 1  # never 0 mod update_grid_every -> keep preset points
                tr._train_epoch()
                sched.step()
                traj.append(float(tr.total_loss))
                if e == 0:
                    out["params_after1"] = np.concatenate(
                        [p.detach().numpy().reshape(-1) for p in m.parameters()])
                    out["losses_step1"] = tr.epoch_loss_detailed.numpy().copy()
            out["traj_f32"] = np.array(traj)
            out["params_after3"] = np.concatenate(
                [p.detach().numpy().reshape(-1) for p in m.parameters()])
    return out


def gradient_reg_case(mp, kind="gradient"):
    """Reference GradientLosses(alpha=0.9) (regularization/gradient.py:25-59), poisson_2D_1C's shipped default --
    or NormalLosses(alpha=0.9) (regularization/normal.py:6-45) / PointwiseLosses(lr=0.01)
    (regularization/pointwise.py:9-152): 3 TrainerOneBatch steps from the poisson golden weights/points; lambdas,
    losses and weights per step."""
    g = np.load(os.path.join(HERE, "poisson_2D_1C.npz"))
    mod = importlib.import_module("examples.poisson_2D_1C.problem")
    conds, din, dout = mod.poisson_2D_1C()
    model = mp.neural_network.FNN(input_dim=din, output_dim=dout, hidden_layers=[64] * 9)
    model.load_state_dict({k: torch.tensor(g["w::" + k]) for k in g["state_keys"]})
    for i, c in enumerate(conds):
        p = torch.tensor(g[f"pts{i}"]).requires_grad_()
        c.generator = _FixedGenerator(p)
        c.points, c.batch_size = p, len(p)
    pinn = mp.PINN(model=model, conditions=conds)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = {"gradient": lambda: mp.regularization.GradientLosses(alpha=0.9),
           "normal": lambda: mp.regularization.NormalLosses(alpha=0.9),
           "pointwise": lambda: mp.regularization.PointwiseLosses(ignored_indexes=[1], lr=0.01),
           # (a large lr so that three steps move the weights visibly)
           "adaptive": lambda: mp.regularization.AdaptiveConditionsLosses(lr=0.05),
           "adaptive_points": lambda: mp.regularization.AdaptiveConditionsAndPointsLosses(lr=0.05)}[kind]()
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg,
                            callbacks_organizer=mp.CallbacksOrganizer([]))
    out = {"lambdas": [], "totals": [], "losses": []}
    for e in range(3):
        # (epoch 0 regenerates the points -- the preset ones again -- which is what creates the point weights)
        tr.current_epoch = e if kind == "adaptive_points" else e + 1
        tr._train_epoch()
        if kind.startswith("adaptive"):
            lam = [reg.adaptive_cond.weight.detach().numpy().copy()]
            if kind == "adaptive_points":
                lam += [reg.adaptive_points[0].weight.detach().numpy()[:16], reg.adaptive_points[-1].weight.detach().numpy()[:16]]
            out["lambdas"].append(np.concatenate(lam))
        elif kind == "pointwise":   # per-point weights of residual 0 (a sample) and of the last residual
            ms = reg.models[0]
            out["lambdas"].append(np.concatenate([ms[0].weight.detach().numpy().reshape(-1)[:16],
                                                  ms[-1].weight.detach().numpy().reshape(-1)[:16]]))
        else:
            out["lambdas"].append(reg.lambda_regularization.detach().numpy().copy())
        out["totals"].append(float(tr.total_loss))
        out["losses"].append(tr.epoch_loss_detailed.numpy().copy())
    out = {k: np.array(v) for k, v in out.items()}
    out["params_after3"] = np.concatenate([p.detach().numpy().reshape(-1) for p in model.parameters()])
    return out


def long_run_case(mp, steps=1000):
    """1 000 free-running TrainerOneBatch steps of the UNMODIFIED reference from the poisson golden weights / points
    (Adam 1e-3, ExponentialLR 0.9999), once in fp32 and once in fp64: total-loss curves, and the fp64 weights after
    100 and 1 000 steps.  The fp64 curve is the truth the loss-curve parity tests measure against; the gap between
    the two curves is the reference's own fp32 envelope (the 1e-4 criterion is ill-posed beyond it)."""
    g = np.load(os.path.join(HERE, "poisson_2D_1C.npz"))
    out = {}
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        mod = importlib.import_module("examples.poisson_2D_1C.problem")
        conds, din, dout = mod.poisson_2D_1C()
        model = mp.neural_network.FNN(input_dim=din, output_dim=dout, hidden_layers=[64] * 9).to(dtype)
        model.load_state_dict({k: torch.tensor(g["w::" + k]).to(dtype) for k in g["state_keys"]})
        for i, c in enumerate(conds):
            p = torch.tensor(g[f"pts{i}"], dtype=dtype).requires_grad_()
            c.generator = _FixedGenerator(p)
            c.points, c.batch_size = p, len(p)
        if dtype == torch.float32:
            pinn = mp.PINN(model=model, conditions=conds)
        else:   # init_output_len builds an fp32 probe point; skip it for fp64
            pinn = mp.PINN.__new__(mp.PINN)
            pinn.model, pinn.conditions = model, conds
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
        tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=steps, update_grid_every=10**9, calc_loss="mean",
                                callbacks_organizer=mp.CallbacksOrganizer([]))
        curve = []
        for e in range(steps):
            tr.current_epoch = e + 1
            tr._train_epoch()
            sched.step()
            curve.append(float(tr.total_loss))
            if dtype == torch.float64 and (e + 1) in (100, steps):
                out[f"params_f64_after{e + 1}"] = np.concatenate(
                    [p.detach().numpy().reshape(-1) for p in model.parameters()])
        out[f"loss_{tag}"] = np.array(curve)
    return out


def long_run_family_case(mp, case, steps=300):
    """`steps` free-running TrainerOneBatch steps of the UNMODIFIED reference from a golden case's weights / points
    (Adam 1e-3, ExponentialLR 0.9999) in fp64 and fp32: the total-loss curves and the fp64 weights after 100 steps.
    Pins the oracle port over a long run for the residual-block and Fourier-feature families (the 1 000-step poisson
    fixture does it for the plain FNN)."""
    name, modname, factory, model_name, kwargs, _ = case
    g = np.load(os.path.join(HERE, name + ".npz"))
    out = {}
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        mod = importlib.import_module(modname)
        conds, din, dout = getattr(mod, factory)()
        model = getattr(mp.neural_network, model_name)(input_dim=din, output_dim=dout, **kwargs).to(dtype)
        model.load_state_dict({k: torch.tensor(g["w::" + k]).to(dtype) for k in g["state_keys"]})
        for i, c in enumerate(conds):
            p = torch.tensor(g[f"pts{i}"], dtype=dtype).requires_grad_()
            c.generator = _FixedGenerator(p)
            c.points, c.batch_size = p, len(p)
        if dtype == torch.float32:
            pinn = mp.PINN(model=model, conditions=conds)
        else:
            pinn = mp.PINN.__new__(mp.PINN)
            pinn.model, pinn.conditions = model, conds
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
        tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=steps, update_grid_every=10**9, calc_loss="mean",
                                callbacks_organizer=mp.CallbacksOrganizer([]))
        curve = []
        for e in range(steps):
            tr.current_epoch = e + 1
            tr._train_epoch()
            sched.step()
            curve.append(float(tr.total_loss))
            if dtype == torch.float64 and e + 1 == 100:
                out["params_f64_after100"] = np.concatenate(
                    [p.detach().numpy().reshape(-1) for p in model.parameters()])
        out[f"loss_{tag}"] = np.array(curve)
    return out


def activation_kats(mp):
    """The reference's own known-answer vectors (tests/neural_network/test_activation_function.py:8-59)."""
    from multipinn.neural_network.activation_function import fused_gelu, fused_sin

    x = torch.tensor([-3.0, -1.0, -0.5, 0.0, 0.5, 1.0, 2.0, 3.0])
    return {"x": x.numpy(), "gelu": fused_gelu(x).numpy(), "sin": fused_sin(x).numpy(),
            "kat_gelu_1": np.float32(0.84134474), "kat_gelu_m1": np.float32(-0.15865526)}


def main():
    mp = import_reference()
    np.savez(os.path.join(HERE, "activation_kats.npz"), **activation_kats(mp))
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        out = run_case(mp, *case)
        path = os.path.join(HERE, case[0] + ".npz")
        np.savez_compressed(path, **out)
        print(case[0], "total_f32", out["total_f32"], "total_f64", out["total_f64"],
              os.path.getsize(path) // 1024, "KiB")
    if not only or "poisson_1k" in only:
        out = long_run_case(mp)
        np.savez_compressed(os.path.join(HERE, "poisson_1k.npz"), **out)
        d = np.abs(out["loss_f32"] - out["loss_f64"]) / out["loss_f64"]
        print("poisson_1k: loss", out["loss_f64"][[0, 99, 999]], "fp32-vs-fp64 rel dev at 100/500/1000:",
              d[99], d[499], d[999], "max", d.max())
    for case in CASES:
        if case[0] in ("navier_stokes_2D", "navier_stokes_3D", "burgers_1D", "allen_cahn_2D_1C") and \
                (not only or case[0] + "_300" in only):
            out = long_run_family_case(mp, case)
            np.savez_compressed(os.path.join(HERE, case[0] + "_300.npz"), **out)
            d = np.abs(out["loss_f32"] - out["loss_f64"]) / out["loss_f64"]
            print(case[0] + "_300: loss", out["loss_f64"][[0, 99, 299]], "fp32-vs-fp64 rel dev max", d.max())
    for kind in ("gradient", "normal", "pointwise", "adaptive", "adaptive_points"):
        if not only or f"{kind}_reg" in only:
            out = gradient_reg_case(mp, kind)
            np.savez_compressed(os.path.join(HERE, f"poisson_{kind}_reg.npz"), **out)
            print(f"{kind}_reg lambdas", out["lambdas"][-1][:6], "totals", out["totals"])


if __name__ == "__main__":
    main()
