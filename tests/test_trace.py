"""Residual tracing (multipinn_b200/trace.py), the Laplace-mode rewrite and the fused train step
(multipinn_b200/fused.py), exercised on CPU with the oracle standing in for the kernels and checked against the
vectors of the unmodified reference (tests/golden)."""
import types

import numpy as np
import pytest
import torch

import multipinn_b200 as mp
from multipinn_b200 import engine, jet, trace
from multipinn_b200.fused import ConditionRunner
from oracle import jet_ref as JR
from oracle import residual_ref as RR
from tests import problems
from tests.oracle_backend import OracleBackend
from tests.test_host_cpu import build_pinn, product_model, product_ns
from tests.test_oracle_golden import NAMES, load, relerr


@pytest.fixture(autouse=True)
def oracle_backend():
    be = OracleBackend()
    old = engine._set_backend_for_tests(be)
    jet.clear_blocks()
    yield be
    engine._set_backend_for_tests(old)


# what the tracer must find for the five BASELINE configs: (order, Laplace weights) of the domain condition
EXPECT_DOMAIN = {
    "regression_1D": (0, None),
    "poisson_2D_1C": (3, (1.0, 1.0)),
    "burgers_1D": (3, (1.0, 0.0)),
    "allen_cahn_2D_1C": (3, (1.0, 0.0)),
    "navier_stokes_2D": (3, (1.0, 1.0)),
    "navier_stokes_3D": (3, (1.0, 1.0, 1.0)),
}


@pytest.mark.parametrize("name", NAMES)
def test_every_example_condition_is_traced(name):
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    for c in conds:
        assert c.program_for(model) is not None, c.trace_note
    blk = conds[0].program_for(model).blocks[0]
    order, lap = EXPECT_DOMAIN[name]
    assert blk.order == order
    assert (blk.lap if order == 3 else None) == lap


@pytest.mark.parametrize("name", NAMES)
def test_traced_residuals_match_reference_golden(name):
    """values of the device program == the residual vectors of the unmodified reference (fp64 goldens)."""
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    k = 0
    for ci, c in enumerate(conds):
        prog = c.program_for(model)
        vals = ConditionRunner(c, prog, model.jet_engine).residual_values(torch.tensor(g[f"pts{ci}"]))
        for kk in range(vals.shape[0]):
            assert relerr(vals[kk].numpy(), g[f"res{ci}_{k}_f64"]) < 1e-5, (ci, k)
            k += 1


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["mean", "legacy"])
def test_fused_step_loss_and_gradient_match_reference(name, mode, oracle_backend):
    """One fused step (no autograd): losses and the flat parameter gradient equal the reference's
    ``calculate_mean_loss`` / ``backward`` (goldens); the eager path gives the same numbers."""
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)   # not Adam: the step leaves the gradient untouched
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=1, update_grid_every=10**9, calc_loss=mode)
    assert tr.fused_step() is not None
    oracle_backend.calls.clear()
    pinn.select_batch(0)
    tr._train_batches()
    kinds = [c[0] for c in oracle_backend.calls]
    assert "resid" in kinds and "bwd" in kinds
    assert relerr(tr.epoch_loss_detailed.numpy(), g["losses_f64"]) < 1e-5
    want_total = g["total_f64"] if mode == "mean" else g["legacy_f64"]
    assert relerr(tr.total_loss.numpy(), want_total) < 1e-5
    if mode == "mean":
        assert relerr(model.jet_engine.flat_grad.numpy(), g["grad_f64"]) < 1e-5
    # and the eager jet path agrees with the fused one
    fused_grad = model.jet_engine.flat_grad.clone()
    tr.use_fused_step = False
    tr.total_loss, tr.epoch_loss_detailed = 0, None
    tr._train_batches()
    assert relerr(model.jet_engine.flat_grad.numpy(), fused_grad.numpy()) < 1e-6
    assert relerr(tr.total_loss.numpy(), want_total) < 1e-5


@pytest.mark.parametrize("name", ["poisson_2D_1C", "navier_stokes_2D", "regression_1D"])
def test_fused_adam_trajectory_matches_reference(name):
    """3 TrainerOneBatch steps with Adam through the fused path (device-side lr / step state) land on the
    reference's weights (goldens params_after1 / params_after3)."""
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.999)
    tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=2, update_grid_every=10**9)
    assert tr.fused_step() is not None and tr._flat_adam is not None
    flat = model.jet_engine.flat
    for step in range(3):
        tr.current_epoch = step
        tr._train_epoch()
        sched.step()
        if step == 0:
            assert relerr(flat.detach().numpy(), g["params_after1"]) < 2e-5
            assert relerr(tr.epoch_loss_detailed.numpy(), g["losses_step1"]) < 1e-5
    assert relerr(flat.detach().numpy(), g["params_after3"]) < 5e-5
    assert float(opt.state[next(model.parameters())]["step"]) == 3.0


def _trace(fn, model, d=2):
    return trace.trace_residual(fn, model, d, torch.zeros(1, d))


def _small_model(d=2, o=1):
    torch.manual_seed(0)
    return mp.FNN(d, o, [16, 16])


def test_laplace_rewrite_cases():
    model = _small_model()

    def sym(model, arg):
        (u,) = mp.unpack(model(arg))
        x, y = mp.unpack(arg)
        u_x, u_y = mp.unpack(mp.grad(u, arg))
        u_xx, _ = mp.unpack(mp.grad(u_x, arg))
        _, u_yy = mp.unpack(mp.grad(u_y, arg))
        return u, x, y, u_x, u_y, u_xx, u_yy

    # anisotropic constant coefficients: one channel with weights (1, 0.25) up to scale
    p = _trace(lambda m, a: [(lambda s: 4.0 * s[5] + s[6] - s[0])(sym(m, a))], model)
    assert p.blocks[0].order == 3 and np.allclose(p.blocks[0].lap, (1.0, 0.25))
    # two components with proportional combinations share the channel
    p = _trace(lambda m, a: (lambda s: [s[5] + s[6], -3.0 * (s[5] + s[6]) / 2.0 + s[3]])(sym(m, a)), model)
    assert p.blocks[0].order == 3 and p.blocks[0].lap == (1.0, 1.0)
    # non-proportional combinations, non-constant coefficients and non-linear uses keep the full order-2 block
    p = _trace(lambda m, a: (lambda s: [s[5] + s[6], s[5] - s[6]])(sym(m, a)), model)
    assert p.blocks[0].order == 2
    p = _trace(lambda m, a: (lambda s: [s[1] * s[5] + s[6]])(sym(m, a)), model)
    assert p.blocks[0].order == 2
    p = _trace(lambda m, a: (lambda s: [s[5] * s[5] + s[6]])(sym(m, a)), model)
    assert p.blocks[0].order == 2
    p = _trace(lambda m, a: (lambda s: [torch.sin(s[5] + s[6])])(sym(m, a)), model)
    assert p.blocks[0].order == 2
    # derivatives that are requested but never used do not raise the block's order
    p = _trace(lambda m, a: (lambda s: [s[0] - s[1]])(sym(m, a)), model)
    assert p.blocks[0].order == 0
    p = _trace(lambda m, a: (lambda s: [s[3] * s[0]])(sym(m, a)), model)
    assert p.blocks[0].order == 1


def _eval(prog, model, x, lap_check=True):
    """Program values on the fp64 oracle jets."""
    ops, flat = model.jet_program(), model.jet_engine
    flat.ensure_flat()
    params = flat.flat.detach().numpy().astype(np.float64)
    jets = []
    for b, blk in enumerate(prog.blocks):
        xb = x
        pp = prog.points_program(b)
        if pp is not None:
            vals, _, _ = RR.evaluate(pp.instrs, pp.consts, pp.outputs, [], x, [], want_seeds=False)
            xb = np.ascontiguousarray(vals.T)
        o, _ = JR.forward(ops, params, xb, blk.order, blk.lap)
        jets.append(o.transpose(1, 0, 2))
    return RR.evaluate(prog.instrs, prog.consts, prog.outputs, [b.o for b in prog.blocks], x, jets)


def test_laplace_mode_equals_full_second_order():
    """The rewritten program on Laplace-mode jets == the eager residual on full order-2 jets, values and seeds
    mapped back (d L / d u_ii = c_i d L / d L-channel)."""
    model = _small_model()
    rng = np.random.default_rng(1)
    x = rng.random((33, 2))

    def fn(m, a):
        (u,) = mp.unpack(m(a))
        xx, yy = mp.unpack(a)
        u_x, u_y = mp.unpack(mp.grad(u, a))
        u_xx, _ = mp.unpack(mp.grad(u_x, a))
        _, u_yy = mp.unpack(mp.grad(u_y, a))
        return [0.5 * u_xx + 2.0 * u_yy + u * u_x - torch.exp(xx) * yy, u_y**2 - (u_xx * 0.25 + u_yy)]

    prog = _trace(fn, model)
    assert prog.blocks[0].order == 3 and np.allclose(prog.blocks[0].lap, (0.25, 1.0))
    vals, sums, dj = _eval(prog, model, x)
    with mp.jet_order(2):
        xt = torch.tensor(x, dtype=torch.float32).requires_grad_()
        eager = fn(model, xt)
    for k in range(2):
        assert relerr(vals[k], eager[k].detach().numpy().astype(np.float64)) < 1e-6


def test_multi_block_programs_shifted_points():
    g = load("regression_1D")
    model, conds, pinn = build_pinn("regression_1D", g)
    prog = conds[0].program_for(model)
    assert len(prog.blocks) == 2 and prog.blocks[0].points is None and prog.blocks[1].points is not None
    pp = prog.points_program(1)
    x = np.linspace(0, 1, 5).reshape(-1, 1)
    vals, _, _ = RR.evaluate(pp.instrs, pp.consts, pp.outputs, [], x, [], want_seeds=False)
    assert np.allclose(vals[0], x[:, 0] + 1.0)
    g = load("burgers_1D")
    model, conds, pinn = build_pinn("burgers_1D", g)
    prog = conds[1].program_for(model)     # periodic boundary: model(arg) and model(arg shifted by 2 in x)
    assert len(prog.blocks) == 2
    pp = prog.points_program(1)
    x = np.array([[-1.0, 0.3], [-1.0, 0.7]])
    vals, _, _ = RR.evaluate(pp.instrs, pp.consts, pp.outputs, [], x, [], want_seeds=False)
    assert np.allclose(vals.T, x + np.array([2.0, 0.0]))


def test_untraceable_residuals_fall_back_to_eager(oracle_backend):
    model = _small_model()
    geo = mp.Hypercube([0, 0], [1, 1])

    def branching(m, a):
        (u,) = mp.unpack(m(a))
        return [u if bool((u > 0).all()) else -u]          # data-dependent control flow

    def batch_reduction(m, a):
        (u,) = mp.unpack(m(a))
        return [u - u.mean()]                               # couples the points of a batch

    def derived_grad(m, a):
        (u,) = mp.unpack(m(a))
        return [u * 2.0]

    for fn in (branching, batch_reduction):
        c = mp.Condition(fn, geo)
        c.init_output_len(model)
        assert c.program_for(model) is None and c.trace_note
        assert c.output_len == 1
    c = mp.Condition(derived_grad, geo)
    c.init_output_len(model)
    assert c.program_for(model) is not None
    # a trainer with an untraced condition stays on the eager path and still trains
    c1, c2 = mp.Condition(batch_reduction, geo), mp.Condition(derived_grad, geo)
    for c in (c1, c2):
        mp.Generator(16, "pseudo").use_for(c)
    pinn = mp.PINN(model, [c1, c2])
    tr = mp.TrainerOneBatch(pinn, torch.optim.Adam(model.parameters(), lr=1e-3), None, num_epochs=1)
    assert tr.fused_step() is None
    tr._train_epoch()
    assert np.isfinite(float(tr.total_loss))


def test_mixed_derivative_selects_full_hessian_mode_only_when_used():
    model = _small_model()

    def unused(m, a):
        (u,) = mp.unpack(m(a))
        u_x, u_y = mp.unpack(mp.grad(u, a))
        u_xx, u_xy = mp.unpack(mp.grad(u_x, a))
        return [u_xx - u]

    def used(m, a):
        (u,) = mp.unpack(m(a))
        u_x, u_y = mp.unpack(mp.grad(u, a))
        u_xx, u_xy = mp.unpack(mp.grad(u_x, a))
        u_yx, u_yy = mp.unpack(mp.grad(u_y, a))
        return [u_xy * u_yx + u_xx + u_yy + u]

    assert _trace(unused, model).blocks[0].order == 3      # requested but dead: still one Laplace channel
    p = _trace(used, model)
    assert p.blocks[0].order == 4                          # value, 2 first, 2 diagonal, 1 mixed
    loads = sorted(ins[2] for ins in p.instrs if ins[0] == trace.OP_LOADJ)
    assert loads == [0, 3, 4, 5]                           # u, u_xx, u_yy, u_xy (u_yx is the same channel)
    # values against the eager residual (which now escalates its plan to order 4 instead of raising)
    rng = np.random.default_rng(2)
    x = rng.random((17, 2))
    vals, _, _ = _eval(p, model, x)
    c = mp.Condition(used, mp.Hypercube([0, 0], [1, 1]))
    c.init_output_len(model)
    assert c.plan_for(model).orders == [4] and c.program_for(model) is not None
    eager = c.get_residual_fn(model)(torch.tensor(x, dtype=torch.float32).requires_grad_())
    assert relerr(vals[0], eager[0].detach().numpy().astype(np.float64)) < 1e-6


def test_full_hessian_fused_step_matches_autograd(oracle_backend):
    """A residual with a mixed derivative through the fused step (order-4 jets, FP32-FMA engine on the device):
    loss and parameter gradient against nested torch autograd of the reference algorithm (oracle port, fp64)."""
    from oracle import autograd_ref as AR
    torch.manual_seed(1)
    model = mp.FNN(2, 1, [16, 16])

    def make(ns):
        def fn(m, a):
            (u,) = ns.unpack(m(a))
            x, y = ns.unpack(a)
            u_x, u_y = ns.unpack(ns.grad(u, a))
            u_xx, u_xy = ns.unpack(ns.grad(u_x, a))
            _, u_yy = ns.unpack(ns.grad(u_y, a))
            return [u_xx * u_yy - u_xy**2 - torch.exp(x) * y, u_x + u_xy]   # Monge-Ampere-like + a linear use
        return fn

    cond = mp.Condition(make(types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack)), mp.Hypercube([0, 0], [1, 1]))
    pts = torch.rand((40, 2), generator=torch.Generator().manual_seed(0)).requires_grad_()
    cond.generator = types.SimpleNamespace(n_points=40, generate=lambda c, m: pts)
    pinn = mp.PINN(model, [cond])
    pinn.update_data()
    tr = mp.TrainerOneBatch(pinn, torch.optim.SGD(model.parameters(), lr=0.0), None, num_epochs=1,
                            update_grid_every=10**9)
    assert tr.fused_step() is not None and cond.program_for(model).blocks[0].order == 4
    pinn.select_batch(0)
    tr._train_batches()
    ref = AR.MODELS["FNN"](2, 1, hidden_layers=[16, 16]).double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    rfn = make(types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack))
    total, losses = AR.mean_loss(ref, [rfn], [pts.detach().double().requires_grad_()])
    total.backward()
    rgrad = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                            for p in ref.parameters()])
    assert relerr(tr.epoch_loss_detailed.numpy(), losses.detach().numpy()) < 1e-5
    assert relerr(float(tr.total_loss), float(total)) < 1e-5
    assert relerr(model.jet_engine.flat_grad.numpy(), rgrad) < 1e-5


def test_program_struct_roundtrip_and_limits():
    g = load("navier_stokes_3D")
    model, conds, pinn = build_pinn("navier_stokes_3D", g)
    prog = conds[0].program_for(model)
    st = prog.struct()
    assert st.n_instr == len(prog.instrs) and st.n_out == 4 and st.n_blocks == 1
    assert st.block_J[0] == 5 and st.block_o[0] == 4     # Laplace mode: value, 3 first derivatives, L
    loads = [(i, ins) for i, ins in enumerate(prog.instrs) if ins[0] == trace.OP_LOADJ]
    for i, (_, a, b, _) in loads:
        assert st.jet_reg[a][b] == i
    model2 = _small_model()

    def long(m, a):
        (u,) = mp.unpack(m(a))
        for _ in range(200):
            u = u * 1.5 + 0.25
        return [u]

    with pytest.raises(trace.TraceError):
        _trace(long, model2)


def test_same_order_conditions_share_one_forward_and_reverse_launch(oracle_backend):
    """poisson: the four value-only Dirichlet conditions are batched -- 2 forward, 5 residual, 2 reverse launches per
    step instead of 5 / 5 / 5 -- and the numbers stay those of the reference."""
    g = load("poisson_2D_1C")
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    tr = mp.TrainerOneBatch(pinn, torch.optim.SGD(model.parameters(), lr=0.0), None, num_epochs=1,
                            update_grid_every=10**9)
    fs = tr.fused_step()
    assert [type(u).__name__ for u in fs.units] == ["ConditionRunner", "BatchedGroup"]
    oracle_backend.calls.clear()
    pinn.select_batch(0)
    tr._train_batches()
    kinds = [c[0] for c in oracle_backend.calls]
    assert kinds.count("fwd") == 2 and kinds.count("bwd") == 2 and kinds.count("resid") == 5
    assert [c[2] for c in oracle_backend.calls if c[0] == "fwd"] == [96, 64]
    assert relerr(tr.epoch_loss_detailed.numpy(), g["losses_f64"]) < 1e-5
    assert relerr(model.jet_engine.flat_grad.numpy(), g["grad_f64"]) < 1e-5


@pytest.mark.parametrize("kind", ["constant", "gradient", "normal"])
def test_regularizer_policies_run_on_the_fused_step(kind, oracle_backend):
    """ConstantLosses / GradientLosses / NormalLosses as loss policies of the fused step (K per-condition gradient
    slabs from one reverse launch per condition, statistics and the weighted combination on the device) give the
    eager jet path's numbers -- which tests/test_host_cpu.py pins to runs of the unmodified reference."""
    g = load("poisson_2D_1C")
    results = []
    for fused in (True, False):
        model, conds, pinn = build_pinn("poisson_2D_1C", g)
        reg = {"constant": lambda: mp.ConstantLosses([1.0, 2.0, 3.0, 4.0, 5.0]),
               "gradient": lambda: mp.GradientLosses(alpha=0.9), "normal": lambda: mp.NormalLosses(alpha=0.9)}[kind]()
        tr = mp.TrainerOneBatch(pinn, torch.optim.Adam(model.parameters(), lr=1e-3), None, num_epochs=3,
                                update_grid_every=10**9, calc_loss=reg)
        tr.use_fused_step = fused
        assert (tr.fused_step() is not None) == fused
        if fused:
            assert tr.fused_step().mode == kind
        out = []
        for e in range(3):
            tr.current_epoch = e + 1
            tr._train_epoch()
            lam = reg.lambda_regularization.numpy().copy() if kind != "constant" else np.zeros(1)
            out.append((float(tr.total_loss), tr.epoch_loss_detailed.numpy().copy(), lam))
        results.append((out, model.jet_engine.flat.detach().numpy().copy()))
    (fo, fw), (eo, ew) = results
    for (ft, fl, flam), (et, el, elam) in zip(fo, eo):
        assert relerr(ft, et) < 1e-6 and relerr(fl, el) < 1e-6
        if kind != "constant":
            assert relerr(flam, elam) < 1e-5
    assert np.abs(fw - ew).max() < 1e-6


def test_model_calls_of_one_residual_are_stacked_into_one_launch(oracle_backend):
    """regression_1D (u(x) * u(x + 1) - x) and burgers' periodic boundary call the model twice: one forward and one
    reverse launch over the 2n stacked points, the shifted points produced on the device."""
    for name, ci in (("regression_1D", 0), ("burgers_1D", 1)):
        g = load(name)
        model, conds, pinn = build_pinn(name, g)
        tr = mp.TrainerOneBatch(pinn, torch.optim.SGD(model.parameters(), lr=0.0), None, num_epochs=1,
                                update_grid_every=10**9)
        fs = tr.fused_step()
        assert fs.runners[ci].stacked
        n = len(g[f"pts{ci}"])
        oracle_backend.calls.clear()
        pinn.select_batch(0)
        tr._train_batches()
        fwd = [c for c in oracle_backend.calls if c[0] == "fwd"]
        assert (fwd[ci][2] if name == "regression_1D" else sorted(c[2] for c in fwd)[-1]) == 2 * n or \
            any(c[2] == 2 * n for c in fwd)
        assert sum(1 for c in oracle_backend.calls if c[0] == "bwd") == len(fs.units)
        assert relerr(tr.epoch_loss_detailed.numpy(), g["losses_f64"]) < 1e-5
        assert relerr(model.jet_engine.flat_grad.numpy(), g["grad_f64"]) < 1e-5


# ----------------------------------------------------------------------------------------------------------------
# idioms of the reference's other examples (scripts/example_sweep.py runs all 46 problem.py files; these are the
# constructs that needed tracer support, pinned here so that they stay on the fused step)
def _fused_vs_autograd(make, d, n_pts, data_fn=None, hidden=(16, 16), seed=3, shift=None):
    """Fused step (traced program, oracle backend) vs nested torch autograd of the reference algorithm in fp64."""
    from oracle import autograd_ref as AR
    torch.manual_seed(seed)
    model = mp.FNN(d, 1, list(hidden))
    ns = types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack)
    geo = mp.Hypercube([0.0] * d, [1.0] * d)
    if data_fn is None:
        cond = mp.Condition(make(ns), geo)
    else:
        cond = mp.ConditionExtra(make(ns), geo, [data_fn])
    pts = torch.rand((n_pts, d), generator=torch.Generator().manual_seed(seed)).requires_grad_()
    cond.generator = types.SimpleNamespace(n_points=n_pts, generate=lambda c, m: pts)
    pinn = mp.PINN(model, [cond])
    pinn.update_data()
    tr = mp.TrainerOneBatch(pinn, torch.optim.SGD(model.parameters(), lr=0.0), None, num_epochs=1,
                            update_grid_every=10**9)
    assert cond.program_for(model) is not None, cond.trace_note
    assert tr.fused_step() is not None
    pinn.select_batch(0)
    tr._train_batches()
    ref = AR.MODELS["FNN"](d, 1, hidden_layers=list(hidden)).double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    rns = types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack)
    p64 = pts.detach().double().requires_grad_()
    if data_fn is None:
        rfn = make(rns)
    else:
        dat = tuple(t.double() for t in (data_fn(pts.detach()),))
        rfn = lambda m, a: make(rns)(m, a, dat)   # noqa: E731
    total, losses = AR.mean_loss(ref, [rfn], [p64])
    total.backward()
    rgrad = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                            for p in ref.parameters()])
    assert relerr(tr.epoch_loss_detailed.numpy(), losses.detach().numpy()) < 1e-5
    assert relerr(model.jet_engine.flat_grad.numpy(), rgrad) < 1e-5
    return cond.program_for(model)


def test_tracer_handles_the_idioms_of_the_reference_examples():
    def make(ns):
        def fn(m, a):
            (u,) = ns.unpack(m(a))
            x, y = ns.unpack(a)
            # examples/heat_2D_simple/problem.py:31 -- a real 0-d tensor on the LEFT of a traced column
            c1 = torch.exp(torch.tensor(1)) + 0.5 * y - u
            # examples/poisson_3D/problem.py:38 -- real tensor * traced column
            c2 = torch.square(torch.tensor(torch.pi)) * torch.cos(torch.pi * x) - u
            # examples/laplace_2D_neumann/problem.py:35-43 -- constants sized by the batch, row-wise sum
            normals = torch.tensor([-1, 0]).repeat(a.shape[0], 1)
            u_n = torch.sum(ns.grad(u, a) * normals, 1)
            pi_t = torch.ones_like(y) * np.pi
            c3 = u_n + (np.pi / (torch.exp(pi_t) - torch.exp(-pi_t))) * (torch.exp(pi_t * y) - torch.exp(-pi_t * y))
            # examples/navier_stokes_3D_pipe_marching/problem.py:84 -- a constant component shaped after a column
            c4 = torch.tensor(np.zeros_like(u.cpu().detach().numpy()), dtype=u.dtype)
            return [c1, c2, c3, c4]
        return fn

    prog = _fused_vs_autograd(make, 2, 24)
    assert len(prog.blocks) == 1 and prog.blocks[0].order == 1


def test_residuals_that_draw_random_numbers_stay_eager():
    model = _small_model()

    def fn(m, a):   # num_diff_random's per-point random step (reference condition/diff.py:84-111)
        f = m(a)
        (d_x,) = mp.unpack(mp.condition.num_diff_random(m, a, f, torch.tensor([[1.0, 0.0]])))
        return [d_x]

    def fn_np(m, a):
        (u,) = mp.unpack(m(a))
        return [u - 1e-9 * float(np.random.rand())]          # far below the probe tolerance: only the RNG check sees it

    for f in (fn, fn_np):
        c = mp.Condition(f, mp.Hypercube([0, 0], [1, 1]))
        c.init_output_len(model)
        assert c.program_for(model) is None and "random" in c.trace_note


def test_derivative_of_a_call_on_shifted_points():
    """examples/portal_1D_1C/problem.py:33-39: ``ur = model(arg + delta)`` then ``grad(ur, arg)``."""
    def make(ns):
        def fn(m, a):
            (ul,) = ns.unpack(m(a))
            (ur,) = ns.unpack(m(a + 2 * torch.pi))
            (ul_x,) = ns.unpack(ns.grad(ul, a))
            (ur_x,) = ns.unpack(ns.grad(ur, a))
            (ur_xx,) = ns.unpack(ns.grad(ur_x, a))
            return [ur - ul, ur_x - ul_x, ur_xx + ur]
        return fn

    prog = _fused_vs_autograd(make, 1, 20)
    assert len(prog.blocks) == 2 and prog.blocks[1].points is not None
    # the eager jet path answers the same (it is what the probe check compared the program with)
    torch.manual_seed(0)
    model = mp.FNN(1, 1, [16, 16])
    x = torch.rand(7, 1).requires_grad_()
    (ur,) = mp.unpack(model(x + 0.5))
    (ur_x,) = mp.unpack(mp.grad(ur, x))
    from oracle import autograd_ref as AR
    ref = AR.MODELS["FNN"](1, 1, hidden_layers=[16, 16]).double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    x64 = x.detach().double().requires_grad_()
    (rr,) = AR.unpack(ref(x64 + 0.5))
    (rr_x,) = AR.unpack(AR.grad(rr, x64))
    assert relerr(ur_x.detach().numpy(), rr_x.detach().numpy()) < 1e-5
    with pytest.raises(jet.UnsupportedDerivative):     # a scaled argument is not a shift
        (us,) = mp.unpack(model(x * 2.0))
        mp.grad(us, x)


def test_condition_extra_data_is_traced_as_per_point_columns(oracle_backend):
    """ConditionExtra (reference condition/condition.py:117-215): the residual takes per-point data -- here a
    normal field -- which the program reads through LOADD from a packed [N, n_data] block."""
    def normals(points):
        p = points.detach()
        v = torch.stack([torch.cos(3.0 * p[:, 0]), torch.sin(2.0 * p[:, 1]) + 0.5], dim=1)
        return v / v.norm(dim=1, keepdim=True)

    def make(ns):
        def fn(m, a, data):
            (u,) = ns.unpack(m(a))
            normal = data[0]
            u_x, u_y = ns.unpack(ns.grad(u, a))
            return [normal[:, 0] * u_x + normal[:, 1] * u_y - 0.3]     # examples/helmholtz/problem.py:26-29
        return fn

    prog = _fused_vs_autograd(make, 2, 28, data_fn=normals)
    assert prog.data_width == 2 and prog.struct().n_data == 2
    assert any(op == trace.OP_LOADD for op, *_ in prog.instrs)
    assert any(c[0] == "resid" for c in oracle_backend.calls)


def test_masks_select_and_max_min_are_traced():
    """Comparison masks, masked assignment (examples/gas_dynamics_1D/problem.py:33-37), torch.where, maximum /
    minimum / clamp: fused loss and gradient against autograd of the same residual."""
    def make(ns):
        def fn(m, a):
            (u,) = ns.unpack(m(a))
            x, y = ns.unpack(a)
            p = 0.5 * torch.exp(-x / 0.7)
            p[p > 0.35] = 0                                        # masked assignment on a per-point column
            w = torch.where((y < 0.5) & ~(x > 0.8) | (u > 10.0), u * 2.0, u - 1.0)    # piecewise residual, combined masks
            q = p.clone()
            q[q < 0.3] = 7.0                                       # must not reach p
            c = torch.clamp(u, min=-0.05, max=0.08) + torch.maximum(u, 0.1 * x) - torch.minimum(u, y - 0.5)
            return [u - p, w, c]
        return fn

    prog = _fused_vs_autograd(make, 2, 64)
    ops = {i[0] for i in prog.instrs}
    assert {trace.OP_GT, trace.OP_SEL, trace.OP_MAX, trace.OP_MIN} <= ops
    # a mask used as a Python truth value is still data-dependent control flow
    model = _small_model()

    def branching(m, a):
        (u,) = mp.unpack(m(a))
        return [u if (u > 0) else -u]

    c = mp.Condition(branching, mp.Hypercube([0, 0], [1, 1]))
    c.init_output_len(model)
    assert c.program_for(model) is None and "branches" in c.trace_note


def test_changed_residual_constants_are_followed(oracle_backend):
    """GradualTrainer (reference trainer/gradual_trainer.py:23-40) changes a problem parameter through a `nonlocal`
    setter while training (examples/convection_1D_gradual_learning/problem.py:16-19).  The traced program has the old
    value baked in: the closure fingerprint must trigger a re-trace, the fused step must take the new program over in
    place, and the numbers must equal a trainer built from scratch with the new value."""
    def problem(betta):
        def inner(m, a):
            (p,) = mp.unpack(m(a))
            p_t, p_x = mp.unpack(mp.grad(p, a))
            return [p_t + betta * p_x]

        def setter(v):
            nonlocal betta
            betta = v
        return inner, setter

    geo = mp.Hypercube([0, 0], [1, 1])
    pts = torch.rand((32, 2), generator=torch.Generator().manual_seed(3)).requires_grad_()

    def trainer_for(fn, model):
        cond = mp.Condition(fn, geo)
        cond.generator = types.SimpleNamespace(n_points=32, generate=lambda c, m: pts)
        pinn = mp.PINN(model, [cond])
        pinn.update_data()
        tr = mp.TrainerOneBatch(pinn, torch.optim.Adam(model.parameters(), lr=1e-3), None, num_epochs=5,
                                update_grid_every=10**9)
        return tr, cond

    torch.manual_seed(0)
    model = mp.FNN(2, 1, [16, 16])
    state = {k: v.clone() for k, v in model.state_dict().items()}
    fn, setter = problem(1.0)
    tr, cond = trainer_for(fn, model)
    fs = tr.fused_step()
    assert fs is not None
    tr.current_epoch = 1
    tr._train_epoch()
    prog0 = cond.program_for(model)
    setter(7.5)                                   # what GradualTrainer.train does between epochs
    tr.current_epoch = 2
    tr._train_epoch()
    assert tr.fused_step() is fs                  # same executor, new program
    assert cond.program_for(model) is not prog0 and 7.5 in cond.program_for(model).consts
    loss_after_change = float(tr.total_loss)
    # reference run: identical first step with betta = 1, then a fresh trainer that has betta = 7.5 from the start
    torch.manual_seed(0)
    model2 = mp.FNN(2, 1, [16, 16])
    model2.load_state_dict(state)
    fn2, setter2 = problem(1.0)
    tr2, _ = trainer_for(fn2, model2)
    tr2.use_fused_step = False                    # eager path evaluates the Python callable every step
    tr2.current_epoch = 1
    tr2._train_epoch()
    setter2(7.5)
    tr2.current_epoch = 2
    tr2._train_epoch()
    assert relerr(loss_after_change, float(tr2.total_loss)) < 1e-5
    assert np.abs(model.jet_engine.flat.detach().numpy() - model2.jet_engine.flat.detach().numpy()).max() < 1e-6


def test_in_place_parameter_writes_are_noticed(oracle_backend):
    """MarchingTrainer.reset_weights (reference trainer/marching.py:35-40) re-initialises the Linear layers in place
    between stages.  The engine compares the parameters' version counters before every launch and re-packs its
    operand images; derivative blocks computed with the old weights are not reused."""
    torch.manual_seed(0)
    model = mp.FNN(2, 1, [16, 16])
    x = torch.rand(5, 2)
    tr_like = model.jet_engine
    tr_like.managed = True                        # as under a Trainer: operands are re-packed only when invalid
    y0 = model(x).detach().clone()
    gen0 = tr_like.generation
    calls = [c for c in oracle_backend.calls if c[0] == "pack"]
    for mod in model.modules():
        if isinstance(mod, torch.nn.Linear):
            mod.reset_parameters()
    y1 = model(x).detach()
    assert tr_like.generation > gen0
    assert not torch.allclose(y0, y1)
    ref = torch.nn.Sequential()
    from oracle import autograd_ref as AR
    r = AR.MODELS["FNN"](2, 1, hidden_layers=[16, 16]).double()
    r.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    assert relerr(y1.numpy(), r(x.double()).detach().numpy()) < 1e-6


def test_replaced_residual_function_is_retraced(oracle_backend):
    """MarchingTrainer stages replace a condition's function, the new one calling the PREVIOUS stage's model as well
    (examples/convection_1D_marching/problem.py:39-56): the cached program must not survive; a residual that calls two
    models does not trace and the trainer continues on the eager path."""
    import copy
    torch.manual_seed(0)
    model = mp.FNN(2, 1, [16, 16])
    geo = mp.Hypercube([0, 0], [1, 1])

    def ic(m, a):
        (u,) = mp.unpack(m(a))
        x, y = mp.unpack(a)
        return [u - torch.sin(y)]

    cond = mp.Condition(ic, geo)
    pts = torch.rand((16, 2), generator=torch.Generator().manual_seed(3)).requires_grad_()
    cond.generator = types.SimpleNamespace(n_points=16, generate=lambda c, m: pts)
    pinn = mp.PINN(model, [cond])
    pinn.update_data()
    tr = mp.TrainerOneBatch(pinn, torch.optim.Adam(model.parameters(), lr=1e-3), None, num_epochs=3,
                            update_grid_every=10**9)
    tr.current_epoch = 1
    tr._train_epoch()
    assert tr.fused_step() is not None
    previous = copy.deepcopy(model)

    def ic_new(m, a):
        (u,) = mp.unpack(m(a))
        (u_prev,) = mp.unpack(previous(a))
        return [u - u_prev]

    cond.function = ic_new
    tr.current_epoch = 2
    tr._train_epoch()
    assert tr.fused_step() is None and "two different models" in cond.trace_note
    with torch.no_grad():
        want = float(((model(pts) - previous(pts)) ** 2).mean())
    tr.current_epoch = 3
    tr._train_epoch()
    assert abs(float(tr.total_loss) - want) / want < 1e-5


@pytest.mark.parametrize("mode", ["mean", "legacy"])
def test_multi_batch_epochs_accumulate_on_the_fused_step(mode, oracle_backend):
    """`Trainer(num_batches=3)` (examples/navier_stokes_2D_obstacle/run_train.py:86): the gradient accumulates over the
    batches and ONE optimizer step follows (reference trainer.py:119-153).  Fused path vs the eager jet path over 3
    epochs: losses and weights."""
    runs = []
    for fused in (True, False):
        g = load("poisson_2D_1C")
        model, conds, pinn = build_pinn("poisson_2D_1C", g)
        n_pts = [len(c.points) for c in conds]
        assert all(n % 2 == 0 for n in n_pts)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        tr = mp.Trainer(pinn, opt, None, num_epochs=3, num_batches=2, update_grid_every=10**9, calc_loss=mode)
        tr.use_fused_step = fused
        hist = []
        for e in range(1, 4):
            tr.current_epoch = e
            oracle_backend.calls.clear()
            tr._train_epoch()
            hist.append((float(tr.total_loss), tr.epoch_loss_detailed.numpy().copy()))
            n_adam = sum(1 for c in oracle_backend.calls if c[0] in ("adam", "adam_dev"))
            assert n_adam == 1                                   # one optimizer step per epoch, not per batch
        if fused:
            assert tr.fused_step() is not None
        runs.append((hist, model.jet_engine.flat.detach().numpy().copy()))
    (h0, w0), (h1, w1) = runs
    for (t0, d0), (t1, d1) in zip(h0, h1):
        assert relerr(t0, t1) < 1e-5 and relerr(d0, d1) < 1e-5
    assert np.abs(w0 - w1).max() < 5e-6
