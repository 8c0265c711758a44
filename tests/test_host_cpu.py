"""Host-side logic of the product (jet-block decoding, residual plans, flat buffers, PINN, trainers)
exercised on CPU with the oracle standing in for the kernels (tests/oracle_backend.py), and checked
against vectors from the unmodified reference (tests/golden)."""
import copy
import io
import os
import types

import numpy as np
import pytest
import torch

import multipinn_b200 as mp
from multipinn_b200 import engine, jet
from oracle import autograd_ref as AR
from tests import problems
from tests.oracle_backend import OracleBackend
from tests.test_oracle_golden import GOLD_KW, NAMES, load, relerr

MODELS = {"FNN": mp.FNN, "ResNet": mp.ResNet, "FFNN": mp.FFNN}


@pytest.fixture(autouse=True)
def oracle_backend():
    be = OracleBackend()
    old = engine._set_backend_for_tests(be)
    jet.clear_blocks()
    yield be
    engine._set_backend_for_tests(old)


def product_ns():
    return types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack, Symbols=mp.Symbols)


def product_model(name, g):
    _, kind, _ = problems.CONFIGS[name]
    m = MODELS[kind](int(g["input_dim"]), int(g["output_dim"]), **GOLD_KW[name])
    sd = {k: torch.tensor(g["w::" + k]) for k in g["state_keys"]}
    m.load_state_dict(sd, strict=True)  # same state_dict keys as the reference classes
    return m


class _Gen:
    def __init__(self, pts):
        self.pts, self.n_points = pts, len(pts)

    def generate(self, cond, model):
        return self.pts


def build_pinn(name, g, world=1):
    model = product_model(name, g)
    conds_spec, din, dout = problems.CONFIGS[name][0](product_ns())
    conds = []
    for i, (fn, (lo, hi)) in enumerate(conds_spec):
        c = mp.Condition(fn, mp.Hypercube(lo, hi))
        c.generator = _Gen(torch.tensor(g[f"pts{i}"]).requires_grad_())
        conds.append(c)
    pinn = mp.PINN(model, conds)
    pinn.update_data()
    for c in conds:
        c.set_batching(world)
    return model, conds, pinn


# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", NAMES)
def test_program_matches_oracle_lowering(name):
    g = load(name)
    m = product_model(name, g)
    prog, flat = AR.build_program(m)
    assert np.array_equal(m.jet_program(), prog)
    assert flat.shape[0] == sum(p.numel() for p in m.parameters())


@pytest.mark.parametrize("name", NAMES)
def test_residuals_loss_and_gradient_match_reference(name, oracle_backend):
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    pinn.select_batch(0)
    k = 0
    for ci, c in enumerate(conds):
        for r in c.get_residual(model):
            assert relerr(r.detach().numpy(), g[f"res{ci}_{k}_f64"]) < 1e-5, (ci, k)
            k += 1
    model.jet_engine.zero_grad()
    total, losses = pinn.calculate_mean_loss()
    assert relerr(total.detach().numpy(), g["total_f64"]) < 1e-5
    assert relerr(torch.stack(losses).detach().numpy(), g["losses_f64"]) < 1e-5
    legacy = torch.sum(torch.stack(pinn.calculate_loss()))
    assert relerr(legacy.detach().numpy(), g["legacy_f64"]) < 1e-5
    total.backward()
    flat_grad = model.jet_engine.flat_grad.numpy()
    assert relerr(flat_grad, g["grad_f64"]) < 1e-5
    # per-parameter .grad are views of the flat buffer
    got = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                          for p in model.parameters()])
    assert np.array_equal(got, flat_grad)


def test_plan_discovers_orders_once(oracle_backend):
    g = load("poisson_2D_1C")
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    assert conds[0].plan_for(model).orders == [2]          # inner: u_xx, u_yy
    assert all(c.plan_for(model).orders == [0] for c in conds[1:])  # Dirichlet sides: values only
    oracle_backend.calls.clear()
    pinn.select_batch(0)
    pinn.calculate_mean_loss()
    fwd = [c for c in oracle_backend.calls if c[0] == "fwd"]
    assert [c[1] for c in fwd] == [2, 0, 0, 0, 0]           # one launch per model call, no re-runs

    g = load("allen_cahn_2D_1C")
    model, conds, pinn = build_pinn("allen_cahn_2D_1C", g)
    assert [c.plan_for(model).orders for c in conds] == [[2], [1], [1], [0]]
    g = load("burgers_1D")
    model, conds, pinn = build_pinn("burgers_1D", g)
    assert [c.plan_for(model).orders for c in conds] == [[2], [0, 0], [0]]


def test_grad_without_plan_escalates_lazily(oracle_backend):
    g = load("poisson_2D_1C")
    model = product_model("poisson_2D_1C", g)
    fn = problems.poisson_2D_1C(product_ns())[0][0][0]
    pts = torch.tensor(g["pts0"]).requires_grad_()
    (r,) = fn(model, pts)  # bare call, no Condition: orders 0 -> 1 -> 2 on demand
    assert relerr(r.detach().numpy(), g["res0_0_f64"]) < 1e-5
    assert [c[1] for c in oracle_backend.calls if c[0] == "fwd"] == [0, 1, 2]
    oracle_backend.calls.clear()
    with mp.jet_order(2):
        (r2,) = fn(model, pts)
    assert [c[1] for c in oracle_backend.calls if c[0] == "fwd"] == [2]
    assert torch.equal(r, r2)


def test_indexing_style_access(oracle_backend):
    """f[:, k] / grad(...)[:, i] indexing (reference examples/poisson_3D/problem.py:22-29) decodes too."""
    g = load("navier_stokes_3D")
    model = product_model("navier_stokes_3D", g)
    pts = torch.tensor(g["pts0"]).requires_grad_()
    with mp.jet_order(2):
        f = model(pts)
        gu = mp.grad(f[:, 1], pts)
        v_yy = mp.grad(gu[:, 1], pts)[:, 1]
    assert relerr(gu.detach().numpy(), g["du_f64"][:, 1, :]) < 1e-5
    assert relerr(v_yy.detach().numpy(), g["d2u_f64"][:, 1, 1]) < 1e-5


def test_unsupported_derivatives_raise(oracle_backend):
    g = load("poisson_2D_1C")
    model = product_model("poisson_2D_1C", g)
    pts = torch.tensor(g["pts0"]).requires_grad_()
    (u,) = mp.unpack(model(pts))
    with pytest.raises(mp.UnsupportedDerivative):
        mp.grad(u * 2.0, pts)                      # derived expression, not a model output view
    with pytest.raises(mp.UnsupportedDerivative):
        mp.grad(u, pts.clone())                    # not the tensor the model saw
    u_x, u_y = mp.unpack(mp.grad(u, pts))
    u_xx, u_xy = mp.unpack(mp.grad(u_x, pts))
    assert bool(torch.isnan(u_xy).all())           # mixed partial is poisoned, never silently wrong
    with pytest.raises(mp.UnsupportedDerivative):
        mp.grad(u_xx, pts)                         # third order

    def mixed(model, arg):
        (u,) = mp.unpack(model(arg))
        u_x, u_y = mp.unpack(mp.grad(u, arg))
        u_xx, u_xy = mp.unpack(mp.grad(u_x, arg))
        return [u_xy]

    # a mixed derivative that is USED raises the condition to the full-Hessian jets (order 4) at the probe
    c = mp.Condition(mixed, mp.Hypercube([0, 0], [1, 1]))
    c.init_output_len(model)
    assert c.plan_for(model).orders == [4]
    (r,) = c.get_residual_fn(model)(pts)
    assert relerr(r.detach().numpy(), _uxy_fp64(model, g["pts0"])) < 1e-5


def _uxy_fp64(model, pts_np):
    """d2u/dxdy by nested autograd of the reference algorithm (oracle port) in fp64."""
    ref = AR.MODELS["FNN"](2, 1, **GOLD_KW["poisson_2D_1C"]).double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    x = torch.tensor(pts_np, dtype=torch.float64).requires_grad_()
    (u,) = AR.unpack(ref(x))
    u_x, _ = AR.unpack(AR.grad(u, x))
    _, u_xy = AR.unpack(AR.grad(u_x, x))
    return u_xy.detach().numpy()


def test_symbols_generated_source():
    s = mp.Symbols(variables="x, y", functions="u")
    fn = s("u, u_x, u_yy")
    assert "(u,) = unpack(model(arg))" in fn.source
    assert "u_x, u_y = unpack(grad(u, arg))" in fn.source
    assert "u_yx, u_yy = unpack(grad(u_y, arg))" in fn.source
    assert fn.source.strip().endswith("return u, u_x, u_yy")
    with pytest.raises(ValueError):
        mp.Symbols("x, x", "u")
    with pytest.raises(ValueError):
        s("w_x")


def test_condition_batching_slices():
    """select_batch / set_batching (reference tests/condition/test_condition.py:59-81)."""
    c = mp.Condition(lambda m, a: [a[:, 0]], mp.Hypercube([0, 0], [1, 1]))
    c.generator = types.SimpleNamespace(n_points=100)
    c.points = torch.arange(200.0).reshape(100, 2)
    c.set_batching(4)
    assert c.batch_size == 25
    c.select_batch(2)
    assert torch.equal(c.batch_points, c.points[50:75])
    with pytest.raises(AssertionError):
        c.set_batching(3)


@pytest.mark.parametrize("name", ["poisson_2D_1C", "navier_stokes_2D", "navier_stokes_3D", "burgers_1D"])
def test_trainer_steps_match_reference(name, oracle_backend):
    """TrainerOneBatch + torch Adam (fused flat Adam underneath) follows the reference's 3-step
    trajectory: loss per step and parameters after steps 1 and 3."""
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=3, update_grid_every=10**9, calc_loss="mean")
    assert tr._flat_adam is not None
    traj = []
    for e in range(3):
        tr.current_epoch = e + 1
        tr._train_epoch()
        sched.step()
        traj.append(float(tr.total_loss))
        if e == 0:
            flat = model.jet_engine.flat.detach().numpy().copy()
            # first Adam step moves every weight by ~lr*g/(|g|+eps): a weight whose gradient is of the
            # order of eps amplifies fp32 round-off in g, hence 0.5% of lr rather than an ulp bound
            assert np.abs(flat - g["params_after1"]).max() < 5e-6
            assert relerr(tr.epoch_loss_detailed.numpy(), g["losses_step1"]) < 1e-5
    assert relerr(np.array(traj), g["traj_f32"]) < 2e-5
    assert np.abs(model.jet_engine.flat.detach().numpy() - g["params_after3"]).max() < 1e-5
    p0 = [p for p in model.parameters() if p.requires_grad][0]
    st = opt.state[p0]
    assert float(st["step"]) == 3.0 and st["exp_avg"].shape == p0.shape


def test_trainer_accumulates_batches(oracle_backend):
    """num_batches=2 accumulation (reference trainer.py:130-153) == mean of the two half-batch grads."""
    g = load("allen_cahn_2D_1C")
    model, conds, pinn = build_pinn("allen_cahn_2D_1C", g)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)  # non-Adam optimizer -> generic step path
    tr = mp.Trainer(pinn, opt, None, num_epochs=1, num_batches=2, update_grid_every=10**9)
    assert tr._flat_adam is None
    tr.current_epoch = 1
    tr._train_epoch()
    got = model.jet_engine.flat_grad.numpy().copy()
    grads = []
    for b in range(2):
        model.jet_engine.zero_grad()
        pinn.select_batch(b)
        total, _ = pinn.calculate_mean_loss()
        total.backward()
        grads.append(model.jet_engine.flat_grad.numpy().copy())
    assert relerr(got, grads[0] + grads[1]) < 1e-6


def test_state_dict_pickle_and_move(oracle_backend):
    g = load("navier_stokes_2D")
    model = product_model("navier_stokes_2D", g)
    pts = torch.tensor(g["pts0"])
    with torch.no_grad():
        y0 = model(pts).clone()
    sd = copy.deepcopy(model.state_dict())
    assert list(sd.keys()) == list(g["state_keys"])
    buf = io.BytesIO()
    torch.save(model, buf)  # what the reference's SaveModel callback does (callbacks/save.py:79)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    with torch.no_grad():
        assert torch.equal(m2(pts), y0)
    # in-place parameter edits through the nn.Parameter views are seen by the engine
    with torch.no_grad():
        next(iter(model.parameters())).mul_(0.5)
        assert not torch.equal(model(pts), y0)
    model.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(model(pts), y0)


def test_inference_under_no_grad_skips_stash(oracle_backend):
    g = load("poisson_2D_1C")
    model = product_model("poisson_2D_1C", g)
    pts = torch.tensor(g["pts0"])
    with torch.no_grad():
        y = model(pts)
    assert not y.requires_grad
    assert relerr(y.numpy(), g["u_f64"]) < 1e-5
    with pytest.raises(RuntimeError):
        model(torch.zeros(4, 3))      # wrong input dim
    with pytest.raises(TypeError):
        model(pts.double())           # fp32 only


def test_gradient_losses_matches_reference(oracle_backend):
    """GradientLosses(alpha=0.9) -- poisson_2D_1C's shipped regularizer -- follows the reference's lambdas,
    weighted totals, component losses and weights over 3 steps (reference regularization/gradient.py:25-59),
    with ONE reverse launch per condition (the reference runs every condition's backward twice)."""
    g = load("poisson_2D_1C")
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", "poisson_gradient_reg.npz"))
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = mp.GradientLosses(alpha=0.9)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg)
    for e in range(3):
        tr.current_epoch = e + 1
        oracle_backend.calls.clear()
        tr._train_epoch()
        assert sum(1 for c in oracle_backend.calls if c[0] == "bwd") == len(conds)
        assert relerr(reg.lambda_regularization.numpy(), gr["lambdas"][e]) < 2e-4
        assert relerr(float(tr.total_loss), gr["totals"][e]) < 1e-5
        assert relerr(tr.epoch_loss_detailed.numpy(), gr["losses"][e]) < 1e-5
    assert np.abs(model.jet_engine.flat.detach().numpy() - gr["params_after3"]).max() < 2e-5


def test_normal_losses_matches_reference(oracle_backend):
    """NormalLosses(alpha=0.9) follows the reference's lambdas, totals, losses and weights over 3 steps
    (reference regularization/normal.py:6-45), with ONE reverse launch per condition."""
    g = load("poisson_2D_1C")
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", "poisson_normal_reg.npz"))
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = mp.NormalLosses(alpha=0.9)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg)
    for e in range(3):
        tr.current_epoch = e + 1
        oracle_backend.calls.clear()
        tr._train_epoch()
        assert sum(1 for c in oracle_backend.calls if c[0] == "bwd") == len(conds)
        assert relerr(reg.lambda_regularization.numpy(), gr["lambdas"][e]) < 2e-4
        assert relerr(float(tr.total_loss), gr["totals"][e]) < 1e-5
        assert relerr(tr.epoch_loss_detailed.numpy(), gr["losses"][e]) < 1e-5
    assert np.abs(model.jet_engine.flat.detach().numpy() - gr["params_after3"]).max() < 2e-5


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "eager"])
def test_pointwise_losses_matches_reference(fused, oracle_backend):
    """PointwiseLosses(ignored_indexes=[1], lr=0.01): per-point weights after each maximising Adam step, weighted
    totals, reported losses and network weights over 3 steps (reference regularization/pointwise.py:9-152).  On the
    fused step the weights are an input of the residual launch (seeds 2 lambda r, sum lambda r^2 accumulated next to
    sum r^2) and their Adam ascent runs on the device; on the eager path the residuals are evaluated once per step
    (one forward per condition)."""
    g = load("poisson_2D_1C")
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", "poisson_pointwise_reg.npz"))
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = mp.PointwiseLosses(ignored_indexes=[1], lr=0.01)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg)
    tr.use_fused_step = fused
    assert (tr.fused_step() is not None) == fused
    for e in range(3):
        tr.current_epoch = e + 1
        oracle_backend.calls.clear()
        tr._train_epoch()
        n_fwd = sum(1 for c in oracle_backend.calls if c[0] == "fwd")
        assert n_fwd == (2 if fused else len(conds))      # fused: domain + the four boundary conditions batched
        ms = reg.models[0]
        assert ms[1] is None and reg.ignored_mask[1]
        lam = np.concatenate([ms[0].weight.detach().numpy().reshape(-1)[:16],
                              ms[-1].weight.detach().numpy().reshape(-1)[:16]])
        assert relerr(lam, gr["lambdas"][e]) < 1e-5
        assert relerr(float(tr.total_loss), gr["totals"][e]) < 1e-5
        assert relerr(tr.epoch_loss_detailed.numpy(), gr["losses"][e]) < 1e-5
    assert np.abs(model.jet_engine.flat.detach().numpy() - gr["params_after3"]).max() < 2e-5


def test_constant_losses(oracle_backend):
    g = load("poisson_2D_1C")
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    pinn.select_batch(0)
    tr = types.SimpleNamespace(pinn=pinn)
    lam = [1.0, 2.0, 3.0, 4.0, 5.0]
    total, losses = mp.ConstantLosses(lam)(tr)
    assert relerr(losses.detach().numpy(), g["losses_f64"] * np.array(lam)) < 1e-5
    assert relerr(total.item(), float((g["losses_f64"] * np.array(lam)).sum())) < 1e-5
    with pytest.raises(Exception):
        mp.ConstantLosses([1.0])(tr)


@pytest.mark.parametrize("name", ["poisson_2D_1C", "burgers_1D", "navier_stokes_2D", "navier_stokes_3D"])
def test_points_grad_matches_reference_backward(name, oracle_backend):
    """Opt-in `points_grad`: after backward() every condition's `points.grad` equals what the reference's nested
    autograd leaves there (reference trainer/trainer.py:141 differentiates w.r.t. the points as a side effect;
    generation/grad_based_generator.py:66 reads it)."""
    g = load(name)
    model, conds, pinn = build_pinn(name, g)
    model.jet_engine.points_grad = True
    pinn.select_batch(0)
    total, _ = pinn.calculate_mean_loss()
    total.backward()
    # the reference algorithm, fp64, same weights and points
    fn, kind, _ = problems.CONFIGS[name]
    spec, din, dout = fn(types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None))
    ref = AR.MODELS[kind](int(g["input_dim"]), int(g["output_dim"]), **GOLD_KW[name]).double()
    ref.load_state_dict({k: torch.tensor(g["w::" + k]).double() for k in g["state_keys"]})
    pts = [torch.tensor(g[f"pts{i}"]).double().requires_grad_() for i in range(len(spec))]
    res = [r for (f, _), x in zip(spec, pts) for r in f(ref, x)]
    # "mean" loss: all squared residual entries over their total count (reference PINN/PINN.py:112-125)
    (sum((r ** 2).sum() for r in res) / sum(r.numel() for r in res)).backward()
    for c, x in zip(conds, pts):
        assert c.points.grad is not None
        assert relerr(c.points.grad.numpy(), x.grad.numpy()) < 1e-5


@pytest.mark.parametrize("kind", ["adaptive", "adaptive_points"])
def test_adaptive_losses_match_reference(kind, oracle_backend):
    """AdaptiveConditionsLosses / AdaptiveConditionsAndPointsLosses (lr=0.05): softmax weights after each ascent
    step, weighted totals, reported losses and network weights over 3 steps vs the unmodified reference's run
    (reference regularization/adaptive.py:75-242, fixtures tests/golden/poisson_adaptive*_reg.npz)."""
    g = load("poisson_2D_1C")
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", f"poisson_{kind}_reg.npz"))
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = mp.AdaptiveConditionsLosses(lr=0.05) if kind == "adaptive" else mp.AdaptiveConditionsAndPointsLosses(lr=0.05)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg)
    # the condition-level policy is a fused loss policy (softmax weights, their Adam ascent and the weighted total are
    # a few device ops inside the captured step); the point-level variant keeps the eager jet path
    assert (tr.fused_step() is not None) == (kind == "adaptive")
    for e in range(3):
        tr.current_epoch = e if kind == "adaptive_points" else e + 1
        tr._train_epoch()
        lam = [reg.adaptive_cond.weight.detach().numpy()]
        if kind == "adaptive_points":
            lam += [reg.adaptive_points[0].weight.detach().numpy()[:16], reg.adaptive_points[-1].weight.detach().numpy()[:16]]
        assert relerr(np.concatenate(lam), gr["lambdas"][e]) < 1e-5
        assert relerr(float(tr.total_loss), gr["totals"][e]) < 1e-5
        assert relerr(tr.epoch_loss_detailed.numpy(), gr["losses"][e]) < 1e-5
    assert np.abs(model.jet_engine.flat.detach().numpy() - gr["params_after3"]).max() < 2e-5


def test_frozen_layers_use_the_generic_optimizer_path(oracle_backend):
    """A model with a frozen Linear layer must not be stepped by the flat fused Adam (it updates the whole
    buffer): the trainer falls back to torch.optim.Adam, which skips parameters without a gradient."""
    g = load("poisson_2D_1C")
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    frozen = list(model.parameters())[2]
    frozen.requires_grad_(False)
    before = frozen.detach().clone()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=2, update_grid_every=10**9)
    assert tr._flat_adam is None
    other = list(model.parameters())[0].detach().clone()
    for e in range(2):
        tr.current_epoch = e + 1
        tr._train_epoch()
    assert torch.equal(frozen.detach(), before)
    assert not torch.equal(list(model.parameters())[0].detach(), other)
    # the frozen Fourier matrix of an FFNN stays on the fused Adam (its gradient slot is never written)
    g3 = load("navier_stokes_3D")
    m3, c3, p3 = build_pinn("navier_stokes_3D", g3)
    tr3 = mp.TrainerOneBatch(p3, torch.optim.Adam(m3.parameters(), lr=1e-3), None, num_epochs=1, update_grid_every=10**9)
    assert tr3._flat_adam is not None


def test_symbols_diff_mode_generates_finite_difference_code(oracle_backend):
    """Symbols(mode="diff") (reference condition/symbols.py:189-234): generated source, and its values against the
    analytic jets (central second differences are accurate to ~1e-3 at eps = 2^-7)."""
    s = mp.Symbols(variables="x, t", functions="u", mode="diff")
    fn = s("u, u_x, u_xx, u_xt")
    src = fn.source
    assert "fns = model(arg)" in src and "(u,) = unpack(fns)" in src
    assert "(u_x,) = unpack(num_diff_random(model, arg, fns, torch.Tensor([[1, 0]])))" in src
    assert "(u_xx,) = unpack(num_diff_second_same(model, arg, fns, torch.Tensor([[1, 0]])))" in src
    assert "(u_xt,) = unpack(num_diff_second_cross(model, arg, torch.Tensor([[1, 0]]), torch.Tensor([[0, 1]])))" in src
    with pytest.raises(ValueError):
        s("u_xxx")
    with pytest.raises(ValueError):
        mp.Symbols("x", "u", mode="nope")("u")
    g = load("allen_cahn_2D_1C")
    model = product_model("allen_cahn_2D_1C", g)
    pts = torch.tensor(g["pts0"])
    torch.manual_seed(0)
    with torch.no_grad():
        u, u_x, u_xx, u_xt = fn(model, pts)
    assert relerr(u.numpy(), g["u_f64"][:, 0]) < 1e-5
    assert relerr(u_xx.numpy(), g["d2u_f64"][:, 0, 0]) < 2e-2
    assert relerr(u_x.numpy(), g["du_f64"][:, 0, 0]) < 5e-2


def test_inference_wrapper(oracle_backend):
    g = load("poisson_2D_1C")
    model = product_model("poisson_2D_1C", g)
    pts = torch.tensor(g["pts0"])
    inf = mp.neural_network.Inference(model, batchsize=32)
    with torch.no_grad():
        y = inf(pts)
    assert y.shape == (len(pts), 1)
    assert relerr(y.numpy(), g["u_f64"]) < 1e-5
    inf.launch_points = 40         # several launches stack to the same answer
    with torch.no_grad():
        assert torch.equal(inf(pts), y)


def test_device_generators(oracle_backend):
    """DeviceGenerator: uniform points inside the box, reproducible with a seed.  ResidualAdaptiveGenerator: the
    error field comes from the value-only residual program and the added points favour large residuals."""
    g = load("poisson_2D_1C")
    model, conds, pinn = build_pinn("poisson_2D_1C", g)
    c = conds[0]
    gen = mp.DeviceGenerator(500, seed=3)
    a = gen.generate(c, model)
    assert a.shape == (500, 2) and a.requires_grad and a.dtype == torch.float32
    assert float(a.min()) >= 0.0 and float(a.max()) <= 0.5
    b = mp.DeviceGenerator(500, seed=3).generate(c, model)
    assert torch.equal(a, b) and not torch.equal(a, gen.generate(c, model))
    rar = mp.ResidualAdaptiveGenerator(64, density_rec_points_num=400, add_points=40, n_points_up_bnd=150, seed=1)
    rar.use_for(c)
    c.points = None
    c.update_points(model)
    assert c.points.shape == (64, 2)
    oracle_backend.calls.clear()
    c.update_points(model)
    assert c.points.shape == (104, 2)
    kinds = [k[0] for k in oracle_backend.calls]
    assert kinds.count("fwd") == 1 and kinds.count("resid") == 1 and "bwd" not in kinds   # value-only error field
    pool = rar._uniform(400, c.geometry, torch.device("cpu"))
    err = rar.error_field(c, model, pool)
    eager = torch.stack([r.detach() for r in c.get_residual_fn(model)(pool.clone().requires_grad_())], 1).abs().sum(1)
    assert relerr(err.numpy(), eager.numpy()) < 1e-5
    c.update_points(model)
    c.update_points(model)
    assert c.points.shape[0] == 184
    c.update_points(model)                 # upper bound reached: unchanged
    assert c.points.shape[0] == 184


def test_trainable_fourier_matrix_gets_a_gradient(oracle_backend):
    """FourierFeatureNetwork(is_trainable=True) (reference fourier_features.py:30-32): the program marks the encoding
    trainable, the reverse pass returns dL/dB, and loss / gradient equal nested autograd of the reference algorithm."""
    torch.manual_seed(4)
    model = mp.FFNN(3, 2, hidden_layers=[16, 16], encoding_dim=16, sigma=1.0, is_trainable=True)
    prog = model.jet_program()
    assert prog[0, 0] == 1 and prog[0, 6] == 4
    assert np.array_equal(prog, AR.build_program(model)[0])

    def make(ns):
        def fn(m, a):
            u, v = ns.unpack(m(a))
            u_x, u_y, u_z = ns.unpack(ns.grad(u, a))
            u_xx, _, _ = ns.unpack(ns.grad(u_x, a))
            _, v_y, _ = ns.unpack(ns.grad(v, a))
            return [u_xx + v_y * u - u_z, v - u_y]
        return fn

    cond = mp.Condition(make(types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack)), mp.Hypercube([0, 0, 0], [1, 1, 1]))
    pts = torch.rand((50, 3), generator=torch.Generator().manual_seed(0)).requires_grad_()
    cond.generator = types.SimpleNamespace(n_points=50, generate=lambda c, m: pts)
    pinn = mp.PINN(model, [cond])
    pinn.update_data()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=1, update_grid_every=10**9)
    assert tr._flat_adam is not None
    ref = AR.MODELS["FFNN"](3, 2, hidden_layers=[16, 16], encoding_dim=16, sigma=1.0, is_trainable=True).double()
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    total, losses = AR.mean_loss(ref, [make(types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack))],
                                 [pts.detach().double().requires_grad_()])
    total.backward()
    rgrad = np.concatenate([p.grad.numpy().reshape(-1) for p in ref.parameters()])
    B0 = model.fourier_encoding.B.detach().clone()
    pinn.select_batch(0)
    tr.current_epoch = 1
    tr._train_epoch()
    assert relerr(tr.epoch_loss_detailed.numpy(), losses.detach().numpy()) < 1e-5
    assert relerr(model.jet_engine.flat_grad.numpy(), rgrad) < 1e-5
    assert float(model.jet_engine.flat_grad[:B0.numel()].abs().max()) > 0
    assert not torch.equal(model.fourier_encoding.B.detach(), B0)          # B is trained


def test_pointwise_fused_follows_the_grid_update_quirk(oracle_backend):
    """After a grid update the reference rebuilds the weight modules (lambda = 1) but keeps stepping the FIRST
    optimizer (reference pointwise.py:60-66, 137-152), so only the first generation of weights is ever trained.  The
    fused policy must do the same as the eager mirror of that code over epochs that cross two grid updates."""
    runs = []
    for fused in (True, False):
        g = load("poisson_2D_1C")
        model, conds, pinn = build_pinn("poisson_2D_1C", g)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        reg = mp.PointwiseLosses(ignored_indexes=[1], lr=0.05)
        tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=7, update_grid_every=3, calc_loss=reg)
        tr.use_fused_step = fused
        totals, lam = [], []
        for e in range(1, 8):          # grid updates (and weight resets) at epochs 3 and 6
            tr.current_epoch = e
            tr._train_epoch()
            totals.append(float(tr.total_loss))
            lam.append(reg.models[0][0].weight.detach().numpy().reshape(-1)[:8].copy())
        assert len(reg.optimizers) == 3
        runs.append((np.array(totals), np.array(lam), model.jet_engine.flat.detach().numpy().copy()))
    (t0, l0, w0), (t1, l1, w1) = runs
    assert relerr(t0, t1) < 1e-5
    assert np.abs(l0 - l1).max() < 1e-5
    assert not np.allclose(l0[1], 1.0) and np.allclose(l0[2:], 1.0)      # trained on epoch 2, reset (for good) at 3
    assert np.abs(w0 - w1).max() < 1e-5
