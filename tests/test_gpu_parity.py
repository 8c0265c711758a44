"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything here goes through the C ABI
of libmpinn_b200.so.  Expected values come from the committed golden vectors (unmodified reference,
fp64 and fp32) and from the oracle evaluated in fp64 on the same seeded inputs.

Tolerance (BASELINE.json north_star): relative error <= 1e-5, measured max-normalised
(max|a-b| / max|b|) on u, grad u, the Hessian diagonal, residuals, losses and the parameter gradient.
"""
import types

import numpy as np
import pytest
import torch

import multipinn_b200 as mp
from multipinn_b200 import _lib, engine, jet
from oracle import autograd_ref as AR
from oracle import jet_ref as JR
from tests import problems
from tests.test_oracle_golden import GOLD_KW, NAMES, load, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-5
MODELS = {"FNN": mp.FNN, "ResNet": mp.ResNet, "FFNN": mp.FFNN}
ENGINES = [_lib.ENGINE_FFMA, _lib.ENGINE_AUTO]
ENGINE_IDS = ["ffma", "auto"]


@pytest.fixture(autouse=True)
def cuda_backend():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    old = engine._set_backend_for_tests(None)   # make sure the real CUDA backend is the one in use
    jet.clear_blocks()
    be = engine.backend()
    assert be.name == "cuda"
    yield be
    engine._set_backend_for_tests(old)


def product_ns():
    return types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack, Symbols=mp.Symbols)


def golden_model(name, g, kernel_engine=_lib.ENGINE_AUTO):
    _, kind, _ = problems.CONFIGS[name]
    m = MODELS[kind](int(g["input_dim"]), int(g["output_dim"]), **GOLD_KW[name])
    m.load_state_dict({k: torch.tensor(g["w::" + k]) for k in g["state_keys"]}, strict=True)
    m = m.cuda()
    m.jet_engine.kernel_engine = kernel_engine
    return m


def full_model(name, seed=0, kernel_engine=_lib.ENGINE_AUTO):
    """Random-init model at the BASELINE config's full depth."""
    fn, kind, kw = problems.CONFIGS[name]
    _, din, dout = fn(product_ns())
    torch.manual_seed(seed)
    m = MODELS[kind](din, dout, **kw).cuda()
    m.jet_engine.kernel_engine = kernel_engine
    return m, din, dout


def oracle_jets(model, x_np, order, dout_np=None):
    prog, flat = AR.build_program(model)
    out, cache = JR.forward(prog, flat, x_np.astype(np.float64), order)
    if dout_np is None:
        return out, None
    dp, _ = JR.backward(prog, flat, x_np.astype(np.float64), order, cache, dout_np.astype(np.float64))
    if prog[0, 0] == JR.OP_FOURIER:
        dp[prog[0, 3]: prog[0, 3] + prog[0, 1] * prog[0, 2] // 2] = 0
    return out, dp


# ------------------------------------------------------------------------------ kernels vs golden
@pytest.mark.parametrize("eng", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("name", NAMES)
def test_forward_jets_match_reference_golden(name, eng):
    g = load(name)
    m = golden_model(name, g, eng)
    x = torch.tensor(g["pts0"]).cuda()
    d = x.shape[1]
    out, _ = m.jet_engine.forward(x, 2, need_stash=False)
    out = out.cpu().numpy()                                   # [J,N,o]
    assert relerr(out[0], g["u_f64"]) < TOL
    assert relerr(np.transpose(out[1:1 + d], (1, 2, 0)), g["du_f64"]) < TOL
    assert relerr(np.transpose(out[1 + d:], (1, 2, 0)), g["d2u_f64"]) < TOL
    for order in (0, 1):
        o2, _ = m.jet_engine.forward(x, order, need_stash=False)
        e = m.jet_engine
        if eng == _lib.ENGINE_FFMA or e.tc_supported(order) == e.tc_supported(2):
            assert np.array_equal(o2.cpu().numpy(), out[: 1 + order * d])  # lower orders are prefixes
        else:   # a 128-wide net: <= 4 channels on the layer-wise tcgen05 engine, the 5-channel jet on FP32-FMA
            assert relerr(o2.cpu().numpy(), out[: 1 + order * d]) < TOL


@pytest.mark.parametrize("eng", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("name", NAMES)
def test_backward_matches_oracle(name, order, eng):
    g = load(name)
    m = golden_model(name, g, eng)
    rng = np.random.default_rng(7)
    x_np = g["pts0"]
    n, d = x_np.shape
    J = 1 + order * d
    dout_np = rng.standard_normal((n, J, int(g["output_dim"]))).astype(np.float32)
    x = torch.tensor(x_np).cuda()
    out, stash = m.jet_engine.forward(x, order, need_stash=True)
    dout = torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda()
    dflat = m.jet_engine.backward(x, order, dout, stash).cpu().numpy()
    ref_out, ref_dp = oracle_jets(m, x_np, order, dout_np)
    assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref_out) < TOL
    assert relerr(dflat, ref_dp) < TOL


@pytest.mark.parametrize("eng", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 1000, 4097])
def test_ragged_point_counts(n, eng):
    """Tile edges: every n gives the same per-point answer as the oracle; nothing leaks across the pad."""
    m, din, dout = full_model("poisson_2D_1C", 1, eng)
    rng = np.random.default_rng(n)
    x_np = rng.random((n, 2), dtype=np.float32) * 0.5
    dout_np = rng.standard_normal((n, 5, 1)).astype(np.float32)
    x = torch.tensor(x_np).cuda()
    out, stash = m.jet_engine.forward(x, 2, need_stash=True)
    dflat = m.jet_engine.backward(x, 2, torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda(),
                                  stash).cpu().numpy()
    ref_out, ref_dp = oracle_jets(m, x_np, 2, dout_np)
    assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref_out) < TOL
    assert relerr(dflat, ref_dp) < TOL


@pytest.mark.parametrize("name,dim", [("poisson_2D_1C", 2), ("navier_stokes_2D", 2), ("navier_stokes_3D", 3)])
def test_persistent_tiles_match_oracle(name, dim):
    """More tiles than CTAs: every CTA walks several tiles, so the operand prefetch that runs ahead across tile
    boundaries (next tile's first layer fetched during this tile's last one) is on the path."""
    m, _, _ = full_model(name, 5, _lib.ENGINE_TCGEN05)
    n = 148 * 128 * 2 + 77
    rng = np.random.default_rng(7)
    x_np = rng.random((n, dim), dtype=np.float32) * 0.5
    x = torch.tensor(x_np).cuda()
    for order in (2, 0):
        out, stash = m.jet_engine.forward(x, order, need_stash=True)
        dout_np = rng.standard_normal((n, 1 + order * dim, out.shape[2])).astype(np.float32)
        dflat = m.jet_engine.backward(x, order, torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda(),
                                      stash).cpu().numpy()
        ref_out, ref_dp = oracle_jets(m, x_np, order, dout_np)
        assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref_out) < TOL, (name, order)
        assert relerr(dflat, ref_dp) < TOL, (name, order)


def test_misaligned_packed_buffer_is_rejected(cuda_backend):
    """The packed operand images are fetched by the TMA engine: the C ABI refuses a buffer that is not 16-byte aligned."""
    m, _, _ = full_model("poisson_2D_1C")
    eng = m.jet_engine
    x = torch.rand((64, 2), device="cuda")
    eng.forward(x, 0, need_stash=False)   # builds the flat / packed buffers
    ops = m.jet_program()
    raw = torch.zeros(cuda_backend.packed_bytes(ops) // 4 + 8, device="cuda")
    out = torch.empty((1, 64, 1), device="cuda")
    with pytest.raises(RuntimeError, match="16-byte"):
        cuda_backend.jet_forward(ops, eng.flat, raw[1:], x, 0, out, None)


@pytest.mark.parametrize("eng", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("name,dim", [("poisson_2D_1C", 2), ("navier_stokes_2D", 2)])
def test_points_gradient_matches_oracle(name, dim, eng):
    """dx of mpinn_jet_backward (dL/d points) against the oracle's reverse pass, both engines, ragged size."""
    m, _, _ = full_model(name, 5, eng)
    n = 777
    rng = np.random.default_rng(11)
    x_np = rng.random((n, dim), dtype=np.float32) * 0.5
    x = torch.tensor(x_np).cuda()
    out, stash = m.jet_engine.forward(x, 2, need_stash=True)
    dout_np = rng.standard_normal((n, 1 + 2 * dim, out.shape[2])).astype(np.float32)
    dout = torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda()
    dflat, dx = m.jet_engine.backward(x, 2, dout, stash, want_dx=True)
    prog, flat = AR.build_program(m)
    _, cache = JR.forward(prog, flat.astype(np.float64), x_np.astype(np.float64), 2)
    ref_dp, ref_dx = JR.backward(prog, flat.astype(np.float64), x_np.astype(np.float64), 2, cache, dout_np.astype(np.float64))
    assert relerr(dx.cpu().numpy(), ref_dx) < TOL
    assert relerr(dflat.cpu().numpy(), ref_dp) < TOL


@pytest.mark.parametrize("order,lap", [(1, None), (2, None), (3, (1.0, 0.5, 2.0))])
def test_points_gradient_through_a_fourier_first_op(order, lap):
    """dL/d(points) through a Fourier-feature encoding (x enters z = xB only): the FP32-FMA engine computes it -- the
    host selects that engine for both passes when `points_grad` is set -- against the oracle's reverse pass; the tcgen05
    reverse kernel refuses the request by name instead of returning nothing."""
    m, _, _ = full_model("navier_stokes_3D", 5)
    e = m.jet_engine
    n = 333
    rng = np.random.default_rng(5)
    x_np = rng.random((n, 3), dtype=np.float32) * 0.5
    x = torch.tensor(x_np).cuda()
    out, stash = e.forward(x, order, need_stash=True, lap=lap)          # tcgen05 stash
    with pytest.raises(RuntimeError, match="Fourier"):
        e.backward(x, order, torch.ones_like(out), stash, want_dx=True, lap=lap)
    e.points_grad = True                                                # -> FP32-FMA engine for both passes
    assert e._engine_id() == _lib.ENGINE_FFMA
    out, stash = e.forward(x, order, need_stash=True, lap=lap)
    dout_np = rng.standard_normal((n, out.shape[0], out.shape[2])).astype(np.float32)
    dout = torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda()
    dflat, dx = e.backward(x, order, dout, stash, want_dx=True, lap=lap)
    prog, flat = AR.build_program(m)
    _, cache = JR.forward(prog, flat.astype(np.float64), x_np.astype(np.float64), order, lap)
    ref_dp, ref_dx = JR.backward(prog, flat.astype(np.float64), x_np.astype(np.float64), order, cache,
                                 dout_np.astype(np.float64), lap)
    assert relerr(dx.cpu().numpy(), ref_dx) < TOL
    kind, din, dout_, w_off = (int(v) for v in np.asarray(prog).reshape(-1, 8)[0][:4])
    assert kind == _lib.OP_FOURIER
    ref_dp[w_off:w_off + din * (dout_ // 2)] = 0.0      # the encoding matrix is frozen here: no gradient slot is written
    assert relerr(dflat.cpu().numpy(), ref_dp) < TOL


def test_empty_input_is_a_noop():
    m, _, _ = full_model("poisson_2D_1C")
    x = torch.zeros((0, 2), device="cuda")
    out, _ = m.jet_engine.forward(x, 2, need_stash=False)
    assert out.shape == (5, 0, 1)


# ------------------------------------------------------------------------------ full path vs golden
@pytest.mark.parametrize("name", NAMES)
def test_product_path_matches_reference_golden(name):
    """Condition -> grad/unpack -> PINN mean loss -> backward, on CUDA, vs the reference's numbers."""
    g = load(name)
    model = golden_model(name, g)
    spec, din, dout = problems.CONFIGS[name][0](product_ns())
    conds = []
    for i, (fn, (lo, hi)) in enumerate(spec):
        c = mp.Condition(fn, mp.Hypercube(lo, hi))
        pts = torch.tensor(g[f"pts{i}"]).cuda().requires_grad_()
        c.generator = types.SimpleNamespace(n_points=len(pts), generate=lambda cond, model, p=pts: p)
        conds.append(c)
    pinn = mp.PINN(model, conds)
    pinn.update_data()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=3, update_grid_every=10**9)
    pinn.select_batch(0)
    k = 0
    for ci, c in enumerate(conds):
        for r in c.get_residual(model):
            assert relerr(r.detach().cpu().numpy(), g[f"res{ci}_{k}_f64"]) < TOL, (ci, k)
            k += 1
    model.jet_engine.zero_grad()
    total, losses = pinn.calculate_mean_loss()
    assert relerr(total.item(), g["total_f64"]) < TOL
    assert relerr(torch.stack(losses).detach().cpu().numpy(), g["losses_f64"]) < TOL
    total.backward()
    assert relerr(model.jet_engine.flat_grad.cpu().numpy(), g["grad_f64"]) < TOL
    # three reference train steps: loss trajectory and weights
    traj = []
    for e in range(3):
        tr.current_epoch = e + 1
        tr._train_epoch()
        sched.step()
        traj.append(float(tr.total_loss))
        if e == 0:
            assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - g["params_after1"]).max() < 5e-6
    assert relerr(np.array(traj), g["traj_f32"]) < 2e-5
    assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - g["params_after3"]).max() < 1e-5


def test_gradient_losses_regularizer_matches_reference():
    """poisson_2D_1C's shipped `regularization.type: Gradient` (reference regularization/gradient.py:25-59):
    lambdas, weighted totals and weights after 3 steps vs the reference's own run (tests/golden)."""
    import os
    g = load("poisson_2D_1C")
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", "poisson_gradient_reg.npz"))
    model = golden_model("poisson_2D_1C", g)
    spec, din, dout = problems.CONFIGS["poisson_2D_1C"][0](product_ns())
    conds = []
    for i, (fn, (lo, hi)) in enumerate(spec):
        c = mp.Condition(fn, mp.Hypercube(lo, hi))
        pts = torch.tensor(g[f"pts{i}"]).cuda().requires_grad_()
        c.generator = types.SimpleNamespace(n_points=len(pts), generate=lambda cond, model, p=pts: p)
        conds.append(c)
    pinn = mp.PINN(model, conds)
    pinn.update_data()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = mp.GradientLosses(alpha=0.9)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg)
    for e in range(3):
        tr.current_epoch = e + 1
        tr._train_epoch()
        assert relerr(reg.lambda_regularization.cpu().numpy(), gr["lambdas"][e]) < 5e-4
        assert relerr(float(tr.total_loss), gr["totals"][e]) < 2e-5
    assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - gr["params_after3"]).max() < 2e-5


# ------------------------------------------------------------------------------ vs autograd on the GPU
@pytest.mark.parametrize("name", ["poisson_2D_1C", "burgers_1D", "allen_cahn_2D_1C", "navier_stokes_2D",
                                  "navier_stokes_3D", "regression_1D"])
def test_full_depth_step_vs_fp64_autograd(name):
    """BASELINE depth (9x64 / 4x128 / ResNet-28 / FFNN), 2048+3x512 points: loss, component losses and
    flat parameter gradient vs the oracle's nested-autograd port run in fp64 on the same device."""
    model, din, dout = full_model(name, 3)
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
    model.jet_engine.zero_grad()
    total, losses = pinn.calculate_mean_loss()
    total.backward()
    rtotal, rlosses = AR.mean_loss(ref, [fn for fn, _ in spec_r],
                                   [torch.tensor(p, dtype=torch.float64).cuda().requires_grad_() for p in pts])
    rtotal.backward()
    rgrad = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy().reshape(-1)
                            for p in ref.parameters()])
    assert relerr(total.item(), rtotal.item()) < TOL
    assert relerr(torch.stack(losses).detach().cpu().numpy(), rlosses.detach().cpu().numpy()) < TOL
    # The parameter gradient is a sum over thousands of points; its fp32 round-off floor is set by the
    # summation, not by the jets.  The reference algorithm's own fp32 floor (same nested autograd, fp32, same
    # device) is measured and printed beside ours; the gate is the plain 1e-5 of BASELINE.md.
    ref32 = AR.MODELS[kind](din, dout, **problems.CONFIGS[name][2]).cuda()
    ref32.load_state_dict(model.state_dict())
    t32, _ = AR.mean_loss(ref32, [fn for fn, _ in spec_r],
                          [torch.tensor(p).cuda().requires_grad_() for p in pts])
    t32.backward()
    g32 = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy().reshape(-1)
                          for p in ref32.parameters()])
    floor = relerr(g32, rgrad)
    ours = relerr(model.jet_engine.flat_grad.cpu().numpy(), rgrad)
    print(f"{name}: param-grad rel err ours {ours:.2e}, reference fp32 autograd {floor:.2e}")
    assert ours < TOL      # BASELINE.md's W.grad gate, not relaxed by the reference's own fp32 floor


# ------------------------------------------------------------------------------ size-independent properties
def test_full_size_properties_poisson_1M():
    """1M collocation points (BASELINE configs[1]): properties that need no reference run.
    (a) evaluating the two halves separately gives the same jets as one launch (points are independent);
    (b) the reverse pass is linear in dout and additive over point subsets;
    (c) order-0 output equals channel 0 of the order-2 output."""
    m, _, _ = full_model("poisson_2D_1C", 11)
    n = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((n, 2), device="cuda", generator=g) * 0.5
    e = m.jet_engine
    out, stash = e.forward(x, 2, need_stash=True)
    assert bool(torch.isfinite(out).all())
    h = n // 2
    oa, sa = e.forward(x[:h].contiguous(), 2, need_stash=True)
    ob, sb = e.forward(x[h:].contiguous(), 2, need_stash=True)
    assert torch.equal(torch.cat([oa, ob], dim=1), out)
    o0, _ = e.forward(x, 0, need_stash=False)
    assert torch.equal(o0[0], out[0])
    dout = torch.randn(out.shape, device="cuda", generator=g)
    gfull = e.backward(x, 2, dout, stash)
    ga = e.backward(x[:h].contiguous(), 2, dout[:, :h].contiguous(), sa)
    gb = e.backward(x[h:].contiguous(), 2, dout[:, h:].contiguous(), sb)
    scale = gfull.abs().max()
    assert float((ga + gb - gfull).abs().max() / scale) < 2e-5
    g2 = e.backward(x, 2, (2.5 * dout).contiguous(), stash)
    assert float((g2 - 2.5 * gfull).abs().max() / scale) < 2e-5
    # spot-check 512 of the million points against the fp64 oracle
    idx = torch.randint(0, n, (512,), device="cuda", generator=g)
    ref_out, _ = oracle_jets(m, x[idx].cpu().numpy(), 2)
    assert relerr(out[:, idx].cpu().numpy().transpose(1, 0, 2), ref_out) < TOL


# ------------------------------------------------------------------------------ tensor-core conventions
@pytest.mark.parametrize("mode,label", [(0, "3xTF32 K-major"), (2, "BF16x3 K-major"), (3, "BF16x3 MN-major stacked"),
                                        (4, "BF16x3 MN-major, M- and N-stacked")])
def test_umma_selftest(cuda_backend, mode, label):
    """The descriptor / swizzle / instruction-descriptor conventions of the tcgen05 engine, in isolation:
    one 128x64x64 (or 128x64x128) product through the same issue helpers the jet kernels use."""
    torch.manual_seed(mode)
    mn = mode in (3, 4)
    a = torch.randn(128, 64, device="cuda")
    b = torch.randn(128 if mn else 64, 64, device="cuda")
    c = torch.zeros(128, 128 if mode == 4 else 64, device="cuda")
    cuda_backend.selftest_umma(mode, a, b, c)
    torch.cuda.synchronize()
    if mn:
        ref = a.double().t() @ b.double()
        got = c[:64].double() + c[64:].double()
        if mode == 4:
            got = got[:, :64] + got[:, 64:]
    else:
        ref = a.double() @ b.double().t()
        got = c.double()
    err = float((got - ref).abs().max() / ref.abs().max())
    fp32 = float(((a.t() @ b if mn else a @ b.t()).double() - ref).abs().max() / ref.abs().max())
    print(f"{label}: rel err {err:.2e} (cuBLAS fp32 {fp32:.2e})")
    assert err < 1e-6


def test_tcgen05_engine_is_used_and_required():
    """ENGINE_TCGEN05 must run the tensor-core kernels for an eligible program and refuse an ineligible one
    (no silent fallback to the FMA engine)."""
    m, _, _ = full_model("poisson_2D_1C", 5, _lib.ENGINE_TCGEN05)
    x = torch.rand((1000, 2), device="cuda") * 0.5
    out, stash = m.jet_engine.forward(x, 2, need_stash=True)
    ref_out, _ = oracle_jets(m, x.cpu().numpy(), 2)
    assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref_out) < TOL
    # 128-wide layers: the layer-wise tcgen05 engine (tests/test_gpu_wide.py) takes up to four jet channels; the full
    # second-order jet of a 2-D problem has five and is refused
    r, _, _ = full_model("burgers_1D", 5, _lib.ENGINE_TCGEN05)
    with pytest.raises(RuntimeError):
        r.jet_engine.forward(torch.rand((64, 2), device="cuda"), 2, need_stash=True)
    xb = torch.rand((500, 2), device="cuda")
    out, _ = r.jet_engine.forward(xb, 3, need_stash=True, lap=(1.0, 0.0))
    ref_out, _ = JR.forward(*AR.build_program(r), xb.cpu().numpy().astype(np.float64), 3, (1.0, 0.0))
    assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref_out) < TOL
    # residual blocks and Fourier-feature networks are covered by the tensor-core engine
    for name, dim in (("navier_stokes_2D", 2), ("navier_stokes_3D", 3)):
        r, _, _ = full_model(name, 5, _lib.ENGINE_TCGEN05)
        xr = torch.rand((777, dim), device="cuda")
        rng = np.random.default_rng(1)
        out, stash = r.jet_engine.forward(xr, 2, need_stash=True)
        dout_np = rng.standard_normal((777, 1 + 2 * dim, out.shape[2])).astype(np.float32)
        dflat = r.jet_engine.backward(xr, 2, torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda(),
                                      stash).cpu().numpy()
        ref_out, ref_dp = oracle_jets(r, xr.cpu().numpy(), 2, dout_np)
        assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref_out) < TOL, name
        assert relerr(dflat, ref_dp) < TOL, name


# ------------------------------------------------------------------------------ small kernels
def test_adam_kernel_matches_torch_adam(cuda_backend):
    torch.manual_seed(0)
    n = 33_537
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=1e-3)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        gr = torch.randn(n, device="cuda") * 10.0 ** float(torch.randint(-6, 2, (1,)))
        ref.grad = gr.clone()
        opt.step()
        cuda_backend.adam_step(p, gr, m, v, step, 1e-3, 0.9, 0.999, 1e-8, 1.0)
        assert float((p - ref.detach()).abs().max()) < 1e-6
    assert float((m - opt.state[ref]["exp_avg"]).abs().max()) < 1e-6 * float(m.abs().max())


@pytest.mark.parametrize("n,stride", [(1, 1), (255, 1), (100_000, 1), (100_000, 3), (1_000_000, 4)])
def test_sumsq_kernel(cuda_backend, n, stride):
    torch.manual_seed(n)
    base = torch.randn(n * stride, device="cuda")
    r = base[::stride]
    acc = torch.zeros(1, device="cuda")
    cuda_backend.sumsq(r, acc)
    ref = float(torch.sum(r.double() ** 2))
    assert abs(acc.item() - ref) / ref < 1e-5


# ------------------------------------------------------------------------------ last-layer bias gradient
@pytest.mark.parametrize("eng", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("n", [700, 5000, 70001])
def test_last_layer_bias_gradient_is_summed_in_double(n, eng):
    """dL/db of the last layer is the column sum of the value-channel seeds, usually the largest entry of the flat
    gradient (reference: loss.backward(), trainer/trainer.py:141).  csrc/api.cu sums it in double precision before
    the engine runs (value_seed_colsum; one CTA without atomics for small launches, double atomics otherwise), so it
    is exact to fp32 rounding even for seeds that cancel to 1e-4 of their magnitude, and identical from run to run."""
    m, din, dout_dim = full_model("poisson_2D_1C", 3, eng)
    e = m.jet_engine
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.rand((n, din), device="cuda", generator=g)
    out, stash = e.forward(x, 1, need_stash=True)
    dout = torch.randn(out.shape, device="cuda", generator=g) * 100.0
    dout[0] -= dout[0].mean(dim=0, keepdim=True) * (1.0 - 1e-4)          # heavy cancellation in the column sum
    exact = dout[0].double().sum(dim=0).cpu().numpy()
    off = sum(p.numel() for p in list(m.parameters())[:-1])
    g1 = e.backward(x, 1, dout, stash).cpu().numpy().copy()
    g2 = e.backward(x, 1, dout, stash).cpu().numpy().copy()
    got = g1[off:off + dout_dim].astype(np.float64)
    assert np.all(np.abs(got - exact) <= 1.2e-7 * np.abs(exact) + 1e-30), (got, exact)
    assert np.array_equal(g1[off:off + dout_dim], g2[off:off + dout_dim])
