"""GPU parity of the round-2 path (run on the B200 box: ``pytest -m gpu``): Laplace-mode jets, the residual-program
kernel, the fused / CUDA-graph train step, Adam with device-side state, and the regularizers on CUDA.  Everything
goes through the C ABI; expected values come from the reference goldens and the fp64 oracle."""
import os
import types

import numpy as np
import pytest
import torch

import multipinn_b200 as mp
from multipinn_b200 import _lib, engine, jet, trace
from multipinn_b200.fused import ConditionRunner
from oracle import autograd_ref as AR
from oracle import jet_ref as JR
from oracle import residual_ref as RR
from tests import problems
from tests.test_gpu_parity import ENGINE_IDS, ENGINES, TOL, full_model, golden_model, product_ns
from tests.test_oracle_golden import GOLD_KW, NAMES, load, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def cuda_backend():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    old = engine._set_backend_for_tests(None)
    jet.clear_blocks()
    be = engine.backend()
    assert be.name == "cuda"
    yield be
    engine._set_backend_for_tests(old)


def golden_pinn(name, g, kernel_engine=_lib.ENGINE_AUTO):
    model = golden_model(name, g, kernel_engine)
    spec, din, dout = problems.CONFIGS[name][0](product_ns())
    conds = []
    for i, (fn, (lo, hi)) in enumerate(spec):
        c = mp.Condition(fn, mp.Hypercube(lo, hi))
        pts = torch.tensor(g[f"pts{i}"]).cuda().requires_grad_()
        c.generator = types.SimpleNamespace(n_points=len(pts), generate=lambda cond, model, p=pts: p)
        conds.append(c)
    pinn = mp.PINN(model, conds)
    pinn.update_data()
    for c in conds:
        c.set_batching(1)
    return model, conds, pinn


LAPS = {1: (1.0,), 2: (1.0, 0.25), 3: (1.0, -0.5, 2.0)}


# ------------------------------------------------------------------------------ Laplace-mode jets
@pytest.mark.parametrize("eng", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("name", NAMES)
def test_laplace_mode_forward_and_reverse_match_oracle(name, eng):
    """order 3: value, first derivatives and L = sum_i c_i d2/dx_i2 against the fp64 oracle, and against the
    order-2 jets of the same kernels combined with the same weights."""
    g = load(name)
    m = golden_model(name, g, eng)
    x_np = g["pts0"]
    n, d = x_np.shape
    lap = LAPS[d]
    x = torch.tensor(x_np).cuda()
    e = m.jet_engine
    out3, stash = e.forward(x, 3, need_stash=True, lap=lap)
    assert out3.shape[0] == d + 2
    prog, flat = AR.build_program(m)
    ref3, cache = JR.forward(prog, flat, x_np.astype(np.float64), 3, lap)
    assert relerr(out3.cpu().numpy().transpose(1, 0, 2), ref3) < TOL
    out2, _ = e.forward(x, 2, need_stash=False)
    comb = sum(lap[i] * out2[1 + d + i] for i in range(d))
    assert relerr(out3[1 + d].cpu().numpy(), comb.cpu().numpy()) < TOL
    # same kernels at both orders -> identical up to rounding; a 128-wide net runs order 3 (four channels) on the layer-wise
    # tcgen05 engine and order 2 (five channels) on the FP32-FMA engine, two roundings of the same numbers
    same_kernels = eng == _lib.ENGINE_FFMA or e.tc_supported(2) == e.tc_supported(3)
    assert relerr(out3[:1 + d].cpu().numpy(), out2[:1 + d].cpu().numpy()) < (1e-6 if same_kernels else TOL)
    rng = np.random.default_rng(3)
    dout_np = rng.standard_normal((n, d + 2, int(g["output_dim"]))).astype(np.float32)
    dout = torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda()
    dflat = e.backward(x, 3, dout, stash, lap=lap).cpu().numpy()
    ref_dp, _ = JR.backward(prog, flat, x_np.astype(np.float64), 3, cache, dout_np.astype(np.float64), lap)
    if prog[0, 0] == JR.OP_FOURIER:
        ref_dp[prog[0, 3]: prog[0, 3] + prog[0, 1] * prog[0, 2] // 2] = 0
    assert relerr(dflat, ref_dp) < TOL


@pytest.mark.parametrize("name,dim", [("poisson_2D_1C", 2), ("navier_stokes_2D", 2), ("navier_stokes_3D", 3)])
def test_laplace_mode_full_depth_many_tiles(name, dim):
    """BASELINE depth, more tiles than CTAs, ragged tail, tcgen05 engine."""
    m, _, _ = full_model(name, 5, _lib.ENGINE_TCGEN05)
    n = 148 * 128 * 2 + 77
    lap = (1.0,) * dim
    rng = np.random.default_rng(9)
    x_np = rng.random((n, dim), dtype=np.float32) * 0.5
    x = torch.tensor(x_np).cuda()
    out, stash = m.jet_engine.forward(x, 3, need_stash=True, lap=lap)
    dout_np = rng.standard_normal((n, dim + 2, out.shape[2])).astype(np.float32)
    dflat = m.jet_engine.backward(x, 3, torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda(),
                                  stash, lap=lap).cpu().numpy()
    prog, flat = AR.build_program(m)
    ref, cache = JR.forward(prog, flat, x_np.astype(np.float64), 3, lap)
    ref_dp, _ = JR.backward(prog, flat, x_np.astype(np.float64), 3, cache, dout_np.astype(np.float64), lap)
    if prog[0, 0] == JR.OP_FOURIER:
        ref_dp[prog[0, 3]: prog[0, 3] + prog[0, 1] * prog[0, 2] // 2] = 0
    assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref) < TOL
    assert relerr(dflat, ref_dp) < TOL


def test_laplace_mode_property_at_1M_points():
    """BASELINE configs[1] size: the L channel equals u_xx + u_yy of the order-2 launch at every one of 1M points,
    and the reverse pass with seeds (g on L) equals the order-2 reverse pass with seeds (g on u_xx, g on u_yy)."""
    m, _, _ = full_model("poisson_2D_1C", 11)
    n = 1_000_000
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((n, 2), device="cuda", generator=gen) * 0.5
    e = m.jet_engine
    out3, st3 = e.forward(x, 3, need_stash=True, lap=(1.0, 1.0))
    out2, st2 = e.forward(x, 2, need_stash=True)
    lap2 = out2[3] + out2[4]
    scale = float(lap2.abs().max())
    assert float((out3[3] - lap2).abs().max()) / scale < TOL
    assert torch.equal(out3[0], out2[0])
    d3 = torch.randn(out3.shape, device="cuda", generator=gen)
    d2 = torch.cat([d3[:3], d3[3:4], d3[3:4]], dim=0).contiguous()
    g3 = e.backward(x, 3, d3, st3, lap=(1.0, 1.0))
    g2 = e.backward(x, 2, d2, st2)
    assert float((g3 - g2).abs().max() / g2.abs().max()) < 2e-5


# ------------------------------------------------------------------------------ residual-program kernel
@pytest.mark.parametrize("name", NAMES)
def test_residual_kernel_matches_oracle_interpreter(name, cuda_backend):
    """values, sums of squares and reverse-pass seeds of every traced example condition: device vs the numpy
    interpreter in fp64 on the same jets; the values also equal the reference's residual vectors (goldens)."""
    g = load(name)
    model, conds, pinn = golden_pinn(name, g)
    k = 0
    for ci, c in enumerate(conds):
        prog = c.program_for(model)
        assert prog is not None, c.trace_note
        x = torch.tensor(g[f"pts{ci}"]).cuda()
        n = x.shape[0]
        r = ConditionRunner(c, prog, model.jet_engine)
        r.bind(n, x.device)
        r.x.copy_(x)
        K = len(prog.outputs)
        w = torch.rand(K, device="cuda") + 0.5
        sums = torch.zeros(K, device="cuda")
        vals = torch.empty((K, n), device="cuda")
        if r.stacked:     # several model calls: one forward over the stacked points, the program reads slices
            B = len(r.blocks)
            for b in r.blocks:
                if b.points_prog is not None:
                    cuda_backend.residual_program(b.points_prog, r.x, [], None, None, None, b.x, point_major=True)
            model.jet_engine.forward(r.xcat, r.blocks[0].order, True, lap=r.blocks[0].lap, out=r.out_cat,
                                     stash=r.stash_cat)
            cuda_backend.residual_program(prog, r.x, [r.out_cat] * B, [r.dout_cat] * B, w, sums, vals,
                                          jet_offset=[i * n for i in range(B)], jet_stride=B * n)
            outs = [r.out_cat[:, i * n:(i + 1) * n] for i in range(B)]
            douts = [r.dout_cat[:, i * n:(i + 1) * n] for i in range(B)]
        else:
            r._forward_blocks(r.x, r.blocks, True)
            cuda_backend.residual_program(prog, r.x, [b.out for b in r.blocks], [b.dout for b in r.blocks], w, sums, vals)
            outs, douts = [b.out for b in r.blocks], [b.dout for b in r.blocks]
        jets = [o.cpu().numpy().astype(np.float64) for o in outs]
        rv, rs, rd = RR.evaluate(prog.instrs, prog.consts, prog.outputs, [b.o for b in prog.blocks],
                                 x.cpu().numpy().astype(np.float64), jets, w.cpu().numpy().astype(np.float64))
        assert relerr(vals.cpu().numpy(), rv) < 2e-6
        assert relerr(sums.cpu().numpy(), rs) < 2e-6
        for dj, ref in zip(douts, rd):
            assert relerr(dj.cpu().numpy(), ref) < 2e-6
        for kk in range(K):
            assert relerr(vals[kk].cpu().numpy(), g[f"res{ci}_{k}_f64"]) < TOL, (ci, k)
            k += 1


def test_residual_kernel_every_opcode(cuda_backend):
    """A synthetic program touching every opcode, ragged sizes, against the numpy interpreter."""
    model = mp.FNN(2, 2, [64, 64, 64]).cuda()

    def fn(m, a):
        u, v = mp.unpack(m(a))
        x, y = mp.unpack(a)
        u_x, u_y = mp.unpack(mp.grad(u, a))
        t = torch.tanh(u) * torch.exp(-v) + torch.abs(u_x) / (2.0 + torch.cos(x) ** 2)
        s = torch.sqrt(y * y + 1.0) - torch.log(x + 1.5) + torch.relu(u_y) - (-v) ** 3 + (u * u + 1.0) ** 0.75
        return [t + torch.sin(s), s - 0.5 * t, u / (1.0 + v * v)]

    prog = trace.trace_residual(fn, model, 2, torch.zeros(1, 2, device="cuda"))
    ops = {i[0] for i in prog.instrs}
    for need in (trace.OP_DIV, trace.OP_NEG, trace.OP_SIN, trace.OP_COS, trace.OP_EXP, trace.OP_TANH, trace.OP_RELU,
                 trace.OP_ABS, trace.OP_SQRT, trace.OP_LOG, trace.OP_POWI, trace.OP_POWF):
        assert need in ops
    c = mp.Condition(fn, mp.Hypercube([0, 0], [1, 1]))
    for n in (1, 127, 128, 129, 5000):
        x = torch.rand((n, 2), device="cuda")
        out, _ = model.jet_engine.forward(x, 1, need_stash=False)
        dout = torch.empty_like(out)
        K = 3
        w = torch.tensor([0.3, 1.0, 2.5], device="cuda")
        sums = torch.zeros(K, device="cuda")
        vals = torch.empty((K, n), device="cuda")
        cuda_backend.residual_program(prog, x, [out], [dout], w, sums, vals)
        rv, rs, rd = RR.evaluate(prog.instrs, prog.consts, prog.outputs, [2], x.cpu().numpy().astype(np.float64),
                                 [out.cpu().numpy().astype(np.float64)], w.cpu().numpy().astype(np.float64))
        assert relerr(vals.cpu().numpy(), rv) < 5e-6, n
        assert relerr(sums.cpu().numpy(), rs) < 5e-6, n
        assert relerr(dout.cpu().numpy(), rd[0]) < 5e-6, n
        # value-only launch (no seeds), point-major layout
        vt = torch.empty((n, K), device="cuda")
        cuda_backend.residual_program(prog, x, [out], None, None, None, vt, point_major=True)
        assert torch.equal(vt.t().contiguous(), vals)


# ------------------------------------------------------------------------------ fused step
@pytest.mark.parametrize("name", NAMES)
def test_fused_graph_step_matches_reference_golden(name):
    """TrainerOneBatch on the fused path with CUDA graphs: losses, gradient and the 3-step Adam trajectory of the
    unmodified reference; two more steps are graph replays and must keep agreeing with the eager jet path."""
    g = load(name)
    model, conds, pinn = golden_pinn(name, g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
    tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=5, update_grid_every=10**9)
    fs = tr.fused_step()
    assert fs is not None and tr._flat_adam is not None
    # eager twin on identical weights
    model_e, conds_e, pinn_e = golden_pinn(name, g)
    opt_e = torch.optim.Adam(model_e.parameters(), lr=1e-3)
    sched_e = torch.optim.lr_scheduler.ExponentialLR(opt_e, gamma=0.9999)
    tr_e = mp.TrainerOneBatch(pinn_e, opt_e, sched_e, num_epochs=5, update_grid_every=10**9)
    tr_e.use_fused_step = False
    traj = []
    for e in range(5):
        tr.current_epoch = tr_e.current_epoch = e + 1
        tr._train_epoch()
        tr_e._train_epoch()
        sched.step()
        sched_e.step()
        traj.append(float(tr.total_loss))
        if e == 0:
            assert relerr(tr.epoch_loss_detailed.cpu().numpy(), g["losses_f64"]) < TOL
            assert relerr(model.jet_engine.flat_grad.cpu().numpy(), g["grad_f64"]) < TOL
            assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - g["params_after1"]).max() < 5e-6
        if e == 2:
            assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - g["params_after3"]).max() < 1e-5
        assert relerr(float(tr.total_loss), float(tr_e.total_loss)) < 2e-5, e
        assert relerr(tr.epoch_loss_detailed.cpu().numpy(), tr_e.epoch_loss_detailed.cpu().numpy()) < 2e-5, e
    assert relerr(np.array(traj[:3]), g["traj_f32"]) < 2e-5
    assert fs.graph_replays >= 4            # steps 3..5: compute + update graphs
    d = (model.jet_engine.flat.detach() - model_e.jet_engine.flat.detach()).abs().max()
    assert float(d) < 2e-5
    p0 = [p for p in model.parameters() if p.requires_grad][0]
    assert float(opt.state[p0]["step"]) == 5.0


def test_fused_step_follows_new_points_and_lr(cuda_backend):
    """A captured step must see points regenerated every epoch (update_grid_every=1, the reference default) and a
    learning rate changed by the scheduler."""
    torch.manual_seed(0)
    model = mp.FNN(2, 1, [64, 64, 64]).cuda()
    model_e = mp.FNN(2, 1, [64, 64, 64]).cuda()
    model_e.load_state_dict(model.state_dict())

    def inner(m, arg):
        (u,) = mp.unpack(m(arg))
        xx, yy = mp.unpack(arg)
        u_x, u_y = mp.unpack(mp.grad(u, arg))
        u_xx, _ = mp.unpack(mp.grad(u_x, arg))
        _, u_yy = mp.unpack(mp.grad(u_y, arg))
        return [u_xx + u_yy + torch.sin(xx) * yy]

    trainers = []
    for mdl, fused in ((model, True), (model_e, False)):
        cond = mp.Condition(inner, mp.Hypercube([0, 0], [1, 1]))
        pts = [torch.rand((700, 2), generator=torch.Generator().manual_seed(s)).cuda().requires_grad_() for s in range(6)]
        cond.generator = types.SimpleNamespace(n_points=700, it=iter(pts),
                                               generate=lambda c, m, it=iter(pts): next(it))
        pinn = mp.PINN(mdl, [cond])
        opt = torch.optim.Adam(mdl.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
        tr = mp.TrainerOneBatch(pinn, opt, sched, num_epochs=6, update_grid_every=1)
        tr.use_fused_step = fused
        trainers.append((tr, sched))
    for e in range(6):
        for tr, sched in trainers:
            tr.current_epoch = e
            tr._train_epoch()
            sched.step()
        a, b = trainers[0][0], trainers[1][0]
        assert relerr(float(a.total_loss), float(b.total_loss)) < 2e-5, e
    assert trainers[0][0].fused_step().graph_replays >= 6
    assert float((model.jet_engine.flat.detach() - model_e.jet_engine.flat.detach()).abs().max()) < 2e-5


def test_adam_device_state_matches_torch_adam(cuda_backend):
    torch.manual_seed(0)
    n = 33_537
    p = torch.randn(n, device="cuda")
    ref = p.clone().requires_grad_()
    opt = torch.optim.Adam([ref], lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    state = torch.zeros(4, device="cuda")
    for step in range(1, 8):
        gr = torch.randn(n, device="cuda") * 10.0 ** float(torch.randint(-6, 2, (1,)))
        ref.grad = gr.clone()
        state[0] = opt.param_groups[0]["lr"]
        opt.step()
        sched.step()
        cuda_backend.adam_tick(state, 0.9, 0.999)
        cuda_backend.adam_step_dev(p, gr, m, v, state, 0.9, 0.999, 1e-8, 1.0)
        assert float(state[1]) == float(step)
        assert float((p - ref.detach()).abs().max()) < 1e-6
    assert float((m - opt.state[ref]["exp_avg"]).abs().max()) < 1e-6 * float(m.abs().max())


# ------------------------------------------------------------------------------ C ABI corner cases
def test_forward_without_stash_through_the_c_abi(cuda_backend):
    """stash = NULL is a value-only pass: the FP32-FMA engine runs it at any order, the tcgen05 engine at order 0
    (inference: nothing is written for a reverse pass) for programs without skip connections; an explicit tcgen05
    request it cannot serve is refused with a message (never an uninitialised return code)."""
    m, _, _ = full_model("poisson_2D_1C", 2)
    e = m.jet_engine
    x = torch.rand((300, 2), device="cuda")
    ref, _ = e.forward(x, 2, need_stash=True)
    ops = m.jet_program()
    for eng_id in (_lib.ENGINE_FFMA, _lib.ENGINE_AUTO):
        out = torch.empty_like(ref)
        cuda_backend.jet_forward(ops, e.flat.detach(), e.packed, x, 2, out, None, eng_id)
        assert relerr(out.cpu().numpy(), ref.cpu().numpy()) < TOL
    with pytest.raises(RuntimeError, match="stash"):
        cuda_backend.jet_forward(ops, e.flat.detach(), e.packed, x, 2, torch.empty_like(ref), None, _lib.ENGINE_TCGEN05)
    out = torch.empty_like(ref[:1])
    cuda_backend.jet_forward(ops, e.flat.detach(), e.packed, x, 0, out, None, _lib.ENGINE_TCGEN05)
    assert torch.equal(out[0], ref[0])   # the same tensor-core arithmetic, no stash
    r, _, _ = full_model("navier_stokes_2D", 2)
    xr = torch.rand((300, 2), device="cuda")
    refr, _ = r.jet_engine.forward(xr, 1, need_stash=True)
    with pytest.raises(RuntimeError, match="stash"):
        cuda_backend.jet_forward(r.jet_program(), r.jet_engine.flat.detach(), r.jet_engine.packed, xr, 0,
                                 torch.empty_like(refr[:1]), None, _lib.ENGINE_TCGEN05)
    with torch.no_grad():                # the host allocates the scratch for residual-block inference itself
        assert torch.equal(r(xr), refr[0])
    with pytest.raises(RuntimeError, match="lap_coef"):
        cuda_backend.jet_forward(ops, e.flat.detach(), e.packed, x, 3, torch.empty((4, 300, 1), device="cuda"), None)
    assert cuda_backend.engine_supports(ops, 2, 3, _lib.ENGINE_TCGEN05)
    b, _, _ = full_model("burgers_1D", 2)
    assert not cuda_backend.engine_supports(b.jet_program(), 2, 2, _lib.ENGINE_TCGEN05)
    # a 128-wide program (FMA engine only) needs no stash for inference
    with torch.no_grad():
        y = b(torch.rand((64, 2), device="cuda"))
    assert y.shape == (64, 1)


def test_reloading_weights_invalidates_packed_operands():
    """load_state_dict / in-place edits after a trainer took over the engine must reach the kernels."""
    g = load("poisson_2D_1C")
    model, conds, pinn = golden_pinn("poisson_2D_1C", g)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    tr = mp.TrainerOneBatch(pinn, torch.optim.Adam(model.parameters(), lr=1e-2), None, num_epochs=2,
                            update_grid_every=10**9)
    x = torch.tensor(g["pts0"]).cuda()
    with torch.no_grad():
        y0 = model(x).clone()
    tr.current_epoch = 1
    tr._train_epoch()
    with torch.no_grad():
        y1 = model(x).clone()
    assert not torch.equal(y0, y1)
    model.load_state_dict(sd0)
    with torch.no_grad():
        assert torch.equal(model(x), y0)
    # in-place re-initialisation behind the trainer's back (MarchingTrainer.reset_weights, reference
    # trainer/marching.py:35-40): noticed through the parameters' version counters, operands re-packed, and the next
    # fused step (a graph replay by now) trains the NEW weights
    for e in (2, 3, 4):
        tr.current_epoch = e
        tr._train_epoch()
    torch.manual_seed(7)
    for mod in model.modules():
        if isinstance(mod, torch.nn.Linear):
            mod.reset_parameters()
    with torch.no_grad():
        y2 = model(x).clone()
    ref = AR.MODELS["FNN"](2, 1, hidden_layers=[64] * 9).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in model.state_dict().items()})
    assert relerr(y2.cpu().numpy(), ref(torch.tensor(g["pts0"]).double()).detach().numpy()) < 1e-5
    before = model.jet_engine.flat.detach().clone()
    tr.current_epoch = 5
    tr._train_epoch()
    step = (model.jet_engine.flat.detach() - before).abs().max()
    assert 0 < float(step) < 2e-2          # one Adam step of size ~lr from the re-initialised weights, not a jump back


# ------------------------------------------------------------------------------ regularizers on CUDA
@pytest.mark.parametrize("kind", ["normal", "pointwise", "adaptive"])
def test_regularizers_match_reference_on_cuda(kind):
    """NormalLosses / PointwiseLosses / AdaptiveConditionsLosses over 3 steps vs the reference's own run (tests/golden),
    on the device, as loss policies of the fused step."""
    g = load("poisson_2D_1C")
    gr = np.load(os.path.join(os.path.dirname(__file__), "golden", f"poisson_{kind}_reg.npz"))
    model, conds, pinn = golden_pinn("poisson_2D_1C", g)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    reg = {"normal": lambda: mp.NormalLosses(alpha=0.9),
           "pointwise": lambda: mp.PointwiseLosses(ignored_indexes=[1], lr=0.01),
           "adaptive": lambda: mp.AdaptiveConditionsLosses(lr=0.05)}[kind]()
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=3, update_grid_every=10**9, calc_loss=reg)
    assert tr.fused_step() is not None      # both are loss policies of the fused step (the third step is a graph replay)
    for e in range(3):
        tr.current_epoch = e + 1
        tr._train_epoch()
        if kind == "normal":
            assert relerr(reg.lambda_regularization.cpu().numpy(), gr["lambdas"][e]) < 5e-4
        elif kind == "adaptive":
            assert relerr(reg.adaptive_cond.weight.detach().cpu().numpy(), gr["lambdas"][e]) < 2e-5
        else:
            ms = reg.models[0]
            lam = np.concatenate([ms[0].weight.detach().cpu().numpy().reshape(-1)[:16],
                                  ms[-1].weight.detach().cpu().numpy().reshape(-1)[:16]])
            assert relerr(lam, gr["lambdas"][e]) < 2e-5
        assert relerr(float(tr.total_loss), gr["totals"][e]) < 2e-5
        assert relerr(tr.epoch_loss_detailed.cpu().numpy(), gr["losses"][e]) < 2e-5
    assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - gr["params_after3"]).max() < 2e-5


@pytest.mark.parametrize("name", ["poisson_2D_1C", "navier_stokes_2D", "navier_stokes_3D"])
def test_points_grad_matches_reference_backward_on_cuda(name):
    g = load(name)
    model, conds, pinn = golden_pinn(name, g)
    model.jet_engine.points_grad = True
    pinn.select_batch(0)
    total, _ = pinn.calculate_mean_loss()
    total.backward()
    fn, kind, _ = problems.CONFIGS[name]
    spec, din, dout = fn(types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None))
    ref = AR.MODELS[kind](int(g["input_dim"]), int(g["output_dim"]), **GOLD_KW[name]).double()
    ref.load_state_dict({k: torch.tensor(g["w::" + k]).double() for k in g["state_keys"]})
    pts = [torch.tensor(g[f"pts{i}"]).double().requires_grad_() for i in range(len(spec))]
    res = [r for (f, _), x in zip(spec, pts) for r in f(ref, x)]
    (sum((r ** 2).sum() for r in res) / sum(r.numel() for r in res)).backward()
    for c, x in zip(conds, pts):
        assert c.points.grad is not None
        assert relerr(c.points.grad.cpu().numpy(), x.grad.numpy()) < TOL


# ------------------------------------------------------------------------------ full-Hessian jets (mixed partials)
@pytest.mark.parametrize("name", ["poisson_2D_1C", "navier_stokes_2D", "navier_stokes_3D", "burgers_1D"])
def test_full_hessian_jets_match_oracle(name, cuda_backend):
    """order 4: value, first, diagonal AND mixed second derivatives (FP32-FMA engine) against the fp64 oracle,
    forward and reverse; the first 1 + 2d channels equal the order-2 jets; an explicit tcgen05 request is refused."""
    g = load(name)
    m = golden_model(name, g)
    x_np = g["pts0"]
    n, d = x_np.shape
    x = torch.tensor(x_np).cuda()
    e = m.jet_engine
    out4, stash = e.forward(x, 4, need_stash=True)
    J = 1 + 2 * d + d * (d - 1) // 2
    assert out4.shape[0] == J
    prog, flat = AR.build_program(m)
    ref, cache = JR.forward(prog, flat, x_np.astype(np.float64), 4)
    assert relerr(out4.cpu().numpy().transpose(1, 0, 2), ref) < TOL
    out2, _ = e.forward(x, 2, need_stash=False)
    assert relerr(out4[:1 + 2 * d].cpu().numpy(), out2.cpu().numpy()) < TOL
    rng = np.random.default_rng(4)
    dout_np = rng.standard_normal((n, J, int(g["output_dim"]))).astype(np.float32)
    dout = torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda()
    dflat = e.backward(x, 4, dout, stash).cpu().numpy()
    ref_dp, _ = JR.backward(prog, flat, x_np.astype(np.float64), 4, cache, dout_np.astype(np.float64))
    if prog[0, 0] == JR.OP_FOURIER:
        ref_dp[prog[0, 3]: prog[0, 3] + prog[0, 1] * prog[0, 2] // 2] = 0
    assert relerr(dflat, ref_dp) < TOL
    assert not cuda_backend.engine_supports(m.jet_program(), d, 4, _lib.ENGINE_TCGEN05)


def test_mixed_partial_residual_trains_on_the_fused_step():
    """A residual that uses u_xy: traced into an order-4 program, fused CUDA-graph step vs the eager jet path and vs
    nested autograd of the reference algorithm in fp64."""
    torch.manual_seed(3)

    def make(ns):
        def fn(m, a):
            (u,) = ns.unpack(m(a))
            x, y = ns.unpack(a)
            u_x, u_y = ns.unpack(ns.grad(u, a))
            u_xx, u_xy = ns.unpack(ns.grad(u_x, a))
            _, u_yy = ns.unpack(ns.grad(u_y, a))
            return [u_xx * u_yy - u_xy**2 - torch.exp(x) * y, u_x + u_xy]
        return fn

    pts = torch.rand((3000, 2), generator=torch.Generator().manual_seed(0))
    trainers = []
    for fused in (True, False):
        torch.manual_seed(3)
        model = mp.FNN(2, 1, [64, 64, 64]).cuda()
        cond = mp.Condition(make(types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack)), mp.Hypercube([0, 0], [1, 1]))
        p = pts.cuda().requires_grad_()
        cond.generator = types.SimpleNamespace(n_points=3000, generate=lambda c, m, p=p: p)
        pinn = mp.PINN(model, [cond])
        pinn.update_data()
        tr = mp.TrainerOneBatch(pinn, torch.optim.Adam(model.parameters(), lr=1e-3), None, num_epochs=4,
                                update_grid_every=10**9)
        tr.use_fused_step = fused
        assert (tr.fused_step() is not None) == fused
        assert cond.plan_for(model).orders == [4] and cond.program_for(model).blocks[0].order == 4
        trainers.append((tr, model))
    ref = AR.MODELS["FNN"](2, 1, hidden_layers=[64, 64, 64]).double().cuda()
    ref.load_state_dict({k: v.double() for k, v in trainers[0][1].state_dict().items()})
    total, losses = AR.mean_loss(ref, [make(types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack))],
                                 [pts.double().cuda().requires_grad_()])
    for e in range(4):
        for tr, _ in trainers:
            tr.current_epoch = e + 1
            tr._train_epoch()
        a, b = trainers[0][0], trainers[1][0]
        if e == 0:
            assert relerr(a.epoch_loss_detailed.cpu().numpy(), losses.detach().cpu().numpy()) < TOL
        assert relerr(float(a.total_loss), float(b.total_loss)) < 2e-5
    d = (trainers[0][1].jet_engine.flat.detach() - trainers[1][1].jet_engine.flat.detach()).abs().max()
    assert float(d) < 2e-5


@pytest.mark.parametrize("order,lap", [(0, None), (1, None), (2, None), (3, (1.0, 0.5, 2.0)), (4, None)])
def test_trainable_fourier_matrix_gradient_matches_oracle(order, lap, cuda_backend):
    """FourierFeatureNetwork(is_trainable=True): dL/dB from the FP32-FMA reverse pass vs the fp64 oracle, every jet
    order; the tcgen05 engine declines the program (it treats B as frozen)."""
    torch.manual_seed(7)
    m = mp.FFNN(3, 4, hidden_layers=[64, 64, 64], encoding_dim=64, sigma=1.0, is_trainable=True).cuda()
    assert not cuda_backend.engine_supports(m.jet_program(), 3, 2, _lib.ENGINE_TCGEN05)
    n = 777
    rng = np.random.default_rng(order)
    x_np = rng.random((n, 3), dtype=np.float32)
    x = torch.tensor(x_np).cuda()
    e = m.jet_engine
    out, stash = e.forward(x, order, need_stash=True, lap=lap)
    dout_np = rng.standard_normal((n, out.shape[0], 4)).astype(np.float32)
    dflat = e.backward(x, order, torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda(), stash,
                       lap=lap).cpu().numpy()
    prog, flat = AR.build_program(m)
    assert prog[0, 6] == 4
    ref, cache = JR.forward(prog, flat, x_np.astype(np.float64), order, lap)
    ref_dp, _ = JR.backward(prog, flat, x_np.astype(np.float64), order, cache, dout_np.astype(np.float64), lap)
    nB = 3 * 32
    assert relerr(out.cpu().numpy().transpose(1, 0, 2), ref) < TOL
    assert relerr(dflat, ref_dp) < TOL
    assert relerr(dflat[:nB], ref_dp[:nB]) < TOL and np.abs(ref_dp[:nB]).max() > 0


# ------------------------------------------------------------------------------ per-point data (LOADD), 8 blocks
@pytest.mark.gpu
def test_residual_kernel_reads_per_point_data_columns(cuda_backend):
    """ConditionExtra data (reference condition/condition.py:117-215) as LOADD columns of a packed [N, n_data] block:
    kernel vs numpy interpreter, ragged sizes, seeds included."""
    model = mp.FNN(2, 1, [64, 64, 64]).cuda()

    def fn(m, a, data):
        (u,) = mp.unpack(m(a))
        normal, scale = data
        u_x, u_y = mp.unpack(mp.grad(u, a))
        return [normal[:, 0] * u_x + normal[:, 1] * u_y - scale, u * scale]

    sample = (torch.zeros(1, 2, device="cuda"), torch.zeros(1, device="cuda"))
    prog = trace.trace_residual(fn, model, 2, torch.zeros(1, 2, device="cuda"), sample)
    assert prog.data_width == 3 and prog.struct().n_data == 3
    for n in (1, 129, 4000):
        x = torch.rand((n, 2), device="cuda")
        items = (torch.randn((n, 2), device="cuda", dtype=torch.float64), torch.rand(n, device="cuda"))
        data = trace.pack_data(items, n)
        assert data.shape == (n, 3) and data.dtype == torch.float32
        out, _ = model.jet_engine.forward(x, 1, need_stash=False)
        dout = torch.empty_like(out)
        w = torch.tensor([0.7, 1.3], device="cuda")
        sums = torch.zeros(2, device="cuda")
        vals = torch.empty((2, n), device="cuda")
        cuda_backend.residual_program(prog, x, [out], [dout], w, sums, vals, data=data)
        rv, rs, rd = RR.evaluate(prog.instrs, prog.consts, prog.outputs, [1], x.cpu().numpy().astype(np.float64),
                                 [out.cpu().numpy().astype(np.float64)], w.cpu().numpy().astype(np.float64),
                                 data=data.cpu().numpy().astype(np.float64))
        assert relerr(vals.cpu().numpy(), rv) < 5e-6, n
        assert relerr(sums.cpu().numpy(), rs) < 5e-6, n
        assert relerr(dout.cpu().numpy(), rd[0]) < 5e-6, n
    with pytest.raises(AssertionError):
        cuda_backend.residual_program(prog, x, [out], [dout], w, sums, vals)        # data block missing


@pytest.mark.gpu
def test_condition_extra_trains_on_the_fused_step():
    """A Neumann-type ConditionExtra next to a plain condition: fused CUDA step (graphs on) against nested autograd of
    the reference algorithm (oracle port, fp64) -- losses, and weights after 3 Adam steps."""
    import types

    from oracle import autograd_ref as AR
    torch.manual_seed(5)
    model = mp.FNN(2, 1, [64, 64, 64]).cuda()

    def normals(points):
        p = points.detach()
        v = torch.stack([torch.cos(3.0 * p[:, 0]), torch.sin(2.0 * p[:, 1]) + 0.5], dim=1)
        return v / v.norm(dim=1, keepdim=True)

    def make_bc(ns):
        def fn(m, a, data):
            (u,) = ns.unpack(m(a))
            u_x, u_y = ns.unpack(ns.grad(u, a))
            return [data[0][:, 0] * u_x + data[0][:, 1] * u_y - 0.3]
        return fn

    def make_in(ns):
        def fn(m, a):
            (u,) = ns.unpack(m(a))
            x, y = ns.unpack(a)
            u_x, u_y = ns.unpack(ns.grad(u, a))
            (u_xx, _) = ns.unpack(ns.grad(u_x, a))
            (_, u_yy) = ns.unpack(ns.grad(u_y, a))
            return [u_xx + u_yy - torch.sin(x) * y]
        return fn

    ns = types.SimpleNamespace(grad=mp.grad, unpack=mp.unpack)
    geo = mp.Hypercube([0, 0], [1, 1])
    c_in, c_bc = mp.Condition(make_in(ns), geo), mp.ConditionExtra(make_bc(ns), geo, [normals])
    gen = torch.Generator().manual_seed(0)
    pts = [torch.rand((300, 2), generator=gen).cuda().requires_grad_(),
           torch.rand((77, 2), generator=gen).cuda().requires_grad_()]
    for c, p in zip((c_in, c_bc), pts):
        c.generator = types.SimpleNamespace(n_points=len(p), generate=lambda cc, m, p=p: p)
    pinn = mp.PINN(model, [c_in, c_bc])
    pinn.update_data()
    ref = AR.MODELS["FNN"](2, 1, hidden_layers=[64, 64, 64]).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in model.state_dict().items()})
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=4, update_grid_every=10**9)
    fs = tr.fused_step()
    assert fs is not None and c_bc.program_for(model).data_width == 2, c_bc.trace_note
    rns = types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack)
    p64 = [p.detach().cpu().double().requires_grad_() for p in pts]
    dat = (normals(pts[1]).cpu().double(),)
    rfns = [make_in(rns), lambda m, a: make_bc(rns)(m, a, dat)]
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    for step in range(4):       # two direct runs, then graph replays
        pinn.select_batch(0)
        tr._train_batches()
        total, losses = AR.train_step(ref, ropt, rfns, p64)
        assert relerr(tr.epoch_loss_detailed.cpu().numpy(), np.asarray([float(v) for v in losses])) < 2e-5, step
        tr.total_loss, tr.epoch_loss_detailed = 0, None
    assert fs.graph_replays >= 1
    rflat = np.concatenate([p.detach().numpy().reshape(-1) for p in ref.parameters()])
    assert np.abs(model.jet_engine.flat.detach().cpu().numpy() - rflat).max() < 1e-5


@pytest.mark.gpu
def test_seven_model_calls_in_one_residual_are_stacked(cuda_backend):
    """Central second differences in three directions (what Symbols(mode="diff") generates for a 3-D Laplacian,
    reference condition/symbols.py:188-234, with fixed steps): 7 model calls, one stacked forward and reverse."""
    torch.manual_seed(2)
    model = mp.FNN(3, 1, [64, 64, 64]).cuda()
    h = 2.0 ** -3

    def fn(m, a):
        f0 = m(a)
        tot = -6.0 * f0[:, 0]
        for i in range(3):
            e = torch.zeros(1, 3, device=a.device)
            e[0, i] = h
            tot = tot + m(a + e)[:, 0] + m(a - e)[:, 0]
        return [tot / (h * h)]

    c = mp.Condition(fn, mp.Hypercube([0, 0, 0], [1, 1, 1]))
    c.init_output_len(model)
    prog = c.program_for(model)
    assert prog is not None and len(prog.blocks) == 7, c.trace_note
    from multipinn_b200.fused import ConditionRunner
    r = ConditionRunner(c, prog, model.jet_engine)
    assert r.stacked
    x = torch.rand((1000, 3), device="cuda")
    got = r.residual_values(x)[0]
    with torch.no_grad():
        want = fn(model, x)[0]     # eager value-only evaluation of the same stencil
    assert relerr(got.cpu().numpy(), want.detach().cpu().numpy()) < 1e-4


@pytest.mark.gpu
def test_residual_kernel_masks_select_max_min(cuda_backend):
    """MAX / MIN / GT / GE / SEL (masked assignment, torch.where, clamp) against the numpy interpreter, seeds included."""
    model = mp.FNN(2, 1, [64, 64, 64]).cuda()

    def fn(m, a):
        (u,) = mp.unpack(m(a))
        x, y = mp.unpack(a)
        p = 0.5 * torch.exp(-x / 0.7)
        p[p > 0.35] = 0
        w = torch.where(y <= 0.5, u * 2.0, u - 1.0)
        c = torch.clamp(u, min=-0.05, max=0.08) + torch.maximum(u, 0.1 * x) - torch.minimum(u, y - 0.5)
        return [u - p, w, c + torch.where(x >= y, u, -u)]

    prog = trace.trace_residual(fn, model, 2, torch.zeros(1, 2, device="cuda"))
    ops = {i[0] for i in prog.instrs}
    assert {trace.OP_GT, trace.OP_GE, trace.OP_SEL, trace.OP_MAX, trace.OP_MIN} <= ops
    for n in (1, 130, 5000):
        x = torch.rand((n, 2), device="cuda")
        out, _ = model.jet_engine.forward(x, 0, need_stash=False)
        dout = torch.empty_like(out)
        w = torch.tensor([0.3, 1.0, 2.5], device="cuda")
        sums = torch.zeros(3, device="cuda")
        vals = torch.empty((3, n), device="cuda")
        cuda_backend.residual_program(prog, x, [out], [dout], w, sums, vals)
        # the interpreter runs in fp32 here: masks must flip at exactly the same points
        rv, rs, rd = RR.evaluate(prog.instrs, prog.consts, prog.outputs, [1], x.cpu().numpy(),
                                 [out.cpu().numpy()], w.cpu().numpy().astype(np.float64))
        assert relerr(vals.cpu().numpy(), rv) < 5e-6, n
        assert relerr(sums.cpu().numpy(), rs) < 5e-6, n
        assert relerr(dout.cpu().numpy(), rd[0]) < 5e-6, n


@pytest.mark.gpu
def test_residual_kernel_per_point_loss_weights(cuda_backend):
    """point_weights / wsums of mpinn_residual_program (PointwiseLosses: loss = sum_k sum_n lambda_kn r_kn^2):
    seeds, weighted and unweighted sums against the numpy interpreter."""
    model = mp.FNN(2, 2, [64, 64, 64]).cuda()

    def fn(m, a):
        u, v = mp.unpack(m(a))
        x, y = mp.unpack(a)
        u_x, u_y = mp.unpack(mp.grad(u, a))
        return [u_x + v * torch.sin(x), u_y - u * y]

    prog = trace.trace_residual(fn, model, 2, torch.zeros(1, 2, device="cuda"))
    for n in (1, 131, 6000):
        x = torch.rand((n, 2), device="cuda")
        out, _ = model.jet_engine.forward(x, 1, need_stash=False)
        dout = torch.empty_like(out)
        lam = torch.rand((2, n), device="cuda") * 3.0
        sums, wsums = torch.zeros(2, device="cuda"), torch.zeros(2, device="cuda")
        vals = torch.empty((2, n), device="cuda")
        cuda_backend.residual_program(prog, x, [out], [dout], None, sums, vals, point_weights=lam, wsums=wsums)
        rv, rs, rd = RR.evaluate(prog.instrs, prog.consts, prog.outputs, [2], x.cpu().numpy().astype(np.float64),
                                 [out.cpu().numpy().astype(np.float64)], None,
                                 point_weights=lam.cpu().numpy().astype(np.float64))
        assert relerr(vals.cpu().numpy(), rv) < 5e-6, n
        assert relerr(sums.cpu().numpy(), rs) < 5e-6, n
        assert relerr(wsums.cpu().numpy(), (lam.cpu().numpy().astype(np.float64) * rv ** 2).sum(axis=1)) < 5e-6, n
        assert relerr(dout.cpu().numpy(), rd[0]) < 5e-6, n


@pytest.mark.gpu
def test_multi_batch_epochs_on_cuda_graphs():
    """Trainer(num_batches=2): two captured compute graphs (overwrite / accumulate) + one update graph per epoch, against
    the eager jet path over 5 epochs."""
    runs = []
    for fused in (True, False):
        g = load("poisson_2D_1C")
        model, conds, pinn = golden_pinn("poisson_2D_1C", g)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        tr = mp.Trainer(pinn, opt, None, num_epochs=5, num_batches=2, update_grid_every=10**9)
        tr.use_fused_step = fused
        hist = []
        for e in range(1, 6):
            tr.current_epoch = e
            tr._train_epoch()
            hist.append(float(tr.total_loss))
        if fused:
            fs = tr.fused_step()
            assert fs is not None and fs._g_compute is not None and fs._g_compute_acc is not None
            assert fs.graph_replays >= 6
        runs.append((np.array(hist), model.jet_engine.flat.detach().cpu().numpy().copy()))
    (h0, w0), (h1, w1) = runs
    assert relerr(h0, h1) < 2e-5
    assert np.abs(w0 - w1).max() < 1e-5
