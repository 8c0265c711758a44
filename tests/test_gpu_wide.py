"""The layer-wise tcgen05 engine for 128-wide FNN programs (csrc/jet_tc_wide.cu; burgers_1D, reference
examples/burgers_1D/configs/config.yaml:7) against the fp64 oracle and against the FP32-FMA engine, through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from multipinn_b200 import _lib  # noqa: E402
import multipinn_b200 as mp  # noqa: E402
from oracle import autograd_ref as AR  # noqa: E402
from oracle import jet_ref as JR  # noqa: E402

TOL = 1e-5


def relerr(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def run(model, x, order, lap, dout_np):
    e = model.jet_engine
    out, stash = e.forward(x, order, need_stash=True, lap=lap)
    dout = torch.tensor(np.ascontiguousarray(dout_np.transpose(1, 0, 2))).cuda()
    dflat = e.backward(x, order, dout, stash, lap=lap)
    return out.cpu().numpy().transpose(1, 0, 2), dflat.cpu().numpy()


@pytest.mark.parametrize("d,order,lap", [(2, 3, (1.0, 0.0)), (2, 3, (0.5, 2.0)), (2, 1, None), (2, 0, None),
                                         (1, 2, None), (1, 3, (1.5,)), (3, 1, None)])
@pytest.mark.parametrize("n", [1, 33, 300, 4097])
def test_wide_engine_matches_oracle(d, order, lap, n):
    torch.manual_seed(3)
    n_out = 1 if d != 3 else 3
    m = mp.FNN(d, n_out, [128, 128, 128, 128]).cuda()
    m.jet_engine.kernel_engine = _lib.ENGINE_TCGEN05          # refuses instead of falling back
    rng = np.random.default_rng(n + 10 * order)
    x_np = rng.random((n, d)).astype(np.float32)
    J = m.jet_engine.n_jets(order)
    dout_np = rng.standard_normal((n, J, n_out)).astype(np.float32)
    out, dflat = run(m, torch.tensor(x_np).cuda(), order, lap, dout_np)
    prog, flat = AR.build_program(m)
    ref_out, cache = JR.forward(prog, flat.astype(np.float64), x_np.astype(np.float64), order, lap)
    ref_dp, _ = JR.backward(prog, flat.astype(np.float64), x_np.astype(np.float64), order, cache,
                            dout_np.astype(np.float64), lap)
    assert relerr(out, ref_out) < TOL
    assert relerr(dflat, ref_dp) < TOL
    # per-tensor check of the parameter gradient (weights, biases of every layer)
    off = 0
    for p_ in m.parameters():
        k = p_.numel()
        assert relerr(dflat[off:off + k], ref_dp[off:off + k]) < 2 * TOL, (tuple(p_.shape), off)
        off += k


def test_wide_engine_agrees_with_ffma_engine_on_many_tiles():
    """More tiles than CTAs (persistent loops, accumulator flush groups, slab fold), burgers depth."""
    torch.manual_seed(0)
    n = 148 * 32 * 5 + 19
    x = torch.rand((n, 2), device="cuda")
    res = {}
    for eng in (_lib.ENGINE_TCGEN05, _lib.ENGINE_FFMA):
        torch.manual_seed(1)
        m = mp.FNN(2, 1, [128, 128, 128, 128]).cuda()
        m.jet_engine.kernel_engine = eng
        g = torch.Generator(device="cuda").manual_seed(5)
        out, stash = m.jet_engine.forward(x, 3, need_stash=True, lap=(1.0, 0.0))
        dout = torch.randn(out.shape, device="cuda", generator=g)
        res[eng] = (out.cpu().numpy(), m.jet_engine.backward(x, 3, dout, stash, lap=(1.0, 0.0)).cpu().numpy())
    a, b = res[_lib.ENGINE_TCGEN05], res[_lib.ENGINE_FFMA]
    assert relerr(a[0], b[0].astype(np.float64)) < TOL
    assert relerr(a[1], b[1].astype(np.float64)) < TOL


def test_wide_engine_limits():
    m = mp.FNN(2, 1, [128, 128, 128]).cuda()
    e = m.jet_engine
    assert e.tc_supported(3) and e.tc_supported(1) and e.tc_supported(0)
    assert not e.tc_supported(2)          # five channels: FP32-FMA engine
    e.kernel_engine = _lib.ENGINE_TCGEN05
    with pytest.raises(RuntimeError):
        e.forward(torch.rand((64, 2), device="cuda"), 2, need_stash=True)
    # value-only inference without a stash falls to the FP32-FMA engine under AUTO
    e.kernel_engine = _lib.ENGINE_AUTO
    out, stash = e.forward(torch.rand((64, 2), device="cuda"), 0, need_stash=False)
    assert out.shape == (1, 64, 1)
