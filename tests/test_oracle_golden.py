"""Pin the oracle (oracle/jet_ref.py, oracle/autograd_ref.py) against vectors produced by the
unmodified reference (tests/golden/*.npz, written by tests/golden/make_golden.py), and against
the reference's own activation known-answer tests
(reference tests/neural_network/test_activation_function.py:8-59)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import autograd_ref as AR
from oracle import jet_ref as JR
from tests import problems

GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ["regression_1D", "poisson_2D_1C", "burgers_1D", "allen_cahn_2D_1C",
         "navier_stokes_2D", "navier_stokes_3D"]
GOLD_KW = {  # depth used by make_golden.CASES
    "regression_1D": dict(hidden_layers=[64] * 3),
    "poisson_2D_1C": dict(hidden_layers=[64] * 9),
    "burgers_1D": dict(hidden_layers=[128] * 2),
    "allen_cahn_2D_1C": dict(hidden_layers=[64] * 3),
    "navier_stokes_2D": dict(hidden_layers=[64] * 3),
    "navier_stokes_3D": dict(hidden_layers=[64] * 3, encoding_dim=64, sigma=1.0),
}


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def oracle_model(name, g, dtype=torch.float32):
    _, kind, _ = problems.CONFIGS[name]
    m = AR.MODELS[kind](int(g["input_dim"]), int(g["output_dim"]), **GOLD_KW[name]).to(dtype)
    sd = {k: torch.tensor(g["w::" + k]).to(dtype) for k in g["state_keys"]}
    m.load_state_dict(sd)
    return m


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_activation_kats():
    k = np.load(os.path.join(GOLD, "activation_kats.npz"))
    x = k["x"].astype(np.float64)
    assert np.allclose(JR.act_derivs(JR.ACT_GELU, x, 0)[0], k["gelu"], atol=1e-6)
    assert np.allclose(JR.act_derivs(JR.ACT_SINE, x, 0)[0], k["sin"], atol=1e-6)
    one = np.array([1.0, -1.0, 0.0])
    got = JR.act_derivs(JR.ACT_GELU, one, 0)[0]
    # the reference's own KATs, atol 1e-5 as in its test file
    assert np.allclose(got, [0.84134474, -0.15865526, 0.0], atol=1e-5)
    assert np.allclose(AR.GELU()(torch.tensor([1.0, -1.0, 0.0])).numpy(),
                       [0.84134474, -0.15865526, 0.0], atol=1e-5)
    s = JR.act_derivs(JR.ACT_SINE, np.array([0.0, np.pi / 2, np.pi, 3 * np.pi / 2, -np.pi / 2]), 0)[0]
    assert np.allclose(s, [0, 1, 0, -1, -1], atol=1e-5)


def test_activation_derivatives_fd():
    z = np.linspace(-4, 4, 41)
    for act in (JR.ACT_GELU, JR.ACT_SINE):
        eps = 1e-5
        d = JR.act_derivs(act, z, 3)
        dp = JR.act_derivs(act, z + eps, 3)
        dm = JR.act_derivs(act, z - eps, 3)
        for k in range(3):
            fd = (dp[k] - dm[k]) / (2 * eps)
            assert np.allclose(fd, d[k + 1], atol=1e-8), (act, k)


@pytest.mark.parametrize("name", NAMES)
def test_jet_forward_matches_reference(name):
    """u, grad u, diag Hessian from the fp64 jet restatement == reference fp64 autograd."""
    g = load(name)
    m = oracle_model(name, g, torch.float64)
    prog, flat = AR.build_program(m)
    x = g["pts0"].astype(np.float64)
    d = x.shape[1]
    out, _ = JR.forward(prog, flat, x, 2)
    assert relerr(out[:, 0], g["u_f64"]) < 1e-12
    assert relerr(np.transpose(out[:, 1:1 + d], (0, 2, 1)), g["du_f64"]) < 1e-12
    assert relerr(np.transpose(out[:, 1 + d:], (0, 2, 1)), g["d2u_f64"]) < 1e-11
    # fp32 evaluation of the same algebra stays inside the 1e-5 budget vs fp64 truth
    out32, _ = JR.forward(prog, flat.astype(np.float32), x.astype(np.float32), 2)
    assert relerr(out32[:, 0], g["u_f64"]) < 1e-5
    assert relerr(np.transpose(out32[:, 1 + d:], (0, 2, 1)), g["d2u_f64"]) < 2e-5


@pytest.mark.parametrize("name", NAMES)
def test_autograd_port_matches_reference(name):
    """Residuals, mean loss, component losses and parameter gradient of the torch port ==
    reference (fp64 to 1e-12; fp32 to fp32 round-off)."""
    g = load(name)
    ns = types.SimpleNamespace(grad=AR.grad, unpack=AR.unpack, Symbols=None)
    conds, din, dout = problems.CONFIGS[name][0](ns)
    fns = [c[0] for c in conds]
    for tag, dtype, tol in (("f64", torch.float64, 1e-11), ("f32", torch.float32, 2e-5)):
        m = oracle_model(name, g, dtype)
        pts = [torch.tensor(g[f"pts{i}"], dtype=dtype).requires_grad_() for i in range(len(fns))]
        k = 0
        for ci, (fn, p) in enumerate(zip(fns, pts)):
            for r in fn(m, p):
                assert relerr(r.detach().numpy(), g[f"res{ci}_{k}_{tag}"]) < tol, (ci, k)
                k += 1
        total, losses = AR.mean_loss(m, fns, pts)
        assert relerr(total.detach().numpy(), g[f"total_{tag}"]) < tol
        assert relerr(losses.detach().numpy(), g[f"losses_{tag}"]) < tol
        legacy, _ = AR.legacy_loss(m, fns, pts)
        assert relerr(legacy.detach().numpy(), g[f"legacy_{tag}"]) < tol
        m.zero_grad()
        total.backward()
        flat = np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).numpy().reshape(-1)
                               for p in m.parameters()])
        assert relerr(flat, g[f"grad_{tag}"]) < (1e-10 if tag == "f64" else 5e-4)


class _JetHarness:
    """Test-local stand-in for the product boundary: ``model(arg)`` returns the value slice of a
    jet block computed by oracle/jet_ref.py and ``grad`` decodes derivative channels from it."""

    def __init__(self, prog, flat, d, dout, order=2):
        self.prog, self.flat, self.d, self.dout, self.order = prog, flat, d, dout, order
        self.blocks = []

    def model(self, arg):
        x = arg.detach().numpy()
        out, cache = JR.forward(self.prog, self.flat, x, self.order)
        blk = torch.tensor(out, requires_grad=True)  # [N,J,o] leaf
        self.blocks.append((x, cache, blk))
        return blk[:, 0, :]

    def grad(self, u, arg):
        d = self.d
        for _, _, blk in self.blocks:
            for k in range(self.dout):
                if u.data_ptr() == blk[:, 0, k].data_ptr():
                    return blk[:, 1:1 + d, k]
                for i in range(d):
                    if u.data_ptr() == blk[:, 1 + i, k].data_ptr():
                        cols = [torch.full_like(u, float("nan")) for _ in range(d)]
                        cols[i] = blk[:, 1 + d + i, k]
                        return torch.stack(cols, 1)
        raise AssertionError("unknown tensor passed to grad")

    def param_grad(self):
        dflat = np.zeros_like(self.flat)
        for x, cache, blk in self.blocks:
            if blk.grad is not None:
                dp, _ = JR.backward(self.prog, self.flat, x, self.order, cache, blk.grad.numpy())
                dflat += dp
        return dflat


@pytest.mark.parametrize("name", NAMES)
def test_jet_backward_matches_reference(name):
    """Parameter gradient of the mean loss through the hand-derived jet reverse pass (fp64) ==
    reference ``loss.backward()`` in fp64."""
    g = load(name)
    m = oracle_model(name, g, torch.float64)
    prog, flat = AR.build_program(m)
    h = _JetHarness(prog, flat, int(g["input_dim"]), int(g["output_dim"]))
    ns = types.SimpleNamespace(grad=h.grad, unpack=AR.unpack, Symbols=None)
    conds, _, _ = problems.CONFIGS[name][0](ns)
    total_sq, count = 0.0, 0
    for i, (fn, _) in enumerate(conds):
        pts = torch.tensor(g[f"pts{i}"], dtype=torch.float64)
        for r in fn(h.model, pts):
            total_sq = total_sq + torch.sum(r**2)
            count += len(r)
    total = total_sq / count
    assert relerr(total.detach().numpy(), g["total_f64"]) < 1e-11
    total.backward()
    dflat = h.param_grad()
    if name == "navier_stokes_3D":  # frozen Fourier B: the reference leaves its grad at None/0
        dflat[: 3 * 32] = 0.0
    assert relerr(dflat, g["grad_f64"]) < 1e-10


def test_adam_restatement():
    g = load("poisson_2D_1C")
    p0 = np.concatenate([g["w::" + k].reshape(-1) for k in g["state_keys"]]).astype(np.float32)
    grad = g["grad_f32"].astype(np.float32)
    m = np.zeros_like(p0)
    v = np.zeros_like(p0)
    p = p0.copy()
    JR.adam_step(p, grad, m, v, 1, 1e-3)
    assert np.abs(p - g["params_after1"]).max() < 2e-7
