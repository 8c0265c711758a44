"""ORACLE (test infrastructure, not product code) -- CPU restatement of the jet algebra.

This file restates, in plain numpy, the arithmetic the multipinn hot path performs with
nested ``torch.autograd.grad`` calls, as a forward Taylor-jet propagation plus a
hand-derived reverse pass.  It is the executable specification of the CUDA kernels in
``multipinn_b200/csrc`` and the "expected" side of the parity tests.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may
import it.  Nothing under ``multipinn_b200/`` does.

Pinning: the reference's own tests hold only the activation known-answer vectors
(reference ``tests/neural_network/test_activation_function.py:8-59``); everything else is
pinned against outputs of the unmodified reference generated in the build container by
``tests/golden/make_golden.py`` (fixtures committed under ``tests/golden/``).

Reference lines restated here
  * ``multipinn/neural_network/activation_function.py:14`` -- gelu = x*0.5*(1+erf(x/1.41421))
  * ``multipinn/neural_network/activation_function.py:26`` -- sine
  * ``multipinn/neural_network/feedforward_neural_network.py:28-52`` -- FNN layer chain
  * ``multipinn/neural_network/light_residual_block.py:23-38``      -- act(W2 act(W1 x+b1)+b2+x)
  * ``multipinn/neural_network/residual_neural_network.py:77-126``   -- ResNet chain
  * ``multipinn/neural_network/fourier_features.py:37-40``           -- [sin(xB), cos(xB)]
  * ``multipinn/condition/diff.py:47``  -- grad(u,arg) == per-point d u / d x (rows independent)

Jet channels per point: 0 = value, 1..d = d/dx_i, d+1..2d = d2/dx_i2.
  order 0 -> J = 1, order 1 -> J = 1+d, order 2 -> J = 1+2d.
  order 3 ("Laplace mode", include/mpinn_b200.h MPINN_ORDER_LAPLACE) -> J = d+2: value, d first derivatives and
  the single channel L = sum_i c_i d2/dx_i2 (weights ``lap``).  L is what the reference's examples build from
  separate nested grad() calls (e.g. examples/poisson_2D_1C/problem.py:26-29 ``u_xx + u_yy``); because the chain
  rule for it closes on itself, a_L = s''(z) sum_i c_i z_i^2 + s'(z) z_L, it can be propagated directly.  The
  test-suite checks it against sum_i c_i of the order-2 channels, which are pinned to the reference goldens.
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy is present in the image; erf is the only thing taken from it
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

# op kinds / activations / flags -- MUST match include/mpinn_b200.h
OP_LINEAR = 0
OP_FOURIER = 1
ACT_NONE = 0
ACT_SINE = 1
ACT_GELU = 2
FLAG_SAVE_SKIP = 1
FLAG_ADD_SKIP = 2

GELU_C = 1.0 / 1.41421  # the reference divides by the literal 1.41421, not sqrt(2)


ORDER_LAPLACE = 3
ORDER_HESSIAN = 4   # value, d first derivatives, d diagonal second derivatives, d(d-1)/2 mixed ones: pairs (i<j) in
                    # the order (0,1), (0,2), (1,2) -- what `grad(u_x, arg)` returns in ALL its columns in the reference
                    # (condition/diff.py:47), used e.g. by examples/surfaces_elliptic/problem.py:15-25


def mixed_pairs(d: int):
    return [(i, j) for i in range(d) for j in range(i + 1, d)]


def n_jets(d: int, order: int) -> int:
    if order == ORDER_LAPLACE:
        return d + 2
    if order == ORDER_HESSIAN:
        return 1 + 2 * d + d * (d - 1) // 2
    return 1 + order * d


def _lap(lap, d, dtype):
    return np.ones(d, dtype=dtype) if lap is None else np.asarray(lap, dtype=dtype)


def act_derivs(act: int, z: np.ndarray, n: int):
    """Return [sigma, sigma', sigma'', sigma'''][:n+1] evaluated at z."""
    if act == ACT_NONE:
        out = [z, np.ones_like(z), np.zeros_like(z), np.zeros_like(z)]
    elif act == ACT_SINE:
        s, c = np.sin(z), np.cos(z)
        out = [s, c, -s, -c]
    elif act == ACT_GELU:
        c = z.dtype.type(GELU_C)
        Phi = 0.5 * (1.0 + _erf(c * z)).astype(z.dtype)
        phi = (c / z.dtype.type(math.sqrt(math.pi))) * np.exp(-(c * z) ** 2)
        c2z2 = (c * z) ** 2
        out = [
            z * Phi,
            Phi + z * phi,
            phi * (2.0 - 2.0 * c2z2),
            -2.0 * c * c * z * phi * (4.0 - 2.0 * c2z2),
        ]
    else:
        raise ValueError(act)
    return out[: n + 1]


def _act_jets(act, z, d, order, lap=None):
    """z: [N,J,W] pre-activation jets -> activation jets [N,J,W]."""
    if act == ACT_NONE:
        return z.copy()
    a = np.empty_like(z)
    ders = act_derivs(act, z[:, 0], min(order, 2))
    a[:, 0] = ders[0]
    if order == ORDER_LAPLACE:
        c = _lap(lap, d, z.dtype)
        q = sum(c[i] * z[:, 1 + i] ** 2 for i in range(d))
        for i in range(d):
            a[:, 1 + i] = ders[1] * z[:, 1 + i]
        a[:, 1 + d] = ders[2] * q + ders[1] * z[:, 1 + d]
        return a
    if order >= 1:
        for i in range(d):
            a[:, 1 + i] = ders[1] * z[:, 1 + i]
    if order in (2, ORDER_HESSIAN):
        for i in range(d):
            a[:, 1 + d + i] = ders[2] * z[:, 1 + i] ** 2 + ders[1] * z[:, 1 + d + i]
    if order == ORDER_HESSIAN:
        for m, (i, j) in enumerate(mixed_pairs(d)):
            a[:, 1 + 2 * d + m] = ders[2] * z[:, 1 + i] * z[:, 1 + j] + ders[1] * z[:, 1 + 2 * d + m]
    return a


def _act_jets_bwd(act, z, abar, d, order, lap=None):
    """Reverse of _act_jets: given abar [N,J,W] return zbar [N,J,W]."""
    if act == ACT_NONE:
        return abar.copy()
    ders = act_derivs(act, z[:, 0], min(order, 2) + 1)
    zb = np.empty_like(abar)
    zb[:, 0] = abar[:, 0] * ders[1]
    if order == ORDER_LAPLACE:
        c = _lap(lap, d, z.dtype)
        aL, zL = abar[:, 1 + d], z[:, 1 + d]
        zb[:, 1 + d] = aL * ders[1]
        zb[:, 0] += aL * ders[2] * zL
        for i in range(d):
            zi = z[:, 1 + i]
            zb[:, 1 + i] = abar[:, 1 + i] * ders[1] + 2.0 * c[i] * aL * ders[2] * zi
            zb[:, 0] += abar[:, 1 + i] * ders[2] * zi + aL * c[i] * ders[3] * zi * zi
        return zb
    if order >= 1:
        for i in range(d):
            zi = z[:, 1 + i]
            zb[:, 1 + i] = abar[:, 1 + i] * ders[1]
            zb[:, 0] += abar[:, 1 + i] * ders[2] * zi
    if order == ORDER_HESSIAN:
        for m, (i, j) in enumerate(mixed_pairs(d)):
            zi, zj, zij, aij = z[:, 1 + i], z[:, 1 + j], z[:, 1 + 2 * d + m], abar[:, 1 + 2 * d + m]
            zb[:, 1 + 2 * d + m] = aij * ders[1]
            zb[:, 1 + i] += aij * ders[2] * zj
            zb[:, 1 + j] += aij * ders[2] * zi
            zb[:, 0] += aij * (ders[3] * zi * zj + ders[2] * zij)
    if order in (2, ORDER_HESSIAN):
        for i in range(d):
            zi = z[:, 1 + i]
            zii = z[:, 1 + d + i]
            aii = abar[:, 1 + d + i]
            zb[:, 1 + d + i] = aii * ders[1]
            zb[:, 1 + i] += 2.0 * aii * ders[2] * zi
            zb[:, 0] += aii * (ders[3] * zi * zi + ders[2] * zii)
    return zb


def input_jets(x: np.ndarray, order: int) -> np.ndarray:
    N, d = x.shape
    J = n_jets(d, order)
    h = np.zeros((N, J, d), dtype=x.dtype)
    h[:, 0] = x
    if order >= 1:
        for i in range(d):
            h[:, 1 + i, i] = 1.0
    return h


def _cos_jets(z, d, order, lap=None):
    """Jets of cos(z): cos' = -sin, cos'' = -cos."""
    c = np.empty_like(z)
    cz, sz = np.cos(z[:, 0]), np.sin(z[:, 0])
    c[:, 0] = cz
    if order >= 1:
        for i in range(d):
            c[:, 1 + i] = -sz * z[:, 1 + i]
    if order in (2, ORDER_HESSIAN):
        for i in range(d):
            c[:, 1 + d + i] = -cz * z[:, 1 + i] ** 2 - sz * z[:, 1 + d + i]
    if order == ORDER_HESSIAN:
        for m, (i, j) in enumerate(mixed_pairs(d)):
            c[:, 1 + 2 * d + m] = -cz * z[:, 1 + i] * z[:, 1 + j] - sz * z[:, 1 + 2 * d + m]
    if order == ORDER_LAPLACE:
        w = _lap(lap, d, z.dtype)
        q = sum(w[i] * z[:, 1 + i] ** 2 for i in range(d))
        c[:, 1 + d] = -cz * q - sz * z[:, 1 + d]
    return c


def forward(program, params: np.ndarray, x: np.ndarray, order: int, lap=None):
    """Run the layer program.  Returns (out jets [N,J,o], cache for backward).

    program: int array [n_ops, 8] = kind, in_dim, out_dim, w_off, b_off, act, flags, 0
    params : flat parameter vector (W stored [out,in] row-major as nn.Linear does)
    """
    program = np.asarray(program, dtype=np.int64).reshape(-1, 8)
    N, d = x.shape
    h = input_jets(x, order)
    cache = []
    skip = None
    for kind, din, dout, w_off, b_off, act, flags, _ in program:
        assert h.shape[2] == din, (h.shape, din)
        if kind == OP_FOURIER:
            m = dout // 2
            B = params[w_off : w_off + din * m].reshape(din, m)
            z = h @ B  # [N,J,m]  (h is the raw input jets here)
            s = _act_jets(ACT_SINE, z, d, order, lap)
            c = _cos_jets(z, d, order, lap)
            cache.append(("fourier", h, z))
            h = np.concatenate([s, c], axis=2)
            continue
        W = params[w_off : w_off + din * dout].reshape(dout, din)
        b = params[b_off : b_off + dout]
        if flags & FLAG_SAVE_SKIP:
            skip = h
        z = h @ W.T
        z[:, 0] += b
        if flags & FLAG_ADD_SKIP:
            z = z + skip
        cache.append(("linear", h, z))
        h = _act_jets(act, z, d, order, lap)
    return h, cache


def backward(program, params: np.ndarray, x: np.ndarray, order: int, cache, dout: np.ndarray, lap=None):
    """Reverse pass.  dout: [N,J,o].  Returns (flat dparams, dx [N,d])."""
    program = np.asarray(program, dtype=np.int64).reshape(-1, 8)
    N, d = x.shape
    dparams = np.zeros_like(params)
    hbar = dout.astype(params.dtype, copy=True)
    skipbar = None
    for (kind, din, dout_, w_off, b_off, act, flags, _), (tag, h, z) in zip(
        program[::-1], cache[::-1]
    ):
        if kind == OP_FOURIER:
            m = dout_ // 2
            B = params[w_off : w_off + din * m].reshape(din, m)
            sbar, cbar = hbar[:, :, :m], hbar[:, :, m:]
            zb = _act_jets_bwd(ACT_SINE, z, sbar, d, order, lap)
            # cosine branch
            cz, sz = np.cos(z[:, 0]), np.sin(z[:, 0])
            zb[:, 0] += cbar[:, 0] * (-sz)
            if order >= 1:
                for i in range(d):
                    zi = z[:, 1 + i]
                    zb[:, 1 + i] += cbar[:, 1 + i] * (-sz)
                    zb[:, 0] += cbar[:, 1 + i] * (-cz) * zi
            if order == ORDER_HESSIAN:
                for q, (i, j) in enumerate(mixed_pairs(d)):
                    zi, zj, zij, cij = z[:, 1 + i], z[:, 1 + j], z[:, 1 + 2 * d + q], cbar[:, 1 + 2 * d + q]
                    zb[:, 1 + 2 * d + q] += cij * (-sz)
                    zb[:, 1 + i] += cij * (-cz) * zj
                    zb[:, 1 + j] += cij * (-cz) * zi
                    zb[:, 0] += cij * (sz * zi * zj - cz * zij)
            if order in (2, ORDER_HESSIAN):
                for i in range(d):
                    zi, zii = z[:, 1 + i], z[:, 1 + d + i]
                    cii = cbar[:, 1 + d + i]
                    zb[:, 1 + d + i] += cii * (-sz)
                    zb[:, 1 + i] += 2.0 * cii * (-cz) * zi
                    zb[:, 0] += cii * (sz * zi * zi - cz * zii)
            if order == ORDER_LAPLACE:
                w = _lap(lap, d, z.dtype)
                cL, zL = cbar[:, 1 + d], z[:, 1 + d]
                zb[:, 1 + d] += cL * (-sz)
                zb[:, 0] += cL * (-cz) * zL
                for i in range(d):
                    zi = z[:, 1 + i]
                    zb[:, 1 + i] += 2.0 * w[i] * cL * (-cz) * zi
                    zb[:, 0] += cL * w[i] * sz * zi * zi
            dB = np.einsum("njk,njm->km", h, zb)
            dparams[w_off : w_off + din * m] += dB.reshape(-1)
            hbar = zb @ B.T
            continue
        W = params[w_off : w_off + din * dout_].reshape(dout_, din)
        zb = _act_jets_bwd(act, z, hbar, d, order, lap)
        if flags & FLAG_ADD_SKIP:
            skipbar = zb
        dW = np.einsum("njo,nji->oi", zb, h)
        dparams[w_off : w_off + din * dout_] += dW.reshape(-1)
        dparams[b_off : b_off + dout_] += zb[:, 0].sum(axis=0)
        hbar = zb @ W
        if flags & FLAG_SAVE_SKIP:
            hbar = hbar + skipbar
            skipbar = None
    dx = hbar[:, 0].copy()  # only the value channel of the input jets depends on x
    return dparams, dx


# --------------------------------------------------------------------------- Adam
def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """torch.optim.Adam (no weight decay / amsgrad), restated; ``step`` is the 1-based count.

    Follows torch/optim/adam.py ``_single_tensor_adam``:
      m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
      p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
    """
    g = g * p.dtype.type(grad_scale)
    m[:] = beta1 * m + (1 - beta1) * g
    v[:] = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1**step
    bc2 = 1 - beta2**step
    step_size = lr / bc1
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p[:] = p - p.dtype.type(step_size) * (m / denom)
    return p, m, v
