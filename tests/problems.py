"""The five BASELINE.json workloads as residual callables, written against an injected namespace.

``ns`` supplies ``grad`` and ``unpack`` (and, for burgers, ``Symbols``), so the very same
residual code runs on the oracle (stock autograd helpers) and on the product
(``multipinn_b200.condition``).  The formulas follow the reference's example problems:
  * regression_1D             ``examples/regression_1D/problem.py:17-35``
  * poisson_2D_1C             ``examples/poisson_2D_1C/problem.py:18-70``
  * burgers_1D                ``examples/burgers_1D/problem.py:9-41``
  * allen_cahn_2D_1C          ``examples/allen_cahn_2D_1C/problem.py:12-62``
  * navier_stokes_2D_obstacle ``examples/navier_stokes_2D_obstacle_multigpu/problem.py:10-49``
  * navier_stokes_3D_pipe     ``examples/navier_stokes_3D_pipe/problem.py:12-70``
Geometries are reduced to axis-aligned boxes ``(low, high)`` (point sampling is outside the
hot path); each problem returns ``[(residual_fn, (low, high)), ...], input_dim, output_dim``.
"""
import math

import torch

PI = math.pi


def regression_1D(ns):
    unpack = ns.unpack

    def inner(model, arg):
        (u,) = unpack(model(arg))
        (x,) = unpack(arg)
        arg_2 = arg + torch.ones_like(arg)
        (u_2,) = unpack(model(arg_2))
        return [u * u_2 - x]

    return [(inner, ([0.0], [4.0]))], 1, 1


def poisson_2D_1C(ns):
    grad, unpack = ns.grad, ns.unpack

    def basic(model, arg):
        f = model(arg)
        (u,) = unpack(f)
        x, y = unpack(arg)
        return f, u, x, y

    def inner(model, arg):
        f, u, x, y = basic(model, arg)
        u_x, u_y = unpack(grad(u, arg))
        u_xx, u_xy = unpack(grad(u_x, arg))
        u_yx, u_yy = unpack(grad(u_y, arg))
        return [u_xx + u_yy + 20 * PI**2 * torch.cos(2 * PI * x) * torch.sin(4 * PI * y)]

    def bc1(model, arg):
        f, u, x, y = basic(model, arg)
        return [u - torch.sin(4 * PI * y)]

    def bc2(model, arg):
        f, u, x, y = basic(model, arg)
        return [u + torch.sin(4 * PI * y) - 0.25]

    def bc3(model, arg):
        f, u, x, y = basic(model, arg)
        return [u - 0.5 * x]

    def bc4(model, arg):
        f, u, x, y = basic(model, arg)
        return [u - 0.5 * x]

    conds = [
        (inner, ([0, 0], [0.5, 0.5])),
        (bc1, ([0, 0], [0, 0.5])),
        (bc2, ([0.5, 0], [0.5, 0.5])),
        (bc3, ([0, 0], [0.5, 0])),
        (bc4, ([0, 0.5], [0.5, 0.5])),
    ]
    return conds, 2, 1


def burgers_1D(ns):
    grad, unpack = ns.grad, ns.unpack
    v = 0.01 / PI
    if getattr(ns, "Symbols", None) is not None:
        symbols = ns.Symbols(variables="x, t", functions="u")
        inner_symbols = symbols("u, u_x, u_t, u_xx")
        initial_symbols = symbols("u, x")
    else:  # what Symbols generates (reference condition/symbols.py:125-158)
        def inner_symbols(model, arg):
            (u,) = unpack(model(arg))
            u_x, u_t = unpack(grad(u, arg))
            u_xx, u_xt = unpack(grad(u_x, arg))
            return u, u_x, u_t, u_xx

        def initial_symbols(model, arg):
            (u,) = unpack(model(arg))
            x, t = unpack(arg)
            return u, x

    def inner(model, arg):
        u, u_x, u_t, u_xx = inner_symbols(model, arg)
        return [u_t + u * u_x - v * u_xx]

    def bc(model, arg):
        arg_x = arg.clone()
        u_x_min = model(arg)[:, 0]
        arg_x[:, 0] += 2.0
        u_x_max = model(arg_x)[:, 0]
        return [u_x_min, u_x_max]

    def ic(model, arg):
        u, x = initial_symbols(model, arg)
        return [u + torch.sin(PI * x)]

    conds = [
        (inner, ([-1, 0], [1, 1])),
        (bc, ([-1, 0], [-1, 1])),
        (ic, ([-1, 0], [1, 0])),
    ]
    return conds, 2, 1


def allen_cahn_2D_1C(ns):
    grad, unpack = ns.grad, ns.unpack

    def basic(model, arg):
        f = model(arg)
        (u,) = unpack(f)
        x, t = unpack(arg)
        return f, u, x, t

    def inner(model, arg):
        f, u, x, t = basic(model, arg)
        u_x, u_t = unpack(grad(u, arg))
        u_xx, u_xt = unpack(grad(u_x, arg))
        return [u_t - 1e-4 * u_xx + 5 * (u**3 - u)]

    def bc(model, arg):
        f, u, x, t = basic(model, arg)
        u_x, u_t = unpack(grad(u, arg))
        return [u_x, u + 1]

    def ic(model, arg):
        f, u, x, t = basic(model, arg)
        return [x**2 * torch.cos(PI * x) - u]

    conds = [
        (inner, ([-1, 0], [1, 1])),
        (bc, ([-1, 0], [-1, 1])),
        (bc, ([1, 0], [1, 1])),
        (ic, ([-1, 0], [1, 0])),
    ]
    return conds, 2, 1


def navier_stokes_2D(ns, re=50):
    grad, unpack = ns.grad, ns.unpack
    i_re = 1 / re

    def basic(model, arg):
        f = model(arg)
        u, v, p = unpack(f)
        x, y = unpack(arg)
        return f, u, v, p, x, y

    def inner(model, arg):
        f, u, v, p, x, y = basic(model, arg)
        u_x, u_y = unpack(grad(u, arg))
        v_x, v_y = unpack(grad(v, arg))
        p_x, p_y = unpack(grad(p, arg))
        u_xx, _ = unpack(grad(u_x, arg))
        v_xx, _ = unpack(grad(v_x, arg))
        _, u_yy = unpack(grad(u_y, arg))
        _, v_yy = unpack(grad(v_y, arg))
        eq1 = u * u_x + v * u_y + p_x - i_re * (u_xx + u_yy)
        eq2 = u * v_x + v * v_y + p_y - i_re * (v_xx + v_yy)
        eq3 = u_x + v_y
        return [eq1, eq2, eq3]

    def bc_x_min(model, arg):
        f, u, v, p, x, y = basic(model, arg)
        return [u - (-0.16 * y**2 + 1), v, p]

    def bc_x_max(model, arg):
        f, u, v, p, x, y = basic(model, arg)
        return [p]

    def block(model, arg):
        f, u, v, p, x, y = basic(model, arg)
        return [u, v]

    conds = [
        (inner, ([-5, -2.5], [5, 2.5])),
        (bc_x_min, ([-5, -2.5], [-5, 2.5])),
        (bc_x_max, ([5, -2.5], [5, 2.5])),
        (block, ([-0.5, -0.5], [0.5, -0.5])),
    ]
    return conds, 2, 3


def navier_stokes_3D(ns, re=60):
    grad, unpack = ns.grad, ns.unpack
    iRe = 1 / re

    def basic(model, arg):
        f = model(arg)
        u, v, w, p = unpack(f)
        x, y, z = unpack(arg)
        return f, u, v, w, p, x, y, z

    def inner(model, arg):
        f, u, v, w, p, x, y, z = basic(model, arg)
        u_x, u_y, u_z = unpack(grad(u, arg))
        v_x, v_y, v_z = unpack(grad(v, arg))
        w_x, w_y, w_z = unpack(grad(w, arg))
        p_x, p_y, p_z = unpack(grad(p, arg))
        u_xx, _, _ = unpack(grad(u_x, arg))
        v_xx, _, _ = unpack(grad(v_x, arg))
        w_xx, _, _ = unpack(grad(w_x, arg))
        _, u_yy, _ = unpack(grad(u_y, arg))
        _, v_yy, _ = unpack(grad(v_y, arg))
        _, w_yy, _ = unpack(grad(w_y, arg))
        _, _, u_zz = unpack(grad(u_z, arg))
        _, _, v_zz = unpack(grad(v_z, arg))
        _, _, w_zz = unpack(grad(w_z, arg))
        eq1 = u * u_x + v * u_y + w * u_z + p_x - iRe * (u_xx + u_yy + u_zz)
        eq2 = u * v_x + v * v_y + w * v_z + p_y - iRe * (v_xx + v_yy + v_zz)
        eq3 = u * w_x + v * w_y + w * w_z + p_z - iRe * (w_xx + w_yy + w_zz)
        eq4 = u_x + v_y + w_z
        return [eq1, eq2, eq3, eq4]

    def input_cond(model, arg):
        f, u, v, w, p, x, y, z = basic(model, arg)
        u_f = 2.0 / (0.25 * 0.25) * (0.25 - (z - 0.5) ** 2) * (0.25 - (y - 0.5) ** 2) / 0.89031168
        u_f = torch.nn.functional.relu(u_f)
        return [(u - u_f), v, w]

    def output_cond(model, arg):
        f, u, v, w, p, x, y, z = basic(model, arg)
        return [p]

    def walls(model, arg):
        f, u, v, w, p, x, y, z = basic(model, arg)
        return [u, v, w]

    conds = [
        (inner, ([0, 0, 0], [9, 3, 1])),
        (input_cond, ([0, 0, 0], [0, 1, 1])),
        (output_cond, ([9, 2, 0], [9, 3, 1])),
        (walls, ([0, 0, 0], [4, 0, 1])),
    ]
    return conds, 3, 4


# name -> (problem fn, model kind, model kwargs)   
CONFIGS = {
    "regression_1D": (regression_1D, "FNN", dict(hidden_layers=[64] * 9)),
    "poisson_2D_1C": (poisson_2D_1C, "FNN", dict(hidden_layers=[64] * 9)),
    "burgers_1D": (burgers_1D, "FNN", dict(hidden_layers=[128] * 4)),
    "allen_cahn_2D_1C": (allen_cahn_2D_1C, "FNN", dict(hidden_layers=[64] * 9)),
    "navier_stokes_2D": (navier_stokes_2D, "ResNet", dict(hidden_layers=[64] * 9)),
    "navier_stokes_3D": (navier_stokes_3D, "FFNN",
                         dict(hidden_layers=[64] * 9, encoding_dim=64, sigma=1.0)),
}


def sample_points(low, high, n, rng):
    """``Hypercube.random_points``: low + side * U[0,1) in fp32 (reference geometry/domain_nd.py:116-118)."""
    import numpy as np

    low = np.asarray(low, dtype="float32")
    high = np.asarray(high, dtype="float32")
    u = rng.random(size=(n, len(low))).astype("float32")
    return (high - low) * u + low
