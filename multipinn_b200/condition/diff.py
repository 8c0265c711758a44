"""Derivative helpers with the reference's names and signatures
(reference ``multipinn/condition/diff.py``).

``grad`` does not call ``torch.autograd.grad``: the derivatives were already produced, in the same
fused kernel launch as the network output, as extra channels of the jet block that ``u`` is a view
of (see ``multipinn_b200/jet.py``).  The finite-difference helpers are value-only model calls and
need no jets.
"""
import torch

from ..jet import UnsupportedDerivative, jet_grad, unpack  # noqa: F401  (re-exported)


def grad(u: torch.Tensor, arg: torch.Tensor) -> torch.Tensor:
    """d u / d arg per point, shape ``[N, dim]`` (reference diff.py:25-47).

    ``u`` must be a column of a network output (``unpack(model(arg))`` or ``model(arg)[:, k]``) or a
    column of a previous ``grad`` result; ``arg`` must be the tensor the model was called with.
    ``grad`` of a first derivative returns the matching *diagonal* second derivative in its own
    column; the mixed-derivative columns are NaN (unsupported by the jet path, rejected if used).
    """
    return jet_grad(u, arg)


def _eps_column(eps, like: torch.Tensor) -> torch.Tensor:
    return torch.as_tensor(eps, dtype=like.dtype, device=like.device).reshape(-1, 1)


def num_diff(model, arg, fns, direction, eps=2**-13):
    """Forward-difference directional derivative (reference diff.py:50-81); ``fns == model(arg)``."""
    h = _eps_column(eps, arg)
    shifted = arg + (direction.to(arg) * h).detach()
    return (model(shifted) - fns) / h


def num_diff_random(model, arg, fns, direction, max_eps=1e-2, min_eps=1e-4):
    """``num_diff`` with a random signed step per point (reference diff.py:84-111)."""
    n = len(arg)
    sign = 1 - 2 * torch.randint(0, 2, (n,), device=arg.device)
    size = torch.empty(n, device=arg.device, dtype=arg.dtype).uniform_(min_eps, max_eps)
    return num_diff(model, arg, fns, direction, sign * size)


def _second_difference(f_pp, f_pm, f_mp, f_mm, denom):
    return (f_pp - f_pm - f_mp + f_mm) / denom


def num_diff_second_same(model, arg, fns, direction, eps=2**-7):
    """Central second difference along one direction (reference diff.py:114-139)."""
    h = _eps_column(eps, arg)
    step = (direction.to(arg) * h).detach()
    return _second_difference(model(arg + step), fns, fns, model(arg - step), h * h)


def num_diff_second_cross(model, arg, direction1, direction2, eps1=2**-7, eps2=2**-7):
    """Four-point mixed second difference (reference diff.py:142-175)."""
    h1, h2 = _eps_column(eps1, arg), _eps_column(eps2, arg)
    s1 = (direction1.to(arg) * h1).detach()
    s2 = (direction2.to(arg) * h2).detach()
    return _second_difference(model(arg + s1 + s2), model(arg + s1 - s2), model(arg - s1 + s2),
                              model(arg - s1 - s2), h1 * h2)


def num_laplace(model, arg, fns, eps=2**-7):
    """Laplacian by averaging second differences over a random orthonormal frame (reference diff.py:178-205)."""
    dim = arg.shape[1]
    frame, _ = torch.linalg.qr(torch.randn((dim, dim), device=arg.device, dtype=arg.dtype))
    total = torch.zeros_like(fns)
    for i in range(dim):
        step = (frame[i:i + 1, :] * eps).detach()
        total = total + (model(arg - step) - 2 * fns + model(arg + step))
    return total / (eps * eps)


def _diff_residual(model, arg, eps=1e-4):
    """Self-check used by the reference's tests (``condition/diff.py:208-245``): analytic ``grad`` minus a
    central difference, one ``[N]`` tensor per (output, input) pair, outputs major."""
    f = model(arg)
    n_out, n_in = f.shape[1], arg.shape[1]
    analytic = [grad(f[:, o], arg) for o in range(n_out)]
    central = []
    for i in range(n_in):
        h = torch.zeros(1, n_in, dtype=arg.dtype, device=arg.device)
        h[0, i] = eps
        central.append((model(arg + h) - model(arg - h)) / (2 * eps))
    return [analytic[o][:, i] - central[i][:, o] for o in range(n_out) for i in range(n_in)]
