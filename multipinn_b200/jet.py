"""The drop-in boundary between user residual code and the jet kernels.

``model(arg)`` runs ONE fused forward that produces, per point, the network outputs together with
their first derivatives and second (diagonal) derivatives -- a *jet block* ``[J, N, o]`` -- and
hands back the value slice ``[N, o]``.  ``grad(u, arg)`` (same signature as the reference's
``multipinn/condition/diff.py:25-47``) does not differentiate anything: it recognises ``u`` as a
view of a live jet block (by storage pointer / offset / stride, so ``unpack`` and ``f[:, k]``
indexing both work) and returns the matching derivative channels as views of the same block.
The block is the output of one ``torch.autograd.Function`` whose backward is the fused reverse
kernel, so ``loss.backward()`` (reference ``trainer/trainer.py:141``) reaches the parameters in
one launch per model call.

How many derivative orders a call needs is discovered by running the residual function once
(the reference does the same probe in ``Condition.init_output_len``, ``condition/condition.py:102-114``)
and remembered per (condition, model-call index) in a ``ResidualPlan``.  A call that turns out to
need more than planned is transparently re-run at the higher order.

Mixed second derivatives: a block of order <= 2 carries only the diagonal ones, and the off-diagonal columns of
``grad(u_x, arg)`` come back NaN-poisoned.  When the probe finds a NaN residual (i.e. a mixed derivative is
actually *used*) the condition's plan is raised to the full-Hessian mode (order 4: d(d-1)/2 extra channels,
FP32-FMA engine) and those columns become real.

Not supported (raises, never silently wrong): derivatives of order >= 3, ``grad`` of anything that is not a direct
view of a model output.
"""
from __future__ import annotations

import threading
import weakref
from typing import Dict, List, Optional

import torch
from torch.autograd.function import once_differentiable

from . import trace as _trace


class UnsupportedDerivative(NotImplementedError):
    pass


def _lap_kw(lap):
    return {"lap": lap} if lap is not None else {}


# ----------------------------------------------------------------------------- residual plans
class ResidualPlan:
    """Derivative order needed by the k-th ``model(arg)`` call inside one residual function."""

    def __init__(self):
        self.orders: List[int] = []
        self.cursor = 0

    def begin(self):
        self.cursor = 0

    def next_order(self) -> int:
        k = self.cursor
        self.cursor += 1
        if k >= len(self.orders):
            self.orders.append(0)
        return k

    def raise_order(self, k: int, order: int):
        if order > self.orders[k]:
            self.orders[k] = order


_tls = threading.local()


def current_plan() -> Optional[ResidualPlan]:
    return getattr(_tls, "plan", None)


class use_plan:
    """Context manager: model calls inside take their jet order from (and report back to) ``plan``."""

    def __init__(self, plan: Optional[ResidualPlan]):
        self.plan = plan

    def __enter__(self):
        self.prev = current_plan()
        _tls.plan = self.plan
        if self.plan is not None:
            self.plan.begin()
        return self.plan

    def __exit__(self, *exc):
        _tls.plan = self.prev
        return False


class jet_order:
    """Manual hint for bare residual calls outside a ``Condition``: ``with jet_order(2): fn(model, x)``."""

    def __init__(self, order: int):
        self.order = int(order)

    def __enter__(self):
        self.prev = getattr(_tls, "default_order", 0)
        _tls.default_order = self.order
        return self

    def __exit__(self, *exc):
        _tls.default_order = self.prev
        return False


# ----------------------------------------------------------------------------- jet blocks
class JetBlock:
    __slots__ = ("out_ref", "arg_ref", "arg_ptr", "shift_of", "n", "d", "o", "order", "engine", "plan", "plan_idx",
                 "offset0", "generation", "grad_mode", "__weakref__")

    def __init__(self, out, arg, order, engine, plan, plan_idx):
        self.out_ref = weakref.ref(out)
        self.arg_ref = weakref.ref(arg)
        self.arg_ptr = arg.data_ptr()
        self.shift_of = _shift_origin(arg)
        if self.shift_of is not None:
            out._jet_points = arg        # a temporary (`x + c`): lives as long as the block's output does
        self.n, self.d = arg.shape
        self.o = out.shape[2]
        self.order = order
        self.engine = engine
        self.plan = plan
        self.plan_idx = plan_idx
        self.offset0 = out.storage_offset()
        self.generation = engine.generation          # parameter version the block was computed with
        self.grad_mode = torch.is_grad_enabled()


def _shift_origin(arg: torch.Tensor):
    """``model(x + c)`` (reference examples/portal_1D_1C/problem.py:36-39 then takes ``grad(u, x)``): when the tensor
    the model is called with is a leaf plus / minus a constant, the data pointer of that leaf -- a derivative with
    respect to it equals the derivative with respect to the shifted points (identity Jacobian)."""
    fn = arg.grad_fn
    if fn is None or type(fn).__name__ not in ("AddBackward0", "SubBackward0"):
        return None
    nxt = fn.next_functions
    if len(nxt) != 2:
        return None
    if type(fn).__name__ == "AddBackward0" and getattr(fn, "_saved_alpha", 1) != 1:
        return None
    (f0, _), (f1, _) = nxt
    leaf_fn = None
    if f0 is not None and f1 is None:
        leaf_fn = f0
    elif f1 is not None and f0 is None and type(fn).__name__ == "AddBackward0":
        leaf_fn = f1
    var = getattr(leaf_fn, "variable", None)          # AccumulateGrad
    if var is None or tuple(var.shape) != tuple(arg.shape):
        return None
    return var.data_ptr()


_blocks: Dict[int, JetBlock] = {}


def _register(block: JetBlock, out: torch.Tensor):
    if len(_blocks) > 256:  # prune entries whose block died
        for k in [k for k, b in _blocks.items() if b.out_ref() is None]:
            del _blocks[k]
    _blocks[out.untyped_storage().data_ptr()] = block


def clear_blocks():
    _blocks.clear()


def _lookup(t: torch.Tensor) -> Optional[JetBlock]:
    try:
        key = t.untyped_storage().data_ptr()
    except Exception:
        return None
    b = _blocks.get(key)
    if b is None:
        return None
    if b.out_ref() is None:
        del _blocks[key]
        return None
    return b


class _JetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, x, engine, order):
        out, stash = engine.forward(x, order, need_stash=True)
        ctx.engine = engine
        ctx.order = order
        ctx.save_for_backward(x, stash)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, stash = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            dflat, dx = ctx.engine.backward(x, ctx.order, dout.contiguous(), stash, want_dx=True)
            return dflat, dx, None, None
        dflat = ctx.engine.backward(x, ctx.order, dout.contiguous(), stash)
        return dflat, None, None, None


def _run_block(engine, arg: torch.Tensor, order: int, plan, plan_idx) -> torch.Tensor:
    if arg.dim() != 2 or arg.shape[1] != engine.d:
        raise RuntimeError(f"expected points of shape [N, {engine.d}], got {tuple(arg.shape)}")
    if arg.dtype != torch.float32:
        raise TypeError("multipinn_b200 evaluates fp32 points only")
    engine.ensure_flat()
    x = arg.detach()
    if x.device != engine.flat.device:
        raise RuntimeError(f"points are on {x.device} but the model is on {engine.flat.device}")
    x = x.contiguous()
    if torch.is_grad_enabled() and engine.flat.requires_grad:
        if engine.points_grad and arg.requires_grad:
            x = arg.contiguous()   # stays in the autograd graph: the reverse pass also returns dL/d(points)
        out = _JetFunction.apply(engine.flat, x, engine, order)
    else:
        out, _ = engine.forward(x, order, need_stash=False)
    _register(JetBlock(out, arg, order, engine, plan, plan_idx), out)
    return out


def model_forward(engine, arg: torch.Tensor) -> torch.Tensor:
    """What ``nn.Module.forward`` of the mirrored network classes calls."""
    if isinstance(arg, _trace.TMat):       # symbolic evaluation of a residual (trace.py)
        return _trace.traced_model_call(engine, arg)
    plan = current_plan()
    if plan is not None:
        idx = plan.next_order()
        order = plan.orders[idx]
    else:
        idx = -1
        order = getattr(_tls, "default_order", 0)
    out = _run_block(engine, arg, order, plan, idx)
    return out[0]


# ----------------------------------------------------------------------------- grad / unpack
_nan_cache: Dict[tuple, torch.Tensor] = {}


def _nan_column(n: int, like: torch.Tensor) -> torch.Tensor:
    key = (n, like.device, like.dtype)
    t = _nan_cache.get(key)
    if t is None:
        if len(_nan_cache) > 16:
            _nan_cache.clear()
        t = torch.full((n,), float("nan"), device=like.device, dtype=like.dtype)
        _nan_cache[key] = t
    return t


def _decode(u: torch.Tensor, b: JetBlock):
    """(jet channel, output index) of the block column that ``u`` is a view of."""
    out = b.out_ref()
    N, o = b.n, b.o
    if u.dim() != 1 or u.shape[0] != N or (N > 1 and u.stride(0) != o):
        raise UnsupportedDerivative(
            "grad(u, arg): u must be a full column of a model output (e.g. from unpack(f) or f[:, k])")
    q = u.storage_offset() - b.offset0
    j, rem = divmod(q, N * o)
    n0, k = divmod(rem, o)
    if n0 != 0 or j < 0 or j >= out.shape[0]:
        raise UnsupportedDerivative("grad(u, arg): u is not aligned with a model output column")
    return j, k


def jet_grad(u: torch.Tensor, arg: torch.Tensor) -> torch.Tensor:
    """``grad(u, arg)`` of the reference (condition/diff.py:25-47), served from the jet block."""
    if _trace.is_traced(u) or _trace.is_traced(arg):
        return _trace.traced_grad(u, arg)
    b = _lookup(u)
    if b is None and u.grad_fn is not None and arg.requires_grad:
        # Not an output of a fused jet network: a plain torch module (the families outside the accelerated
        # path keep the reference's stock-autograd behaviour, condition/diff.py:47).  Outputs of a jet network
        # that were transformed first (u*v, ...) are NOT connected to `arg` in autograd and still raise below.
        try:
            return torch.autograd.grad(u, arg, torch.ones_like(u), create_graph=True)[0]
        except RuntimeError:
            pass
    if b is None:
        raise UnsupportedDerivative(
            "grad(u, arg): u is not a view of a model output produced by a multipinn_b200 network. "
            "Derivatives of derived expressions (e.g. grad(u*v, arg)) are not supported by the jet "
            "path; differentiate the outputs and apply the product/chain rule in the residual.")
    if tuple(arg.shape) != (b.n, b.d) or (arg.data_ptr() != b.arg_ptr and arg.data_ptr() != b.shift_of):
        raise UnsupportedDerivative("grad(u, arg): arg is not the tensor the model was called with")
    if arg.data_ptr() != b.arg_ptr:      # model(arg + c): the block's own points
        arg = b.arg_ref()
        if arg is None:
            raise UnsupportedDerivative("grad(u, arg): the shifted points the model was called with are gone")
    j, k = _decode(u, b)
    d = b.d
    if j > d:
        raise UnsupportedDerivative("third and higher derivatives are not supported by the jet path")
    need = 1 if j == 0 else 2
    if b.order == 3:
        raise UnsupportedDerivative("internal: Laplace-mode blocks are only produced for traced residuals")
    if b.order < need:
        # the plan under-estimated this call: run it again at the needed order and remember
        if b.plan is not None and b.plan_idx >= 0:
            b.plan.raise_order(b.plan_idx, need)
        out = None
        for other in _blocks.values():  # an earlier escalation of the same call may still be alive
            if other.engine is b.engine and other.arg_ptr == b.arg_ptr and other.n == b.n \
                    and other.order >= need and other.order <= 2 and other.out_ref() is not None \
                    and other.generation == b.engine.generation and other.grad_mode == torch.is_grad_enabled():
                out, b = other.out_ref(), other
                break
        if out is None:
            out = _run_block(b.engine, arg, need, b.plan, b.plan_idx)
            b = _blocks[out.untyped_storage().data_ptr()]
    else:
        out = b.out_ref()
    if j == 0:
        return out[1:1 + d, :, k].t()
    i = j - 1
    if b.order == 4:      # full-Hessian block: the off-diagonal columns are real
        pairs = [(a, c) for a in range(d) for c in range(a + 1, d)]
        cols = [out[1 + d + i, :, k] if c == i else out[1 + 2 * d + pairs.index((min(i, c), max(i, c))), :, k]
                for c in range(d)]
    else:
        cols = [_nan_column(b.n, out)] * d
        cols = list(cols)
        cols[i] = out[1 + d + i, :, k]
    res = torch.stack(cols, dim=1)
    res._jet_block = out  # keep the block (and its registry entry) alive as long as the result is
    return res


def unpack(batch: torch.Tensor) -> tuple:
    """Same as the reference's ``unpack`` (condition/diff.py:4-22): columns of ``batch`` as views."""
    if _trace.is_traced(batch):
        return _trace.traced_unpack(batch)
    return torch.unbind(batch, dim=1)
