"""``Condition`` / ``ConditionExtra`` with the reference's interface
(reference ``multipinn/condition/condition.py:6-215``): a residual callable bound to a geometry and
a point generator, sliced per batch / per rank.

Addition over the reference: each (condition, model) pair keeps a ``ResidualPlan`` recording which
jet order every ``model(arg)`` call inside the residual function needs.  It is filled by the probe
evaluation the reference already performs in ``init_output_len`` and consulted on every later call,
so the fused forward computes exactly the derivative channels the residual will ask for.
"""
from __future__ import annotations

import weakref
from typing import Callable, List, Union

import numpy as np
import torch

from .. import trace as _trace
from ..jet import ResidualPlan, UnsupportedDerivative, use_plan


def closure_fingerprint(fn, depth: int = 3) -> tuple:
    """Numbers and small tensors a residual callable reads from its closure cells and module globals (recursing into
    the plain functions it closes over).  A traced program has them baked in as constants; examples such as
    ``examples/convection_1D_gradual_learning/problem.py:16-19`` change one through a ``nonlocal`` setter during
    training (reference ``trainer/gradual_trainer.py:30-35``), which must re-trace the residual."""
    items, seen = [], set()

    def value(v):
        if isinstance(v, (bool, int, float, complex, str)):
            return v
        if isinstance(v, torch.Tensor) and v.numel() <= 16:
            if v.is_cuda:       # reading it back would synchronise the device at every step: identity + version instead
                return ("cuda-tensor", id(v), v._version)
            return tuple(v.detach().reshape(-1).tolist())
        if isinstance(v, np.ndarray) and v.size <= 16:
            return tuple(v.reshape(-1).tolist())
        if isinstance(v, (list, tuple)) and len(v) <= 16 and all(isinstance(e, (bool, int, float)) for e in v):
            return tuple(v)
        return None

    def visit(f, d):
        code = getattr(f, "__code__", None)
        if code is None or id(f) in seen or d < 0:
            return
        seen.add(id(f))
        for name, cell in zip(code.co_freevars, f.__closure__ or ()):
            try:
                v = cell.cell_contents
            except ValueError:
                continue
            val = value(v)
            if val is not None:
                items.append((name, val))
            elif callable(v):
                visit(v, d - 1)
        g = getattr(f, "__globals__", {})
        for name in code.co_names:
            val = value(g.get(name)) if name in g else None
            if val is not None:
                items.append(("global:" + name, val))

    visit(fn, depth)
    return tuple(items)


class Condition:
    def __init__(self, function: Callable, geometry) -> None:
        self.function = function
        self.geometry = geometry
        self.generator = None
        self.points = None
        self.batch_size = None
        self.batch_points = None
        self.output_len = None
        self._plans = weakref.WeakKeyDictionary()
        self._programs = weakref.WeakKeyDictionary()   # model -> traced ResidualProgram, or None (eager path)
        self._fingerprints = weakref.WeakKeyDictionary()   # model -> closure_fingerprint at trace time
        self.trace_note = None                         # why the last trace attempt fell back, for diagnostics

    # -- jet-order bookkeeping -------------------------------------------------------------
    def plan_for(self, model) -> ResidualPlan:
        try:
            plan = self._plans.get(model)
        except TypeError:  # unhashable / non-weakrefable stand-ins (mocks)
            return None
        if plan is None:
            plan = ResidualPlan()
            self._plans[model] = plan
        return plan

    # -- traced residual programs (fused train step) ------------------------------------------
    def program_for(self, model):
        """The device program this condition's residual was traced into for ``model``, or None.  Re-traced when a
        constant the callable closes over has changed since (see ``closure_fingerprint``)."""
        try:
            prog = self._programs.get(model)
        except TypeError:
            return None
        if model in self._programs and not self._same_residual(model):
            self._trace(model, self._probe_point(model))
            prog = self._programs.get(model)
        return prog

    def _same_residual(self, model) -> bool:
        """The callable traced for ``model`` is still this condition's ``function`` (MarchingTrainer stages replace
        it: examples/convection_1D_marching/problem.py:53-54) and closes over the same constants."""
        fn, fp = self._fingerprints.get(model, (None, None))
        return fn is self.function and fp == closure_fingerprint(self.function)

    def _trace(self, model, probe: torch.Tensor) -> None:
        """Try to record the residual as a device program and check it against the eager residual at a few
        points of the bounding box.  Any failure leaves the condition on the eager jet path."""
        from ..neural_network.base import JetModule
        self.trace_note = None
        if not isinstance(model, JetModule) or \
                type(self).get_residual not in (Condition.get_residual, ConditionExtra.get_residual):
            return
        try:
            self._programs[model] = None
            self._fingerprints[model] = (self.function, closure_fingerprint(self.function))
        except TypeError:
            return
        try:
            prog = _trace.trace_residual(self.function, model, model.jet_engine.d, probe, self._trace_data(probe))
            self._check_program(model, prog, probe)
        except (_trace.TraceError, UnsupportedDerivative) as e:
            self.trace_note = f"{type(e).__name__}: {e}"
            return
        except Exception as e:  # user code doing something the symbolic columns do not support
            self.trace_note = f"{type(e).__name__}: {e}"
            return
        self._programs[model] = prog

    def _trace_data(self, probe: torch.Tensor):
        """Sample of the per-point data the residual takes besides the points (None for a plain Condition)."""
        return None

    def _check_points(self, probe: torch.Tensor, n: int) -> torch.Tensor:
        low = torch.as_tensor(self.geometry.bbox[0], dtype=torch.float32, device=probe.device)
        high = torch.as_tensor(self.geometry.bbox[1], dtype=torch.float32, device=probe.device)
        gen = torch.Generator(device="cpu").manual_seed(1234)
        # (device named explicitly: the reference's run scripts set torch.set_default_device("cuda"), utils/device.py:47)
        u = torch.rand((n, low.numel()), generator=gen, dtype=torch.float32, device="cpu").to(probe.device)
        return (low + (high - low) * u).contiguous()

    def _check_program(self, model, prog, probe: torch.Tensor, n: int = 8) -> None:
        from ..fused import ConditionRunner
        x = self._check_points(probe, n)
        with torch.no_grad():
            got = ConditionRunner(self, prog, model.jet_engine).residual_values(x)
        want = self.get_residual_fn(model)(x.clone().requires_grad_())
        if len(want) != got.shape[0]:
            raise _trace.TraceError("traced program has a different number of components")
        for k, w in enumerate(want):
            w = w.detach().reshape(-1).to(got.dtype)
            scale = float(w.abs().max()) + 1e-6
            if not bool(((got[k] - w).abs() <= 2e-3 * scale + 1e-5).all()):
                raise _trace.TraceError(
                    f"traced program disagrees with the eager residual on component {k} "
                    f"(max diff {float((got[k] - w).abs().max()):.3e}, scale {scale:.3e})")

    def _call(self, model, *args):
        with use_plan(self.plan_for(model)):
            return self.function(model, *args)

    # -- reference interface ---------------------------------------------------------------
    def update_points(self, model=None) -> None:
        self.points = self.generator.generate(self, model)

    def select_batch(self, i: int) -> None:
        start = i * self.batch_size
        self.batch_points = self.points[start:start + self.batch_size]

    def get_residual(self, model) -> List[torch.Tensor]:
        return self._call(model, self.batch_points)

    def get_residual_fn(self, model) -> Callable:
        def residual(arg):
            return self._call(model, arg)

        return residual

    def set_batching(self, num_batches: int) -> None:
        assert self.generator.n_points % num_batches == 0, "Points must divide evenly into batches"
        self.batch_size = self.generator.n_points // num_batches

    def _probe_point(self, model) -> torch.Tensor:
        device = next(model.parameters()).device
        low = torch.as_tensor(self.geometry.bbox[0], dtype=torch.float32, device=device)
        high = torch.as_tensor(self.geometry.bbox[1], dtype=torch.float32, device=device)
        return ((low + high) * 0.5).reshape(1, -1).requires_grad_()

    def init_output_len(self, model) -> None:
        """Evaluate the residual once at the bbox centre: sets ``output_len`` (as the reference does)
        and, as a side effect, fills the jet-order plan and validates that the residual only uses
        derivatives the jet path provides."""
        probe = self._probe_point(model)
        residuals = self.get_residual_fn(model)(probe)
        if any(isinstance(r, torch.Tensor) and bool(torch.isnan(r).any()) for r in residuals) and probe.shape[1] >= 2:
            # a NaN-poisoned mixed second derivative is used: raise every second-order model call of this residual to
            # the full-Hessian mode and evaluate again
            plan = self.plan_for(model)
            if plan is not None and any(o == 2 for o in plan.orders):
                plan.orders = [4 if o == 2 else o for o in plan.orders]
                residuals = self.get_residual_fn(model)(probe)
        self.output_len = len(residuals)
        for k, r in enumerate(residuals):
            if isinstance(r, torch.Tensor) and bool(torch.isnan(r).any()):
                raise UnsupportedDerivative(
                    f"residual component {k} of {getattr(self.function, '__name__', self.function)} is NaN at "
                    "the probe point (also with the full-Hessian jets): it uses a derivative the fused jet path "
                    "does not compute, or the residual itself is not finite there.")
        self._trace(model, probe)


class ConditionExtra(Condition):
    """Condition whose residual takes auxiliary per-point data, e.g. boundary normals
    (reference ``condition/condition.py:117-215``)."""

    def __init__(self, function: Callable, geometry, data_gen: List[Union[str, Callable]]):
        super().__init__(function, geometry)
        data_gen = list(data_gen)
        if "normals" in data_gen:
            assert hasattr(geometry, "boundary_normal"), "Geometry must support normal vectors"
            data_gen[data_gen.index("normals")] = ConditionExtra.generator_for_normals(geometry)
        self.data_gen = lambda points: tuple(fn(points) for fn in data_gen)
        self.data = None
        self.batch_data = None

    @staticmethod
    def generator_for_normals(geometry) -> Callable:
        return lambda points: torch.tensor(
            geometry.boundary_normal(points.detach().cpu().numpy()), device=points.device)

    def update_points(self, model=None) -> None:
        self.points = self.generator.generate(self, model)
        self.data = self.data_gen(self.points)

    def select_batch(self, i: int) -> None:
        start = i * self.batch_size
        self.batch_points = self.points[start:start + self.batch_size]
        self.batch_data = self.data[start:start + self.batch_size]

    def get_residual(self, model) -> List[torch.Tensor]:
        return self._call(model, self.batch_points, self.batch_data)

    def get_residual_fn(self, model) -> Callable:
        def residual(arg):
            return self._call(model, arg, self.data_gen(arg))

        return residual

    def _trace_data(self, probe: torch.Tensor):
        return self.data_gen(probe)

    def _check_points(self, probe: torch.Tensor, n: int) -> torch.Tensor:
        # data generators such as boundary normals are only meaningful ON the geometry: check there when it can
        # be sampled (fixed seed; the caller's numpy random stream is left as it was)
        try:
            import numpy as np
            state = np.random.get_state()
            np.random.seed(1234)
            try:
                pts = self.geometry.random_points(n, "pseudo")
            finally:
                np.random.set_state(state)
            x = torch.as_tensor(np.asarray(pts), dtype=torch.float32, device=probe.device).contiguous()
            if x.dim() == 2 and x.shape[0] >= 1 and x.shape[1] == probe.shape[1] and bool(torch.isfinite(x).all()):
                return x
        except Exception:   # noqa: BLE001 -- geometry without a sampler: fall back to the bounding box
            pass
        return super()._check_points(probe, n)
