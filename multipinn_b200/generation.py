"""Point generators with the reference's interface (reference ``multipinn/generation/generator.py:7-72``),
plus a ``device`` argument so points land where the model lives, and two generators that never leave the device:
uniform sampling of box geometries, and residual-driven refinement (the reference's RAR-D scheme,
``generation/adaptive_generator.py:12-160``) whose error field comes from the value-only residual program
(``fused.ConditionRunner.residual_values``: forward jets + residual kernel, no stash, no reverse pass)."""
from typing import Iterable, Optional, Union

import numpy as np
import torch

from .condition.condition import Condition


class Generator:
    def __init__(self, n_points, sampler="pseudo", device=None):
        self.n_points = n_points
        self.sampler = sampler
        self.device = device

    def generate(self, condition: Condition, model):
        points = condition.geometry.random_points(self.n_points, self.sampler)
        device = self.device
        if device is None and model is not None:
            device = next(model.parameters()).device
        return torch.as_tensor(points, device=device).requires_grad_()

    def use_for(self, condition: Union[Condition, Iterable[Condition]]) -> None:
        if isinstance(condition, Condition):
            condition.generator = self
        else:
            for cond in condition:
                cond.generator = self


def _box_of(geometry):
    """(low, high) of an axis-aligned box geometry (this package's or the reference's ``Hypercube``), else None."""
    low, high = getattr(geometry, "low", None), getattr(geometry, "high", None)
    if low is None or high is None or type(geometry).__name__ != "Hypercube":
        return None
    return np.asarray(low, dtype="float32"), np.asarray(high, dtype="float32")


class DeviceGenerator(Generator):
    """``Generator(n, "pseudo")`` for box geometries with the random numbers drawn on the device
    (``low + side * U[0,1)``, the formula of the reference's ``Hypercube.random_points``, geometry/domain_nd.py:116-118):
    no host sampling and no host-to-device copy when the points are regenerated every epoch.  Other geometries
    fall back to the host sampler of the base class."""

    def __init__(self, n_points, sampler="pseudo", device=None, seed: Optional[int] = None):
        super().__init__(n_points, sampler, device)
        self.seed = seed
        self._rng = None

    def _uniform(self, n, geometry, device):
        box = _box_of(geometry)
        if box is None or self.sampler != "pseudo":
            return None
        if self._rng is None or self._rng.device != device:
            self._rng = torch.Generator(device=device)
            if self.seed is not None:
                self._rng.manual_seed(self.seed)
        low = torch.as_tensor(box[0], device=device)
        side = torch.as_tensor(box[1] - box[0], device=device)
        return torch.rand((n, len(box[0])), generator=self._rng, device=device, dtype=torch.float32) * side + low

    def generate(self, condition: Condition, model):
        device = self.device
        if device is None and model is not None:
            device = next(model.parameters()).device
        pts = self._uniform(self.n_points, condition.geometry, torch.device(device) if device is not None else torch.device("cpu"))
        if pts is None:
            return super().generate(condition, model)
        return pts.requires_grad_()


class ResidualAdaptiveGenerator(DeviceGenerator):
    """Residual-based adaptive refinement with distribution (RAR-D, reference ``AdaptiveGeneratorRAR_D``,
    adaptive_generator.py:76-160) on the device: start uniform; at every later update evaluate the error field
    ``sum_k |r_k|`` on ``density_rec_points_num`` fresh uniform points, draw ``add_points`` of them without
    replacement with probability ``err^p / mean(err^p) + c`` and append them, up to ``n_points_up_bnd``.
    The error field is one value-only pass of the traced residual program when the condition has one (else the eager
    residual under ``no_grad``)."""

    def __init__(self, n_points, power_coeff=3, add_coeff=1, density_rec_points_num=10000, add_points=50,
                 n_points_up_bnd=None, device=None, seed=None):
        super().__init__(n_points, "pseudo", device, seed)
        self.power_coeff, self.add_coeff = power_coeff, add_coeff
        self.density_rec_points_num = density_rec_points_num
        self.add_points = add_points
        self.n_points_up_bnd = n_points_up_bnd if n_points_up_bnd is not None else density_rec_points_num

    def error_field(self, condition: Condition, model, pool: torch.Tensor) -> torch.Tensor:
        prog = condition.program_for(model)
        if prog is not None:
            from .fused import ConditionRunner
            with torch.no_grad():
                return ConditionRunner(condition, prog, model.jet_engine).residual_values(pool).abs().sum(dim=0)
        with torch.no_grad():
            res = condition.get_residual_fn(model)(pool)
        return torch.stack([r.detach() for r in res], dim=1).abs().sum(dim=1)

    def generate(self, condition: Condition, model):
        if condition.points is None:
            return super().generate(condition, model)
        if condition.points.shape[0] >= self.n_points_up_bnd:
            return condition.points
        device = condition.points.device
        pool = self._uniform(self.density_rec_points_num, condition.geometry, device)
        if pool is None:
            pool = torch.as_tensor(condition.geometry.random_points(self.density_rec_points_num), device=device)
        err = self.error_field(condition, model, pool.contiguous())
        w = err.double() ** self.power_coeff
        w = w / w.mean() + self.add_coeff
        idx = torch.multinomial(w / w.sum(), self.add_points, replacement=False, generator=self._rng)
        return torch.cat([condition.points.detach(), pool[idx]], dim=0).requires_grad_()
