"""Point generator with the reference's interface (reference ``multipinn/generation/generator.py:7-72``),
plus a ``device`` argument so points land where the model lives."""
from typing import Iterable, Union

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
