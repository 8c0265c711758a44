"""Loss assembly with the reference's interface (reference ``multipinn/PINN/PINN.py:8-158``).

Every per-component ``sum(r**2)`` goes through the fused sum-of-squares kernel (strided input,
warp-shuffle reduction, one atomic per CTA) instead of materialising ``r**2``."""
from typing import List

import torch

from ..condition.condition import Condition
from ..engine import backend


class _SumSquares(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r):
        acc = torch.zeros(1, dtype=r.dtype, device=r.device)
        backend().sumsq(r.detach(), acc)
        ctx.save_for_backward(r)
        return acc.reshape(())

    @staticmethod
    def backward(ctx, g):
        (r,) = ctx.saved_tensors
        return r * (2.0 * g)


def sum_squares(r: torch.Tensor) -> torch.Tensor:
    """``torch.sum(r**2)`` for a 1-D residual vector."""
    if r.dim() != 1 or r.dtype != torch.float32:
        return torch.sum(r**2)
    return _SumSquares.apply(r)


class PINN:
    def __init__(self, model: torch.nn.Module, conditions: List[Condition]) -> None:
        self.model = model
        self.conditions = conditions
        self.model.eval()
        for cond in self.conditions:
            cond.init_output_len(self.model)

    def update_data(self) -> None:
        for cond in self.conditions:
            cond.update_points(model=self.model)

    def select_batch(self, i) -> None:
        for cond in self.conditions:
            cond.select_batch(i)

    def condition_loss(self, cond: Condition) -> List[torch.Tensor]:
        return [r**2 for r in cond.get_residual(self.model)]

    @staticmethod
    def count(loss_arr):
        return [len(x) for x in loss_arr]

    @staticmethod
    def mean_loss(loss_arr):
        return [torch.mean(x) for x in loss_arr]

    @staticmethod
    def sum_loss(loss_arr):
        return [torch.sum(x) for x in loss_arr]

    def _sums_and_counts(self):
        sums, counts = [], []
        for cond in self.conditions:
            for r in cond.get_residual(self.model):
                sums.append(sum_squares(r))
                counts.append(len(r))
        return sums, counts

    def calculate_loss(self) -> List[torch.Tensor]:
        """Per-component means (``calc_loss="legacy"`` sums them)."""
        sums, counts = self._sums_and_counts()
        return [s / c for s, c in zip(sums, counts)]

    def calculate_mean_loss(self):
        """``(sum_k sum_n r_kn^2 / sum_k N_k, [mean_n r_kn^2]_k)`` (reference PINN.py:109-125)."""
        sums, counts = self._sums_and_counts()
        total_mean = torch.sum(torch.stack(sums)) / sum(counts)
        return total_mean, [s / c for s, c in zip(sums, counts)]

    def calculate_loss_and_count(self):
        sums, counts = self._sums_and_counts()
        return [s / c for s, c in zip(sums, counts)], counts

    def calculate_loss_on_points(self, cond: Condition, points: torch.Tensor) -> List[torch.Tensor]:
        residuals = cond.get_residual_fn(self.model)(points)
        return [sum_squares(r) / len(r) for r in residuals]
