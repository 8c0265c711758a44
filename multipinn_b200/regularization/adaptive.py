"""Self-adaptive loss weights trained by gradient ASCENT on the weighted loss, same numbers as the reference's
``AdaptiveConditionsLosses`` / ``AdaptiveConditionsAndPointsLosses`` (``multipinn/regularization/adaptive.py:7-242``):

    w = softmax(theta),  theta_0 = log(initial weights),  total = sum_k w_k L_k,
    theta <- Adam(lr, maximize=True, weight_decay=1e-6) step on d total / d theta, taken at the START of the next call

Condition level: ``L_k`` are the component means and ``theta_0 = log(N_k)``.  Point level (the "AndPoints" variant):
every component first reduces its per-point squared residuals with its own softmax weights (``theta_0 = 0``, i.e.
the plain mean), re-created whenever the trainer regenerates the points.

The weights only multiply residual sums, so everything below the loss values is the unchanged fused jet path: the
component losses are autograd scalars whose backward is one fused reverse launch per condition.
"""
from typing import List

import torch
from torch import nn

from .basic import BasicLosses


class AdaptiveWeightSum(nn.Module):
    """``sum(x * softmax(theta))`` with ``theta`` ascending the value it produced on the previous call."""

    def __init__(self, initial_weights, lr: float = 1e-3):
        super().__init__()
        self.lr = lr
        self.weight = torch.as_tensor(initial_weights, dtype=torch.float32).detach()
        self.train_weight = nn.Parameter(torch.log(self.weight))
        self.optimizer = torch.optim.Adam((self.train_weight,), lr=lr, maximize=True, weight_decay=1e-6)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.train_weight.device != x.device:
            self.to(x.device)
        # the gradient left by the backward() of the previous step's total loss moves theta first
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.weight = torch.softmax(self.train_weight, dim=0)
        return torch.sum(x * self.weight)


class AdaptiveConditionsLosses(BasicLosses):
    def __init__(self, lr: float = 1e-3):
        super().__init__()
        self.lr = lr
        self.adaptive_cond = None

    def weight_sum(self, mean_losses: List[torch.Tensor]):
        stacked = torch.stack(mean_losses)
        return self.adaptive_cond(stacked), stacked.detach()

    def __call__(self, trainer):
        if self.adaptive_cond is None:
            mean_losses, counts = trainer.pinn.calculate_loss_and_count()
            self.adaptive_cond = AdaptiveWeightSum(counts, self.lr)
        else:
            mean_losses = trainer.pinn.calculate_loss()
        return self.weight_sum(mean_losses)


class AdaptiveConditionsAndPointsLosses(AdaptiveConditionsLosses):
    def __init__(self, lr: float = 1e-3):
        super().__init__(lr)
        self.adaptive_points = tuple()

    def __call__(self, trainer):
        per_point = []
        for cond in trainer.pinn.conditions:
            per_point += trainer.pinn.condition_loss(cond)
        if trainer.current_epoch % trainer.update_grid_every == 0:   # new points: start their weights afresh
            self.adaptive_points = tuple(AdaptiveWeightSum(torch.ones_like(c), self.lr) for c in per_point)
            if self.adaptive_cond is None:
                self.adaptive_cond = AdaptiveWeightSum([len(c) for c in per_point], self.lr)
        reduced = [self.adaptive_points[i](c) for i, c in enumerate(per_point)]
        return self.weight_sum(reduced)
