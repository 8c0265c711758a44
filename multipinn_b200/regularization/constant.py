from typing import List

import torch

from .basic import BasicLosses


class ConstantLosses(BasicLosses):
    """Fixed per-component weights (reference ``regularization/constant.py:8-36``)."""

    def __init__(self, lambdas: List[float]):
        super().__init__()
        self.lambdas = lambdas

    def __call__(self, trainer):
        losses = trainer.pinn.calculate_loss()
        if len(losses) != len(self.lambdas):
            raise Exception(
                f"The number of lambdas(number = {len(self.lambdas)}) does not match the number of losses(number = {len(losses)})")
        losses = torch.stack([los * lam for los, lam in zip(losses, self.lambdas)])
        return torch.sum(losses), losses
