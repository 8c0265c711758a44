"""Per-point loss weights trained by gradient ascent ("Investigating and mitigating failure modes in PINNs"), same
behaviour as the reference's ``PointwiseLosses`` (``multipinn/regularization/pointwise.py:9-152``):

    total = sum_k (k ignored ? sum_n r_kn^2 : sum_n lambda_kn r_kn^2),   lambda_kn initialised to 1,
    lambda <- maximising optimizer step (Adam, lr) on d total / d lambda before every evaluation but the first,
    reported losses = [mean_n r_kn^2]_k of the unweighted residuals.

Everything is ordinary autograd on the residual vectors, so it runs on the fused path as is; the one difference
from the reference is that the residuals are evaluated once per call and shared between the weighted total and the
reported losses (the reference evaluates them twice, three times on a first launch).  Quirks kept on purpose: the
weights of a batch are a bias-free ``Linear(N, 1)``; after a grid update the weight models are rebuilt but new
optimizers are appended while ``optimizers[current_batch]`` keeps being stepped (reference :137-152).
"""
import itertools
from typing import List

import torch

from .basic import BasicLosses


class PointwiseLosses(BasicLosses):
    def __init__(self, ignored_indexes: List[int] = None, optimizer: torch.optim.Optimizer = torch.optim.Adam,
                 lr: float = 0.01):
        super().__init__()
        self.optimizer_def = optimizer
        self.optimizers = []
        self.lr = lr
        self.models = []
        self.ignored_indexes = ignored_indexes if ignored_indexes is not None else []
        self.ignored_mask = None

    def _residuals(self, trainer):
        model = trainer.pinn.model
        return [r for condition in trainer.pinn.conditions for r in condition.get_residual(model)]

    def __call__(self, trainer):
        if trainer.current_epoch % trainer.update_grid_every == 0 and trainer.current_batch == 0:
            self.ignored_mask = None
            self.models = []
        residuals = self._residuals(trainer)
        if len(self.models) < trainer.num_batches:
            self.models.append([])
            self._first_launch(trainer, residuals)
        else:
            self.optimizers[trainer.current_batch].step()
            self.optimizers[trainer.current_batch].zero_grad()
        real_losses = torch.stack([(r.detach() ** 2).mean() for r in residuals], dim=0)
        losses = []
        for index, residual in enumerate(residuals):
            if self.ignored_mask[index]:
                losses.append((residual ** 2).sum())
            else:
                losses.append(self.models[trainer.current_batch][index](residual ** 2)[0])
        return torch.sum(torch.stack(losses, dim=0)), real_losses

    def _first_launch(self, trainer, residuals):
        self.ignored_mask = []
        params = []
        for index, residual in enumerate(residuals):
            if index in self.ignored_indexes:
                self.ignored_mask.append(True)
                self.models[trainer.current_batch].append(None)
            else:
                self.ignored_mask.append(False)
                model = torch.nn.Linear(residual.shape[0], 1, bias=False, device=residual.device)
                model.weight.data.fill_(1.0)
                self.models[trainer.current_batch].append(model)
                params.append(model.parameters())
        self.optimizers.append(self.optimizer_def(params=itertools.chain(*params), lr=self.lr, maximize=True))
