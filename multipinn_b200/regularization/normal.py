"""Loss balancing by the spread of the last layer's gradient, same numbers as the reference's ``NormalLosses``
(``multipinn/regularization/normal.py:6-45``):

    lambda_k <- (1 - alpha) lambda_k + alpha * std(dL_0/dW_last) / std(dL_k/dW_last),   k >= 1
    total = L_0 + sum_k lambda_k L_k

``W_last`` is found with the reference's own expression, ``list(list(model.children())[-1].children())[-1]``.
As in ``GradientLosses`` here, the K per-condition gradients come from one fused reverse launch each (the reference
back-propagates every loss twice), the statistics are read from the flat gradient vector and the gradient of the
weighted total is assembled from them.
"""
import torch

from ..neural_network.base import JetModule
from .basic import BasicLosses
from .gradient import _PrecomputedGradient


class NormalLosses(BasicLosses):
    def __init__(self, alpha: float = 0.9):
        super().__init__()
        self.lambda_regularization = None
        self.alpha = alpha

    def __call__(self, trainer):
        """Requirement (as in the reference): the first loss belongs to the inner area."""
        model = trainer.pinn.model
        losses = trainer.pinn.calculate_loss()
        if len(trainer.pinn.conditions) != len(losses):
            raise Exception("Regularization does not supported system ODE")
        if not isinstance(model, JetModule):
            raise TypeError("NormalLosses here works on the fused jet networks of multipinn_b200.neural_network")
        engine = model.jet_engine
        if self.lambda_regularization is None:
            self.lambda_regularization = torch.ones(len(losses) - 1, device=losses[0].device)
        last_layers = list(list(model.children())[-1].children())[-1]
        off = next(o for p, o in zip(engine._params, engine._offsets) if p is last_layers.weight)
        sl = slice(off, off + last_layers.weight.numel())
        grads = [torch.autograd.grad(los, engine.flat)[0] for los in losses]   # one fused reverse launch each
        var_f = torch.std(grads[0][sl])
        total_grad = grads[0].clone()
        total = losses[0].detach().clone()
        for ind in range(1, len(losses)):
            regularization = var_f / torch.std(grads[ind][sl])
            regularization = (1 - self.alpha) * self.lambda_regularization[ind - 1] + self.alpha * regularization
            self.lambda_regularization[ind - 1] = regularization
            total_grad += regularization * grads[ind]
            total = total + regularization * losses[ind].detach()
        total_loss = _PrecomputedGradient.apply(engine.flat, total, total_grad)
        return total_loss, torch.stack([los.detach() for los in losses], dim=0)
