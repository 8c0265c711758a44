"""Gradient-statistics loss balancing (Wang et al., "Understanding and mitigating gradient flow pathologies in
PINNs"), same numbers as the reference's ``GradientLosses`` (``multipinn/regularization/gradient.py:25-59``):

    lambda_k <- (1 - alpha) lambda_k + alpha * max|dL_0/dtheta| / mean_tensors(mean|dL_k/dtheta|),   k >= 1
    total = L_0 + sum_k lambda_k L_k

The reference obtains the per-condition gradients with one extra ``backward(retain_graph=True)`` per condition
and then back-propagates the weighted total once more.  Here every condition owns its own fused reverse launch,
so the K per-condition gradients are computed exactly once, the statistics are taken on the flat gradient
vector, and the gradient of the weighted total is assembled as ``g_0 + sum_k lambda_k g_k`` -- the ``backward()``
the trainer then calls on the returned loss only hands that vector over.
"""
import torch

from ..neural_network.base import JetModule
from .basic import BasicLosses


class _PrecomputedGradient(torch.autograd.Function):
    """A scalar whose gradient w.r.t. the flat parameter vector is already known."""

    @staticmethod
    def forward(ctx, flat, value, grad):
        ctx.save_for_backward(grad)
        return value.detach().clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


class GradientLosses(BasicLosses):
    def __init__(self, alpha: float = 0.9):
        super().__init__()
        self.lambda_regularization = None
        self.alpha = alpha
        self._tensor_mean_weights = None

    def _weights(self, engine):
        """w_i = 1 / (T * numel(tensor(i))) over the T trainable tensors: sum_i |g_i| w_i == mean of per-tensor means."""
        if self._tensor_mean_weights is None or self._tensor_mean_weights.shape != engine.flat.shape:
            w = torch.zeros_like(engine.flat_grad)
            trainable = [(p, off) for p, off in zip(engine._params, engine._offsets) if p.requires_grad]
            for p, off in trainable:
                w[off:off + p.numel()] = 1.0 / (len(trainable) * p.numel())
            self._tensor_mean_weights = w
        return self._tensor_mean_weights

    def __call__(self, trainer):
        model = trainer.pinn.model
        losses = trainer.pinn.calculate_loss()
        if len(trainer.pinn.conditions) != len(losses):
            raise Exception("Regularization does not supported system ODE")
        if not isinstance(model, JetModule):
            raise TypeError("GradientLosses here works on the fused jet networks of multipinn_b200.neural_network")
        engine = model.jet_engine
        if self.lambda_regularization is None:
            self.lambda_regularization = torch.ones(len(losses) - 1, device=losses[0].device)
        grads = [torch.autograd.grad(los, engine.flat)[0] for los in losses]   # one fused reverse launch each
        w = self._weights(engine)
        max_f = grads[0].abs().max()
        total_grad = grads[0].clone()
        total = losses[0].detach().clone()
        for ind in range(1, len(losses)):
            regularization = max_f / (grads[ind].abs() * w).sum()
            regularization = (1 - self.alpha) * self.lambda_regularization[ind - 1] + self.alpha * regularization
            self.lambda_regularization[ind - 1] = regularization
            total_grad += regularization * grads[ind]
            total = total + regularization * losses[ind].detach()
        total_loss = _PrecomputedGradient.apply(engine.flat, total, total_grad)
        return total_loss, torch.stack([los.detach() for los in losses], dim=0)
