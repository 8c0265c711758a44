"""``Inference`` with the reference's interface (reference ``multipinn/neural_network/inference_wrapper.py:5-40``):
evaluate a model on a large point set in batches and stack the predictions.

The reference walks a ``DataLoader`` with 256-point mini-batches to bound the memory of an eager forward; a fused
jet network evaluates a million points in one value-only launch (no stash is written), so ``batchsize`` only caps
the size of one launch here (rows are independent: the result does not depend on it)."""
import torch
import torch.nn as nn


class Inference(nn.Module):
    def __init__(self, model: nn.Module, batchsize: int = 256):
        super().__init__()
        self.batchsize = batchsize
        self.model = model
        self.device = next(model.parameters()).device
        self.launch_points = max(int(batchsize), 1 << 20)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device)
        outs = [self.model.forward(x[i:i + self.launch_points]) for i in range(0, max(len(x), 1), self.launch_points)]
        return torch.vstack(outs)
