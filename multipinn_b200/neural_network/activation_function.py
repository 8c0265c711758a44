"""Activation modules with the reference's names and values
(reference ``multipinn/neural_network/activation_function.py:5-54``).

Inside a network these never run as separate ops: the jet kernels evaluate the activation and its
first three derivatives in the GEMM epilogue (csrc/common.cuh).  The modules exist so the layer
structure, ``state_dict`` keys and pickles match the reference, and for stand-alone element-wise use.
"""
import torch
from torch import nn

GELU_DIVISOR = 1.41421  # the reference's literal, not sqrt(2)


def fused_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * 0.5 * (1.0 + torch.erf(x / GELU_DIVISOR))


def fused_sin(x: torch.Tensor) -> torch.Tensor:
    return torch.sin(x)


class GELU(nn.Module):
    def forward(self, input):
        return fused_gelu(input)


class Sine(nn.Module):
    def forward(self, input):
        return fused_sin(input)
