"""Fourier-feature network (reference ``multipinn/neural_network/fourier_features.py:9-76``):
``gamma(x) = [sin(xB), cos(xB)]`` with ``B ~ N(0, sigma^2)`` followed by an ``FNN``.
The encoding jets are evaluated in registers inside the fused kernel (``x@B`` once, not twice).  A frozen ``B`` (the
default) runs on the tcgen05 engine; ``is_trainable=True`` adds ``dL/dB`` to the reverse pass of the FP32-FMA engine."""
from typing import List, Literal

import torch
import torch.nn as nn

from .base import JetModule, ProgramBuilder
from .feedforward_neural_network import FNN, XavierFNN, fnn_rows


class FourierEncoding(nn.Module):
    def __init__(self, input_dim: int, encoding_dim: int, sigma: float, is_trainable: bool = False):
        super().__init__()
        self.input_dim = input_dim
        self.encoding_dim = encoding_dim
        self.sigma = sigma
        self.is_trainable = is_trainable
        w = torch.zeros((input_dim, encoding_dim // 2))
        nn.init.normal_(w)
        self.B = nn.Parameter(w * self.sigma, requires_grad=is_trainable)

    def extra_repr(self) -> str:
        return f"input_dim={self.input_dim}, encoding_dim={self.encoding_dim}, sigma={self.sigma}"

    def forward(self, x: torch.Tensor):
        z = x @ self.B
        return torch.cat([torch.sin(z), torch.cos(z)], dim=1)


class FourierFeatureNetwork(JetModule):
    def __init__(self, input_dim: int, output_dim: int, hidden_layers: List[int], encoding_dim: int,
                 sigma: float = 1.0, is_trainable: bool = False, xavier_init: bool = False,
                 xavier_init_mode: Literal["norm", "uniform"] = "norm", use_jit: bool = False):
        super().__init__()
        self.fourier_encoding = FourierEncoding(input_dim, encoding_dim, sigma, is_trainable)
        if xavier_init:
            self.fnn = XavierFNN(encoding_dim, output_dim, hidden_layers, init_mode=xavier_init_mode)
        else:
            self.fnn = FNN(encoding_dim, output_dim, hidden_layers)

    def jet_program(self):
        b = ProgramBuilder(self)
        b.fourier(self.fourier_encoding.B)
        fnn_rows(b, self.fnn.network_layers)
        return b.array()
