"""Residual block ``act(W2 act(W1 x + b1) + b2 + x)``
(reference ``multipinn/neural_network/light_residual_block.py:5-38``).  Inside ``ResNet`` it is
compiled into two LINEAR ops with SAVE_SKIP / ADD_SKIP flags; stand-alone it is a plain module."""
import torch
import torch.nn as nn


class LightResidualBlock(nn.Module):
    def __init__(self, activation, features: int):
        super().__init__()
        self.linear_first = nn.Sequential(nn.Linear(features, features), activation)
        self.linear_second = nn.Linear(features, features)
        self.activation = activation

    def forward(self, x: torch.Tensor):
        return self.activation(self.linear_second(self.linear_first(x)) + x)
