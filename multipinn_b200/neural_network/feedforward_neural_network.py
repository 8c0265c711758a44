"""Fully-connected network with the reference's layout
(reference ``multipinn/neural_network/feedforward_neural_network.py:9-87``):
``Linear -> Sine``, then ``Linear -> GELU`` per further hidden size, then a plain ``Linear``.
Attribute names (``network_layers``) and ``state_dict`` keys are the reference's."""
from typing import List, Literal

import torch.nn as nn

from .. import _lib
from .activation_function import GELU, Sine
from .base import JetModule, ProgramBuilder, _act_code


def fnn_rows(builder: ProgramBuilder, network_layers) -> None:
    mods = list(network_layers)
    for m in mods[:-1]:
        builder.linear(m[0], _act_code(m[1]))
    builder.linear(mods[-1], _lib.ACT_NONE)


class FNN(JetModule):
    def __init__(self, input_dim: int, output_dim: int, hidden_layers: List[int]):
        super().__init__()
        widths = [int(w) for w in hidden_layers]  # accepts an OmegaConf ListConfig
        stack = [nn.Sequential(nn.Linear(input_dim, widths[0]), Sine())]
        for fan_in, fan_out in zip(widths[:-1], widths[1:]):
            stack.append(nn.Sequential(nn.Linear(fan_in, fan_out), GELU()))
        stack.append(nn.Linear(widths[-1], output_dim))
        self.network_layers = nn.ModuleList(stack)

    def jet_program(self):
        b = ProgramBuilder(self)
        fnn_rows(b, self.network_layers)
        return b.array()


class XavierFNN(FNN):
    def __init__(self, input_dim: int, output_dim: int, hidden_layers: List[int],
                 init_mode: Literal["norm", "uniform"] = "norm"):
        super().__init__(input_dim, output_dim, hidden_layers)
        init = {"norm": nn.init.xavier_normal_, "uniform": nn.init.xavier_uniform_}[init_mode]
        for m in self.network_layers:
            lin = m[0] if isinstance(m, nn.Sequential) else m
            init(lin.weight)
            nn.init.zeros_(lin.bias)
