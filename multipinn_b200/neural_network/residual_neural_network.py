"""Residual MLP with the reference's layout
(reference ``multipinn/neural_network/residual_neural_network.py:10-126``): ``Linear -> Sine``; per
hidden size ``blocks[i]`` residual blocks followed by ``Linear -> GELU`` (the last stage has no
trailing Linear); a plain output ``Linear``.  Attribute ``layers`` and ``state_dict`` keys match."""
from __future__ import annotations

import torch.nn as nn

from .. import _lib
from .activation_function import GELU, Sine
from .base import JetModule, ProgramBuilder, _act_code
from .light_residual_block import LightResidualBlock


class ResNet(JetModule):
    def __init__(self, input_dim: int, output_dim: int, hidden_layers: list, blocks: list = None,
                 res_block: nn.Module = LightResidualBlock):
        super().__init__()
        widths = [int(w) for w in hidden_layers]
        counts = [int(c) for c in blocks] if blocks else [1] * len(widths)
        stages = [nn.Sequential(nn.Linear(input_dim, widths[0]), Sine())]
        # As in the reference (:54-70) the intermediate stages pair hidden_layers[:-1] with blocks[:-1], so a
        # `blocks` list shorter than `hidden_layers` drops the unpaired stages; the last stage always uses
        # blocks[-1] at hidden_layers[-1]; a stage never has fewer than one block (:102-105).
        for width, count, nxt in zip(widths[:-1], counts[:-1], widths[1:]):
            seq = [res_block(GELU(), width) for _ in range(max(count, 1))]
            seq.append(nn.Sequential(nn.Linear(width, nxt), GELU()))
            stages.append(nn.Sequential(*seq))
        stages.append(nn.Sequential(*[res_block(GELU(), widths[-1]) for _ in range(max(counts[-1], 1))]))
        stages.append(nn.Linear(widths[-1], output_dim))
        self.layers = nn.ModuleList(stages)

    def jet_program(self):
        b = ProgramBuilder(self)
        mods = list(self.layers)
        b.linear(mods[0][0], _act_code(mods[0][1]))
        for stage in mods[1:-1]:
            for m in stage:
                if isinstance(m, nn.Sequential):
                    b.linear(m[0], _act_code(m[1]))
                elif hasattr(m, "linear_first") and hasattr(m, "linear_second"):
                    b.linear(m.linear_first[0], _act_code(m.linear_first[1]), _lib.FLAG_SAVE_SKIP)
                    b.linear(m.linear_second, _act_code(m.activation), _lib.FLAG_ADD_SKIP)
                else:
                    raise NotImplementedError(f"residual block {type(m).__name__} has no jet lowering")
        b.linear(mods[-1], _lib.ACT_NONE)
        return b.array()
