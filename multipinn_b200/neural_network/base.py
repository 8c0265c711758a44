"""Common base of the mirrored network classes: builds the layer program once and routes
``forward`` through the fused jet engine."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, jet
from ..engine import JetEngine
from .activation_function import GELU, Sine


def _act_code(module) -> int:
    if isinstance(module, Sine):
        return _lib.ACT_SINE
    if isinstance(module, GELU):
        return _lib.ACT_GELU
    raise NotImplementedError(
        f"activation {type(module).__name__} has no jet implementation (supported: Sine, GELU)")


class ProgramBuilder:
    """Collects ``mpinn_op_t`` rows; offsets follow ``model.parameters()`` order."""

    def __init__(self, model: nn.Module):
        self.offsets: Dict[int, int] = {}
        off = 0
        for p in model.parameters():
            self.offsets[id(p)] = off
            off += p.numel()
        self.rows: List[List[int]] = []

    def linear(self, lin: nn.Linear, act: int, flags: int = 0):
        self.rows.append([_lib.OP_LINEAR, lin.in_features, lin.out_features,
                          self.offsets[id(lin.weight)], self.offsets[id(lin.bias)], act, flags, 0])

    def fourier(self, B: torch.Tensor):
        flags = _lib.FLAG_TRAINABLE if B.requires_grad else 0   # a trainable encoding runs on the FP32-FMA engine
        self.rows.append([_lib.OP_FOURIER, B.shape[0], 2 * B.shape[1], self.offsets[id(B)], 0, 0, flags, 0])

    def array(self) -> np.ndarray:
        return np.asarray(self.rows, dtype=np.int32)


class JetModule(nn.Module):
    """nn.Module whose forward is one fused jet-kernel launch."""

    def jet_program(self) -> np.ndarray:  # pragma: no cover - overridden
        raise NotImplementedError

    @property
    def jet_engine(self) -> JetEngine:
        eng = self.__dict__.get("_jet_engine")
        if eng is None:
            eng = JetEngine(self, self.jet_program())
            self.__dict__["_jet_engine"] = eng
        return eng

    def forward(self, input_tensor: torch.Tensor) -> torch.Tensor:
        return jet.model_forward(self.jet_engine, input_tensor)

    # The kernels read packed copies of the weights; a trainer re-packs after its own updates, every other
    # in-place change of the parameters has to invalidate them.
    def load_state_dict(self, *args, **kwargs):
        res = super().load_state_dict(*args, **kwargs)
        eng = self.__dict__.get("_jet_engine")
        if eng is not None:
            eng.invalidate()
        return res

    def _apply(self, fn, *args, **kwargs):
        res = super()._apply(fn, *args, **kwargs)
        eng = self.__dict__.get("_jet_engine")
        if eng is not None:
            eng.invalidate()
        return res

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_jet_engine", None)  # device buffers are rebuilt lazily after unpickling
        return state
