"""ORACLE (test infrastructure, not product code) -- the reference's algorithm, as the reference runs it.

A compact restatement of multipinn's hot path in stock PyTorch: plain ``nn.Module`` MLPs,
nested ``torch.autograd.grad(..., create_graph=True)`` for the PDE derivatives, squared
residual means, ``loss.backward()`` and ``torch.optim.Adam``.  It is what the product is
compared with (same points, same weights) and what ``bench.py`` times as the CPU baseline.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import it.

Pinned against the unmodified reference (imported from /root/reference in the build
container) by ``tests/golden/make_golden.py``; see ``tests/test_oracle_golden.py``.

Reference lines restated
  * FNN            ``multipinn/neural_network/feedforward_neural_network.py:19-52``
  * activations    ``multipinn/neural_network/activation_function.py:5-54``
  * ResNet         ``multipinn/neural_network/residual_neural_network.py:15-126``
  * residual block ``multipinn/neural_network/light_residual_block.py:13-38``
  * Fourier        ``multipinn/neural_network/fourier_features.py:17-76``
  * grad / unpack  ``multipinn/condition/diff.py:4-47``
  * loss assembly  ``multipinn/PINN/PINN.py:50-125``
  * train step     ``multipinn/trainer/trainer.py:119-153,219-258,318-338``
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import jet_ref as J


# ------------------------------------------------------------------ activations
# The reference wraps both activations in @torch.jit.script (activation_function.py:5-26): on CUDA the TorchScript
# fuser then runs each activation -- and each level of its autodiff -- as one fused kernel, which is what makes the
# reference's own torch-CUDA step as fast as it is.  SCRIPTED selects that (default, the reference's real code
# path) or plain eager modules (bench.py reports both).
SCRIPTED = True


@torch.jit.script
def _scripted_gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / 1.41421))


@torch.jit.script
def _scripted_sin(x):
    return torch.sin(x)


class GELU(nn.Module):
    def forward(self, x):
        if SCRIPTED:
            return _scripted_gelu(x)
        return x * 0.5 * (1.0 + torch.erf(x / 1.41421))


class Sine(nn.Module):
    def forward(self, x):
        if SCRIPTED:
            return _scripted_sin(x)
        return torch.sin(x)


# ------------------------------------------------------------------ networks
class FNN(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_layers):
        super().__init__()
        hidden_layers = list(hidden_layers)
        mods = [nn.Sequential(nn.Linear(input_dim, hidden_layers[0]), Sine())]
        for a, b in zip(hidden_layers[:-1], hidden_layers[1:]):
            mods.append(nn.Sequential(nn.Linear(a, b), GELU()))
        mods.append(nn.Linear(hidden_layers[-1], output_dim))
        self.network_layers = nn.ModuleList(mods)

    def forward(self, x):
        for m in self.network_layers:
            x = m(x)
        return x


class LightResidualBlock(nn.Module):
    def __init__(self, activation, features):
        super().__init__()
        self.linear_first = nn.Sequential(nn.Linear(features, features), activation)
        self.linear_second = nn.Linear(features, features)
        self.activation = activation

    def forward(self, x):
        out = self.linear_second(self.linear_first(x))
        return self.activation(out + x)


class ResNet(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_layers, blocks=None):
        super().__init__()
        hidden_layers = list(hidden_layers)
        blocks = list(blocks) if blocks else [1] * len(hidden_layers)
        mods = [nn.Sequential(nn.Linear(input_dim, hidden_layers[0]), Sine())]

        def stage(count, fin, fout, tail):
            seq = [LightResidualBlock(GELU(), fin) for _ in range(count)]
            if tail:
                seq.append(nn.Sequential(nn.Linear(fin, fout), GELU()))
            return nn.Sequential(*seq)

        for a, k, b in zip(hidden_layers[:-1], blocks[:-1], hidden_layers[1:]):
            mods.append(stage(k, a, b, True))
        mods.append(stage(blocks[-1], hidden_layers[-1], hidden_layers[-1], False))
        mods.append(nn.Linear(hidden_layers[-1], output_dim))
        self.layers = nn.ModuleList(mods)

    def forward(self, x):
        for m in self.layers:
            x = m(x)
        return x


class FourierEncoding(nn.Module):
    def __init__(self, input_dim, encoding_dim, sigma, is_trainable=False):
        super().__init__()
        w = torch.zeros((input_dim, encoding_dim // 2))
        nn.init.normal_(w)
        self.B = nn.Parameter(w * sigma, requires_grad=is_trainable)

    def forward(self, x):
        return torch.cat([torch.sin(x @ self.B), torch.cos(x @ self.B)], dim=1)


class FourierFeatureNetwork(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_layers, encoding_dim, sigma=1.0,
                 is_trainable=False):
        super().__init__()
        self.fourier_encoding = FourierEncoding(input_dim, encoding_dim, sigma, is_trainable)
        self.fnn = FNN(encoding_dim, output_dim, hidden_layers)

    def forward(self, x):
        return self.fnn(self.fourier_encoding(x))


MODELS = {"FNN": FNN, "ResNet": ResNet, "FFNN": FourierFeatureNetwork}


# ------------------------------------------------------------------ derivative helpers
def unpack(batch):
    return torch.unbind(batch, dim=1)


def grad(u, arg):
    return torch.autograd.grad(u, arg, torch.ones_like(u), create_graph=True)[0]


# ------------------------------------------------------------------ program view of a model
def linear_layers(model) -> List[nn.Linear]:
    return [m for m in model.modules() if isinstance(m, nn.Linear)]


def build_program(model):
    """(program int array [n,8], flat fp64 params) in ``model.parameters()`` order.

    Works for any module built from the classes above *or* from same-named classes with the
    same attribute layout (the product's mirror classes), so tests can cross-check.
    """
    offs = {}
    off = 0
    for p in model.parameters():
        offs[id(p)] = off
        off += p.numel()
    ops = []

    def lin(l, act, flags=0):
        ops.append([J.OP_LINEAR, l.in_features, l.out_features, offs[id(l.weight)],
                    offs[id(l.bias)], act, flags, 0])

    def fnn_ops(fnn):
        mods = list(fnn.network_layers)
        lin(mods[0][0], J.ACT_SINE)
        for m in mods[1:-1]:
            lin(m[0], J.ACT_GELU)
        lin(mods[-1], J.ACT_NONE)

    name = type(model).__name__
    if name in ("FNN", "XavierFNN"):
        fnn_ops(model)
    elif name in ("FourierFeatureNetwork",):
        B = model.fourier_encoding.B
        ops.append([J.OP_FOURIER, B.shape[0], 2 * B.shape[1], offs[id(B)], 0, 0, 4 if B.requires_grad else 0, 0])
        fnn_ops(model.fnn)
    elif name == "ResNet":
        mods = list(model.layers)
        lin(mods[0][0], J.ACT_SINE)
        for stage in mods[1:-1]:
            for m in stage:
                if hasattr(m, "linear_first"):
                    lin(m.linear_first[0], J.ACT_GELU, J.FLAG_SAVE_SKIP)
                    lin(m.linear_second, J.ACT_GELU, J.FLAG_ADD_SKIP)
                else:
                    lin(m[0], J.ACT_GELU)
        lin(mods[-1], J.ACT_NONE)
    else:
        raise TypeError(name)
    flat = np.concatenate([p.detach().cpu().double().numpy().reshape(-1) for p in model.parameters()])
    return np.asarray(ops, dtype=np.int32), flat


# ------------------------------------------------------------------ loss + step (PINN.py / trainer.py)
def mean_loss(model, conditions: Sequence[Callable], points: Sequence[torch.Tensor]):
    """``calc_loss="mean"``: total = sum_k sum_n r_kn^2 / sum_k N_k ; losses_k = mean_n r_kn^2.

    conditions[i](model, points[i]) -> list of residual vectors (PINN.py:109-125).
    """
    sums, counts = [], []
    for fn, pts in zip(conditions, points):
        for r in fn(model, pts):
            sums.append(torch.sum(r**2))
            counts.append(len(r))
    total = torch.sum(torch.stack(sums)) / sum(counts)
    return total, torch.stack([s / c for s, c in zip(sums, counts)])


def legacy_loss(model, conditions, points):
    """``calc_loss="legacy"``: total = sum_k mean_n r_kn^2 (trainer.py:155-167)."""
    means = []
    for fn, pts in zip(conditions, points):
        for r in fn(model, pts):
            means.append(torch.mean(r**2))
    losses = torch.stack(means)
    return torch.sum(losses), losses


def train_step(model, optimizer, conditions, points, calc_loss=mean_loss):
    """One ``TrainerOneBatch._train_batches`` (trainer.py:318-338): returns (total, losses)."""
    optimizer.zero_grad()
    total, losses = calc_loss(model, conditions, points)
    total.backward()
    optimizer.step()
    return total.detach(), losses.detach()
