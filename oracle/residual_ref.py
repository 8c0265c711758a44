"""ORACLE (test infrastructure, not product code) -- numpy interpreter of a traced residual program.

Restates what ``csrc/residual.cu`` computes for one condition: the residual components ``r_k[n]``
(reference: the list a residual callable returns, e.g. ``examples/poisson_2D_1C/problem.py:24-37``,
evaluated by ``Condition.get_residual``, ``multipinn/condition/condition.py:62-71``), their sums of
squares (reference ``multipinn/PINN/PINN.py:123``) and ``d/d(jets)`` of ``sum_k w_k sum_n r_k[n]^2``
(what autograd propagates from the loss back to the network outputs).  Instruction ``i`` defines
register ``i``; opcodes are those of ``include/mpinn_b200.h`` (MPINN_RP_*).

Pinned by ``tests/test_trace.py``: traced programs of all six golden configs reproduce the reference's
residual vectors / losses, and the seeds reproduce torch autograd of the eager residual.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import it.
"""
from __future__ import annotations

import numpy as np

CONST, LOADX, LOADJ, LOADD = 0, 1, 2, 3
ADD, SUB, MUL, DIV = 10, 11, 12, 13
NEG, SIN, COS, EXP, TANH, RELU, ABS, SQRT, LOG = 20, 21, 22, 23, 24, 25, 26, 27, 28
POWI, POWF = 30, 31
MAX, MIN = 14, 15
GT, GE, SEL = 40, 41, 42


def evaluate(instrs, consts, outputs, block_o, x, jets, weights=None, want_seeds=True, data=None, point_weights=None):
    """instrs: list of (op, a, b, c); jets: list of [J, N, o] arrays (one per model call); data: [N, n_data]
    per-point auxiliary columns read by LOADD (a ConditionExtra's data tuple, reference
    condition/condition.py:117-215), or None.
    Returns (values [K, N], sums [K], djets list or None)."""
    x = np.asarray(x)
    dt = x.dtype
    N = x.shape[0]
    v = [None] * len(instrs)
    for i, (op, a, b, c) in enumerate(instrs):
        if op == CONST:
            v[i] = np.full(N, consts[a], dtype=dt)
        elif op == LOADX:
            v[i] = x[:, a].copy()
        elif op == LOADJ:
            o = block_o[a]
            v[i] = np.asarray(jets[a])[b // o, :, b % o].astype(dt)
        elif op == LOADD:
            v[i] = np.asarray(data)[:, a].astype(dt)
        elif op == ADD:
            v[i] = v[a] + v[b]
        elif op == SUB:
            v[i] = v[a] - v[b]
        elif op == MUL:
            v[i] = v[a] * v[b]
        elif op == DIV:
            v[i] = v[a] / v[b]
        elif op == MAX:
            v[i] = np.maximum(v[a], v[b])
        elif op == MIN:
            v[i] = np.minimum(v[a], v[b])
        elif op == GT:
            v[i] = (v[a] > v[b]).astype(dt)
        elif op == GE:
            v[i] = (v[a] >= v[b]).astype(dt)
        elif op == SEL:      # torch.where(mask, x, y) / p[mask] = value
            v[i] = np.where(v[a] != 0, v[b], v[c])
        elif op == NEG:
            v[i] = -v[a]
        elif op == SIN:
            v[i] = np.sin(v[a])
        elif op == COS:
            v[i] = np.cos(v[a])
        elif op == EXP:
            v[i] = np.exp(v[a])
        elif op == TANH:
            v[i] = np.tanh(v[a])
        elif op == RELU:
            v[i] = np.maximum(v[a], 0)
        elif op == ABS:
            v[i] = np.abs(v[a])
        elif op == SQRT:
            v[i] = np.sqrt(v[a])
        elif op == LOG:
            v[i] = np.log(v[a])
        elif op == POWI:
            v[i] = v[a] ** int(b)
        elif op == POWF:
            v[i] = v[a] ** v[b]
        else:
            raise ValueError(f"bad opcode {op}")
    values = np.stack([v[r] for r in outputs])
    sums = (values.astype(np.float64) ** 2).sum(axis=1)
    if not want_seeds:
        return values, sums, None
    w = np.ones(len(outputs)) if weights is None else np.asarray(weights, dtype=np.float64)
    adj = [np.zeros(N, dtype=dt) for _ in instrs]
    for k, r in enumerate(outputs):     # point_weights [K, N]: lambda_k[n] of PointwiseLosses (loss = sum w_k lambda r^2)
        pw = 1.0 if point_weights is None else np.asarray(point_weights)[k].astype(dt)
        adj[r] = adj[r] + 2.0 * w[k] * pw * v[r]
    for i in range(len(instrs) - 1, -1, -1):
        op, a, b, c = instrs[i]
        g = adj[i]
        if op < ADD:
            continue
        if op == ADD:
            adj[a] = adj[a] + g
            adj[b] = adj[b] + g
        elif op == SUB:
            adj[a] = adj[a] + g
            adj[b] = adj[b] - g
        elif op == MUL:
            adj[a] = adj[a] + g * v[b]
            adj[b] = adj[b] + g * v[a]
        elif op == DIV:
            adj[a] = adj[a] + g / v[b]
            adj[b] = adj[b] - g * v[i] / v[b]
        elif op in (MAX, MIN):    # the gradient follows the selected operand (a on ties; torch splits ties evenly)
            pick_a = (v[a] >= v[b]) if op == MAX else (v[a] <= v[b])
            adj[a] = adj[a] + np.where(pick_a, g, 0)
            adj[b] = adj[b] + np.where(pick_a, 0, g)
        elif op in (GT, GE):
            pass                   # masks are piecewise constant
        elif op == SEL:
            adj[b] = adj[b] + np.where(v[a] != 0, g, 0)
            adj[c] = adj[c] + np.where(v[a] != 0, 0, g)
        elif op == NEG:
            adj[a] = adj[a] - g
        elif op == SIN:
            adj[a] = adj[a] + g * np.cos(v[a])
        elif op == COS:
            adj[a] = adj[a] - g * np.sin(v[a])
        elif op == EXP:
            adj[a] = adj[a] + g * v[i]
        elif op == TANH:
            adj[a] = adj[a] + g * (1 - v[i] ** 2)
        elif op == RELU:
            adj[a] = adj[a] + np.where(v[a] > 0, g, 0)
        elif op == ABS:
            adj[a] = adj[a] + g * np.sign(v[a])
        elif op == SQRT:
            adj[a] = adj[a] + g * 0.5 / v[i]
        elif op == LOG:
            adj[a] = adj[a] + g / v[a]
        elif op == POWI:
            adj[a] = adj[a] + g * int(b) * v[a] ** (int(b) - 1)
        elif op == POWF:
            adj[a] = adj[a] + g * v[b] * v[a] ** (v[b] - 1)
    djets = [np.zeros_like(np.asarray(j), dtype=dt) for j in jets]
    for i, (op, a, b, c) in enumerate(instrs):
        if op == LOADJ:
            o = block_o[a]
            djets[a][b // o, :, b % o] += adj[i]
    return values, sums, djets
