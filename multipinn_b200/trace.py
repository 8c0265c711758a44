"""Residual tracing: the user's residual callable is run ONCE on symbolic per-point columns and recorded as a
small straight-line program over the jet channels of each ``model(arg)`` call, the point coordinates and
constants.  The device then evaluates the program, the sum of squares of every residual component and the
``dL/d(jet)`` seeds of the reverse pass in one launch (``csrc/residual.cu``), instead of ~100 eager tensor
kernels plus autograd bookkeeping per step.

What is traced is exactly the arithmetic the reference's examples perform between ``model(arg)`` and the list
of residuals they return (``examples/*/problem.py``): ``unpack`` / ``grad`` / ``Symbols`` (reference
``multipinn/condition/diff.py:4-47``, ``condition/symbols.py``), ``+ - * /``, integer and real powers, and the
elementwise ``torch`` functions in ``_UNARY``.  Anything else (data-dependent control flow, reductions over the
batch, unsupported torch calls, model calls on modified points) raises ``TraceError`` and the condition keeps
the eager jet path -- never a silently different result: a traced program is additionally checked against the
eager residual on probe points before it is used (``Condition.init_output_len``).

The optimiser pass ``laplace_rewrite`` recognises residuals that use second derivatives only through ONE fixed
linear combination ``sum_i c_i d2u/dx_i2`` (a Laplacian, or a single ``u_xx``) with constant coefficients.  Such a
block runs in the kernels' Laplace mode (``order = 3``): one second-order jet channel instead of ``d``.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------- IR
# opcodes -- MUST match include/mpinn_b200.h (MPINN_RP_*) and oracle/residual_ref.py
OP_CONST, OP_LOADX, OP_LOADJ, OP_LOADD = 0, 1, 2, 3
OP_ADD, OP_SUB, OP_MUL, OP_DIV = 10, 11, 12, 13
OP_NEG, OP_SIN, OP_COS, OP_EXP, OP_TANH, OP_RELU, OP_ABS, OP_SQRT, OP_LOG = 20, 21, 22, 23, 24, 25, 26, 27, 28
OP_POWI, OP_POWF = 30, 31
OP_MAX, OP_MIN = 14, 15                 # element-wise maximum / minimum
OP_GT, OP_GE, OP_SEL = 40, 41, 42       # a > b, a >= b as 1.0 / 0.0 ;  SEL m x y = x where m != 0 else y

MAX_INSTR, MAX_CONST, MAX_OUT, MAX_BLOCKS, MAX_JO, MAX_DATA = 192, 64, 8, 8, 32, 16
ORDER_LAPLACE = 3
ORDER_HESSIAN = 4        # diagonal AND mixed second derivatives: channels 2d+1.. are the pairs (0,1), (0,2), (1,2)
ENABLE_LAPLACE = True    # module switch for A/B measurements (bench.py --no-laplace); the rewrite is exact

_BINARY = {OP_ADD, OP_SUB, OP_MUL, OP_DIV}
_UNARY_OPS = {OP_NEG, OP_SIN, OP_COS, OP_EXP, OP_TANH, OP_RELU, OP_ABS, OP_SQRT, OP_LOG}
_REGS2 = _BINARY | {OP_MAX, OP_MIN, OP_GT, OP_GE, OP_POWF}    # operands a, b are registers
_REGS1 = _UNARY_OPS | {OP_POWI}                               # operand a is a register


def _map_operands(op, a, b, c, f):
    """(a, b, c) with every REGISTER operand passed through f."""
    if op in _REGS2:
        return f(a), f(b), c
    if op in _REGS1:
        return f(a), b, c
    if op == OP_SEL:
        return f(a), f(b), f(c)
    return a, b, c


def _register_operands(op, a, b, c):
    if op in _REGS2:
        return [a, b]
    if op in _REGS1:
        return [a]
    if op == OP_SEL:
        return [a, b, c]
    return []


class TraceError(Exception):
    """The residual cannot be expressed as a device program; the caller falls back to the eager jet path."""


def n_jets(d: int, order: int) -> int:
    if order == ORDER_LAPLACE:
        return d + 2
    if order == ORDER_HESSIAN:
        return 1 + 2 * d + d * (d - 1) // 2
    return 1 + order * d


def mixed_index(d: int, i: int, j: int) -> int:
    """Channel (full numbering) of d2/dx_i dx_j, i != j."""
    pairs = [(a, c) for a in range(d) for c in range(a + 1, d)]
    return 1 + 2 * d + pairs.index((min(i, j), max(i, j)))


class RInstr(ctypes.Structure):
    _fields_ = [("op", ctypes.c_int16), ("a", ctypes.c_int16), ("b", ctypes.c_int16), ("c", ctypes.c_int16)]


class RProg(ctypes.Structure):
    """Mirror of ``mpinn_rprog_t``."""
    _fields_ = [
        ("n_instr", ctypes.c_int32), ("n_const", ctypes.c_int32), ("n_out", ctypes.c_int32),
        ("n_blocks", ctypes.c_int32), ("d", ctypes.c_int32), ("n_data", ctypes.c_int32),
        ("reserved0", ctypes.c_int32), ("reserved1", ctypes.c_int32),
        ("block_J", ctypes.c_int32 * MAX_BLOCKS), ("block_o", ctypes.c_int32 * MAX_BLOCKS),
        ("out_reg", ctypes.c_int16 * MAX_OUT),
        ("jet_reg", (ctypes.c_int16 * MAX_JO) * MAX_BLOCKS),
        ("instr", RInstr * MAX_INSTR),
        ("consts", ctypes.c_float * MAX_CONST),
    ]


class Block:
    """One ``model(arg)`` call inside the residual: output width, derivative order actually used and, in Laplace
    mode, the coefficients ``c_i`` of its single second-order channel."""

    def __init__(self, o: int, engine=None, points: Optional[List[int]] = None):
        self.o = o
        self.engine = engine
        self.order = 0
        self.lap: Optional[Tuple[float, ...]] = None
        # registers holding the d coordinates the model was called with; None = the condition's points themselves
        # (model(arg)); a list for calls on derived points such as model(arg + 1)
        self.points = points


class ResidualProgram:
    """Finalised program: ``instrs`` is SSA (instruction i defines register i)."""

    def __init__(self, d: int, instrs: List[Tuple[int, int, int, int]], consts: List[float],
                 outputs: List[int], blocks: List[Block], data_width: int = 0):
        self.d = d
        self.data_width = data_width     # columns of the per-point data block LOADD reads (ConditionExtra)
        self.instrs = instrs
        self.consts = consts
        self.outputs = outputs
        self.blocks = blocks
        self._struct: Optional[RProg] = None

    # channel index of (block, jet j, output k) inside the [J, N, o] block
    def jets_of(self, b: int) -> int:
        return n_jets(self.d, self.blocks[b].order)

    def check_limits(self):
        if len(self.instrs) > MAX_INSTR:
            raise TraceError(f"residual program has {len(self.instrs)} instructions (limit {MAX_INSTR})")
        if len(self.consts) > MAX_CONST:
            raise TraceError(f"residual program has {len(self.consts)} constants (limit {MAX_CONST})")
        if len(self.outputs) > MAX_OUT:
            raise TraceError(f"residual has {len(self.outputs)} components (limit {MAX_OUT})")
        if self.data_width > MAX_DATA:
            raise TraceError(f"residual reads {self.data_width} per-point data columns (limit {MAX_DATA})")
        if len(self.blocks) > MAX_BLOCKS:
            raise TraceError(f"residual calls the model {len(self.blocks)} times (supported: 1..{MAX_BLOCKS})")
        for b, blk in enumerate(self.blocks):
            if self.jets_of(b) * blk.o > MAX_JO:
                raise TraceError("too many jet channels x outputs in one block")

    def struct(self) -> RProg:
        if self._struct is not None:
            return self._struct
        self.check_limits()
        s = RProg()
        s.n_instr, s.n_const, s.n_out, s.n_blocks = len(self.instrs), len(self.consts), len(self.outputs), len(self.blocks)
        s.d, s.n_data = self.d, self.data_width
        for b, blk in enumerate(self.blocks):
            s.block_J[b] = self.jets_of(b)
            s.block_o[b] = blk.o
            for i in range(MAX_JO):
                s.jet_reg[b][i] = -1
        for k, r in enumerate(self.outputs):
            s.out_reg[k] = r
        for i, (op, a, b, c) in enumerate(self.instrs):
            s.instr[i].op, s.instr[i].a, s.instr[i].b, s.instr[i].c = op, a, b, c
            if op == OP_LOADJ:
                s.jet_reg[a][b] = i
        for i, v in enumerate(self.consts):
            s.consts[i] = v
        self._struct = s
        return s

    def points_program(self, b: int) -> Optional["ResidualProgram"]:
        """Program (no model calls) computing the [N, d] points of block ``b`` from the condition's points, or None
        when the block is evaluated at the condition's points themselves."""
        regs = self.blocks[b].points
        if regs is None:
            return None
        sub = eliminate_dead(ResidualProgram(self.d, self.instrs, self.consts, list(regs), [], self.data_width))
        if any(op == OP_LOADJ for op, *_ in sub.instrs):
            raise TraceError("model called on points that depend on a model output")
        sub.check_limits()
        return sub

    def describe(self) -> str:
        names = {v: k[3:] for k, v in globals().items() if k.startswith("OP_")}
        lines = [f"d={self.d} blocks=" + ",".join(f"(o={b.o},order={b.order},lap={b.lap})" for b in self.blocks)]
        for i, (op, a, b, c) in enumerate(self.instrs):
            lines.append(f"  r{i} = {names.get(op, op)} {a} {b} {c}" + (f"  [{self.consts[a]}]" if op == OP_CONST else ""))
        lines.append("  out " + ", ".join(f"r{r}" for r in self.outputs))
        return "\n".join(lines)


# ---------------------------------------------------------------------------------------------- tracer
class _Tracer:
    def __init__(self, d: int, like: torch.Tensor):
        self.d = d
        self.like = like  # a real tensor: device / dtype answers for user code that asks
        self.instrs: List[Tuple[int, int, int, int]] = []
        self.consts: List[float] = []
        self._const_reg: Dict[float, int] = {}
        self._leaf_reg: Dict[tuple, int] = {}
        self.blocks: List[Block] = []
        # per block: requested derivative level of every (j, k) leaf is implied by j
        self.points: Optional["TMat"] = None
        self.n_data = 0

    def emit(self, op: int, a: int = 0, b: int = 0, c: int = 0) -> int:
        self.instrs.append((op, a, b, c))
        return len(self.instrs) - 1

    def const(self, v) -> "TVar":
        v = float(np.float32(v))
        if v != v:
            raise TraceError("NaN constant in residual")
        r = self._const_reg.get(v)
        if r is None:
            self.consts.append(v)
            r = self.emit(OP_CONST, len(self.consts) - 1)
            self._const_reg[v] = r
        return TVar(self, r)

    def leaf(self, key: tuple) -> "TVar":
        r = self._leaf_reg.get(key)
        if r is None:
            if key[0] == "x":
                r = self.emit(OP_LOADX, key[1])
            elif key[0] == "d":    # column of the per-point data block
                r = self.emit(OP_LOADD, key[1])
            else:  # ("j", block, j, k) with j in the FULL second-order numbering (0, 1..d, d+1..2d)
                r = self.emit(OP_LOADJ, key[1], key[2], key[3])
            self._leaf_reg[key] = r
        v = TVar(self, r)
        v.leaf = key
        return v

    def lift(self, v) -> "TVar":
        if isinstance(v, TVar):
            if v.tr is not self:
                raise TraceError("mixing values of two traces")
            return v
        if isinstance(v, _Poison):
            v.touch()
        if isinstance(v, (bool, np.bool_)):
            raise TraceError("boolean in residual arithmetic")
        if isinstance(v, (int, float, np.integer, np.floating)):
            return self.const(v)
        if isinstance(v, torch.Tensor):
            if v.numel() == 1:
                return self.const(float(v.detach().reshape(()).cpu()))
            raise TraceError("residual mixes traced columns with a real tensor (per-point data is not traced)")
        if isinstance(v, np.ndarray) and v.size == 1:
            return self.const(float(v.reshape(())))
        raise TraceError(f"unsupported operand of type {type(v).__name__} in residual")


class _Poison:
    """A mixed second derivative (or any channel the jet path does not carry).  Harmless while unused."""

    def __init__(self, why: str):
        self.why = why

    def touch(self):
        from .jet import UnsupportedDerivative
        raise UnsupportedDerivative(self.why)

    def _bad(self, *a, **k):
        self.touch()

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _bad
    __neg__ = __pow__ = __rpow__ = __abs__ = _bad

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        for a in args:
            if isinstance(a, _Poison):
                a.touch()
        raise TraceError("poisoned value")


def _int_exponent(e) -> Optional[int]:
    if isinstance(e, (bool, np.bool_)):
        return None
    if isinstance(e, (int, np.integer)):
        return int(e)
    if isinstance(e, (float, np.floating)) and float(e) == int(e) and abs(e) <= 64:
        return int(e)
    if isinstance(e, torch.Tensor) and e.numel() == 1:
        return _int_exponent(float(e))
    return None


class _Detached:
    """``u.detach()`` of a traced column.  Its VALUE cannot be used (a detached factor changes the gradient and is
    not recorded), but user code that only asks for its shape -- ``np.zeros_like(p.cpu().detach().numpy())``,
    reference examples/navier_stokes_3D_pipe_marching/problem.py:84 -- gets the one-row stand-in of the symbolic
    batch size, filled with NaN so that any real use shows up in the probe check."""

    def cpu(self): return self
    def numpy(self): return np.full((1,), np.nan, dtype=np.float32)

    def _bad(self, *a, **k):
        raise TraceError("detach() inside a residual changes its gradient; not traced")

    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = __truediv__ = __rtruediv__ = _bad
    __neg__ = __pow__ = __rpow__ = __abs__ = __getitem__ = _bad

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        raise TraceError("detach() inside a residual changes its gradient; not traced")


class TVar:
    """Symbolic per-point scalar column (what a ``[N]`` tensor is in eager mode)."""

    __slots__ = ("tr", "reg", "leaf")

    def __init__(self, tr: _Tracer, reg: int):
        self.tr = tr
        self.reg = reg
        self.leaf = None

    # -- arithmetic
    def _bin(self, op, other, swap=False):
        o = self.tr.lift(other)
        a, b = (o, self) if swap else (self, o)
        return TVar(self.tr, self.tr.emit(op, a.reg, b.reg))

    def __add__(self, o): return self._bin(OP_ADD, o)
    def __radd__(self, o): return self._bin(OP_ADD, o, True)
    def __sub__(self, o): return self._bin(OP_SUB, o)
    def __rsub__(self, o): return self._bin(OP_SUB, o, True)
    def __mul__(self, o): return self._bin(OP_MUL, o)
    def __rmul__(self, o): return self._bin(OP_MUL, o, True)
    def __truediv__(self, o): return self._bin(OP_DIV, o)
    def __rtruediv__(self, o): return self._bin(OP_DIV, o, True)
    def __neg__(self): return self._un(OP_NEG)
    def __pos__(self): return self
    def __abs__(self): return self._un(OP_ABS)

    def _un(self, op):
        return TVar(self.tr, self.tr.emit(op, self.reg))

    def __pow__(self, e):
        if isinstance(e, TVar):
            raise TraceError("power with a per-point exponent")
        k = _int_exponent(e)
        if k is not None:
            if k == 1:
                return self
            if -32768 <= k <= 32767:
                return TVar(self.tr, self.tr.emit(OP_POWI, self.reg, k))
        c = self.tr.lift(e)
        return TVar(self.tr, self.tr.emit(OP_POWF, self.reg, c.reg))

    def __rpow__(self, base):
        raise TraceError("constant ** per-point value")

    # data-dependent control flow cannot be recorded
    def __bool__(self):
        raise TraceError("residual branches on a per-point value")

    # comparisons give per-point masks (1.0 / 0.0 columns) for torch.where / masked assignment; using one as a
    # Python truth value is data-dependent control flow and still raises (__bool__)
    def __gt__(self, o): return self._bin(OP_GT, o)
    def __ge__(self, o): return self._bin(OP_GE, o)
    def __lt__(self, o): return self._bin(OP_GT, o, True)
    def __le__(self, o): return self._bin(OP_GE, o, True)
    __hash__ = object.__hash__

    def __eq__(self, o):
        raise TraceError("equality comparison of per-point values in residual")

    def __setitem__(self, mask, value):
        """``p[p > c] = v`` (reference examples/gas_dynamics_1D/problem.py:33-37): this handle now names the selected
        column, which is what the in-place write means for everything that uses ``p`` afterwards."""
        if not isinstance(mask, TVar):
            raise TraceError("only masked assignment with a traced mask is supported on a traced column")
        v = self.tr.lift(value)
        self.reg = self.tr.emit(OP_SEL, mask.reg, v.reg, self.reg)
        self.leaf = None

    # masks combine like booleans: ~m, m1 & m2, m1 | m2 (torch only defines these on bool tensors, i.e. on masks)
    def __invert__(self): return self.tr.const(1.0) - self
    def __and__(self, o): return self * o
    def __or__(self, o): return self._bin(OP_MAX, o)
    __rand__, __ror__ = __and__, __or__

    def maximum(self, o): return self._bin(OP_MAX, o)
    def minimum(self, o): return self._bin(OP_MIN, o)

    def clamp(self, min=None, max=None):
        r = self
        if min is not None:
            r = r._bin(OP_MAX, min)
        if max is not None:
            r = r._bin(OP_MIN, max)
        return r

    def __len__(self):
        return 1      # symbolic batch size, see TMat.shape

    @property
    def shape(self):
        return (1,)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def __getitem__(self, idx):
        raise TraceError("indexing a traced column")

    # -- tensor-method spellings user code uses
    def sin(self): return self._un(OP_SIN)
    def cos(self): return self._un(OP_COS)
    def exp(self): return self._un(OP_EXP)
    def tanh(self): return self._un(OP_TANH)
    def relu(self): return self._un(OP_RELU)
    def abs(self): return self._un(OP_ABS)
    def sqrt(self): return self._un(OP_SQRT)
    def log(self): return self._un(OP_LOG)
    def neg(self): return self._un(OP_NEG)
    def square(self): return self ** 2
    def pow(self, e): return self ** e
    def clone(self):
        c = TVar(self.tr, self.reg)      # a new handle: masked assignment on the clone must not reach the original
        c.leaf = self.leaf
        return c

    def cpu(self): return self
    def to(self, *a, **k): return self
    def float(self): return self
    def detach(self): return _Detached()
    def add(self, o): return self + o
    def sub(self, o): return self - o
    def mul(self, o): return self * o
    def div(self, o): return self / o

    @property
    def device(self): return self.tr.like.device

    @property
    def dtype(self): return self.tr.like.dtype

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _torch_function(func, args, kwargs or {})


class TMat:
    """Symbolic ``[N, C]`` matrix: the points, a model output or a ``grad`` result."""

    def __init__(self, tr: _Tracer, cols: list, kind: str, block: int = -1):
        self.tr = tr
        self.cols = list(cols)
        self.kind = kind      # "points" | "value" | "grad" | "derived"
        self.block = block

    # The batch size is symbolic.  It answers 1: every tensor user code builds from it (`torch.ones(arg.shape[0])`,
    # `torch.tensor([-1, 0]).repeat(arg.shape[0], 1)`) then has a leading dimension of one and broadcasts over the
    # points exactly as its N-row original would -- provided its rows are all equal, which is all that can be said
    # without data.  Anything else (aranges, random draws, a division by N) makes the traced program differ from
    # the eager residual at the probe points and is rejected there (`Condition._check_program`).
    @property
    def shape(self):
        return (1, len(self.cols))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self): return self.tr.like.device

    @property
    def dtype(self): return self.tr.like.dtype

    def __len__(self):
        return 1

    def __iter__(self):
        raise TraceError("iterating over the rows of a traced matrix")

    def _col_index(self, idx):
        if not (isinstance(idx, tuple) and len(idx) == 2 and idx[0] == slice(None)):
            raise TraceError("only [:, k] / [:, a:b] indexing of traced matrices is supported")
        return idx[1]

    def __getitem__(self, idx):
        k = self._col_index(idx)
        if isinstance(k, (int, np.integer)):
            c = self.cols[int(k)]
            if isinstance(c, _Poison):
                return c
            return c
        if isinstance(k, slice):
            return TMat(self.tr, self.cols[k], "derived", self.block)
        raise TraceError("unsupported column index")

    def __setitem__(self, idx, value):
        k = self._col_index(idx)
        if not isinstance(k, (int, np.integer)):
            raise TraceError("only single-column assignment is traced")
        self.cols[int(k)] = self.tr.lift(value)

    def clone(self):
        return TMat(self.tr, [c.clone() if isinstance(c, TVar) else c for c in self.cols], self.kind, self.block)

    # column-wise arithmetic (what `arg + torch.ones_like(arg)` or `arg * 2` mean for an [N, C] tensor)
    def _colwise(self, other, fn):
        if isinstance(other, TMat):
            if len(other.cols) != len(self.cols):
                raise TraceError("shape mismatch in traced matrix arithmetic")
            cols = [fn(a, b) for a, b in zip(self.cols, other.cols)]
        elif isinstance(other, torch.Tensor) and other.numel() == len(self.cols) and other.dim() <= 2:
            vals = other.detach().reshape(-1).cpu().tolist()   # a [C] / [1, C] row broadcast over the points
            cols = [fn(a, v) for a, v in zip(self.cols, vals)]
        else:
            cols = [fn(a, other) for a in self.cols]
        return TMat(self.tr, cols, "derived")

    def __add__(self, o): return self._colwise(o, lambda a, b: a + b)
    def __radd__(self, o): return self._colwise(o, lambda a, b: b + a)
    def __sub__(self, o): return self._colwise(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._colwise(o, lambda a, b: b - a)
    def __mul__(self, o): return self._colwise(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._colwise(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._colwise(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._colwise(o, lambda a, b: b / a)
    def __neg__(self): return TMat(self.tr, [-a for a in self.cols], "derived")

    def sum(self, *a, **k):
        return _torch_function(torch.sum, (self,) + a, k)

    def detach(self):
        raise TraceError("detach() of traced points")

    def to(self, *a, **k): return self
    def cpu(self): return self
    def float(self): return self

    def requires_grad_(self, *a, **k):
        return self

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _torch_function(func, args, kwargs or {})


_UNARY = {
    torch.sin: OP_SIN, torch.cos: OP_COS, torch.exp: OP_EXP, torch.tanh: OP_TANH, torch.relu: OP_RELU,
    torch.nn.functional.relu: OP_RELU, torch.abs: OP_ABS, torch.sqrt: OP_SQRT, torch.log: OP_LOG,
    torch.neg: OP_NEG, torch.negative: OP_NEG,
}
_BIN = {torch.add: OP_ADD, torch.sub: OP_SUB, torch.subtract: OP_SUB, torch.mul: OP_MUL, torch.multiply: OP_MUL,
        torch.div: OP_DIV, torch.true_divide: OP_DIV, torch.divide: OP_DIV,
        # `real_tensor <op> traced` arrives as the Tensor method (the real tensor is args[0])
        torch.Tensor.add: OP_ADD, torch.Tensor.sub: OP_SUB, torch.Tensor.mul: OP_MUL, torch.Tensor.div: OP_DIV,
        torch.Tensor.true_divide: OP_DIV, torch.Tensor.__add__: OP_ADD, torch.Tensor.__sub__: OP_SUB,
        torch.Tensor.__mul__: OP_MUL, torch.Tensor.__truediv__: OP_DIV}
_BIN_SWAPPED = {torch.Tensor.__radd__: OP_ADD, torch.Tensor.__rsub__: OP_SUB, torch.Tensor.__rmul__: OP_MUL,
                torch.Tensor.__rtruediv__: OP_DIV}


def _first_traced(args):
    for a in args:
        if isinstance(a, (TVar, TMat)):
            return a
        if isinstance(a, (list, tuple)):
            r = _first_traced(a)
            if r is not None:
                return r
    return None


def _torch_function(func, args, kwargs):
    t = _first_traced(args)
    if t is None:
        raise TraceError(f"torch.{getattr(func, '__name__', func)} without a traced operand")
    tr = t.tr
    if kwargs.get("inplace") is False:     # torch.nn.functional.relu(x) arrives with inplace=False
        kwargs = {k: v for k, v in kwargs.items() if k != "inplace"}
    if func in _UNARY and len(args) == 1 and not kwargs:
        return tr.lift(args[0])._un(_UNARY[func])
    if func is torch.Tensor.to and isinstance(args[0], torch.Tensor):   # real_tensor.to(traced): dtype / device of it
        return args[0].to(dtype=tr.like.dtype, device=tr.like.device)
    if func in _BIN_SWAPPED and len(args) == 2 and not kwargs:
        func, args = {OP_ADD: torch.add, OP_SUB: torch.sub, OP_MUL: torch.mul,
                      OP_DIV: torch.div}[_BIN_SWAPPED[func]], (args[1], args[0])
    if func in _BIN and len(args) == 2 and not kwargs and isinstance(args[1], TMat) and not isinstance(args[0], TMat):
        m = args[1]      # row / scalar <op> traced matrix
        return {OP_ADD: m.__radd__, OP_SUB: m.__rsub__, OP_MUL: m.__rmul__, OP_DIV: m.__rtruediv__}[_BIN[func]](args[0])
    if func in _BIN and len(args) == 2 and not kwargs and not isinstance(args[0], TMat):
        return tr.lift(args[0])._bin(_BIN[func], args[1])
    if func in (torch.sum, torch.Tensor.sum) and isinstance(args[0], TMat):
        dim = args[1] if len(args) > 1 else kwargs.get("dim", kwargs.get("axis"))
        if dim in (1, -1) and not kwargs.get("keepdim", False):    # sum over the columns: still one value per point
            cols = [tr.lift(c) for c in args[0].cols]
            acc = cols[0]
            for c in cols[1:]:
                acc = acc + c
            return acc
        raise TraceError("sum over the points of a traced matrix")
    if func in (torch.pow, torch.Tensor.pow, torch.Tensor.__pow__) and len(args) == 2:
        if not isinstance(args[0], TVar):
            raise TraceError("constant ** per-point value")
        return args[0] ** args[1]
    if func in (torch.maximum, torch.minimum) and len(args) == 2 and not kwargs:
        return tr.lift(args[0])._bin(OP_MAX if func is torch.maximum else OP_MIN, args[1])
    if func in (torch.clamp, torch.clip) and isinstance(args[0], TVar):
        lo = args[1] if len(args) > 1 else kwargs.get("min")
        hi = args[2] if len(args) > 2 else kwargs.get("max")
        return args[0].clamp(lo, hi)
    if func is torch.where and len(args) == 3 and not kwargs:
        if not isinstance(args[0], TVar):
            raise TraceError("torch.where with a mask that is not a traced comparison")
        x, y = tr.lift(args[1]), tr.lift(args[2])
        return TVar(tr, tr.emit(OP_SEL, args[0].reg, x.reg, y.reg))
    if func is torch.square and len(args) == 1:
        return tr.lift(args[0]) ** 2
    if func in (torch.zeros_like, torch.ones_like) and len(args) == 1 and not kwargs:
        c = tr.const(0.0 if func is torch.zeros_like else 1.0)
        if isinstance(args[0], TVar):
            return c
        return TMat(tr, [c] * len(args[0].cols), "derived")
    if func in _BIN and len(args) == 2 and not kwargs and isinstance(args[0], TMat):
        return {OP_ADD: args[0].__add__, OP_SUB: args[0].__sub__, OP_MUL: args[0].__mul__,
                OP_DIV: args[0].__truediv__}[_BIN[func]](args[1])
    if func is torch.unbind:
        m = args[0]
        dim = args[1] if len(args) > 1 else kwargs.get("dim", 0)
        if isinstance(m, TMat) and dim == 1:
            return tuple(m.cols)
    if func in (torch.stack, torch.cat, torch.concat):
        seq = args[0]
        dim = args[1] if len(args) > 1 else kwargs.get("dim", 0)
        if func is torch.stack and dim == 1 and all(isinstance(s, TVar) for s in seq):
            return TMat(tr, list(seq), "derived")
    raise TraceError(f"torch function {getattr(func, '__name__', func)} is not traced")


# ---------------------------------------------------------------------------------------------- model / grad hooks
def is_traced(x) -> bool:
    return isinstance(x, (TVar, TMat, _Poison))


def _point_regs(tr: _Tracer, arg: TMat) -> Optional[Tuple[int, ...]]:
    """Registers of the columns of a points matrix; None when they are the condition's own coordinates."""
    if len(arg.cols) != tr.d:
        raise TraceError(f"model called on a matrix with {len(arg.cols)} columns (expected {tr.d})")
    regs = []
    for c in arg.cols:
        if not isinstance(c, TVar):
            raise TraceError("model called on points with an unsupported column")
        regs.append(c.reg)
    if all(c.leaf == ("x", i) for i, c in enumerate(arg.cols)):
        return None
    return tuple(regs)


def traced_model_call(engine, arg: TMat) -> TMat:
    """``model(arg)`` on traced points: opens a jet block (or finds the one already opened for the same points)
    and returns its value matrix."""
    tr = arg.tr
    if tr.blocks and tr.blocks[0].engine is not engine:
        raise TraceError("residual calls two different models")
    regs = _point_regs(tr, arg)
    for b, blk in enumerate(tr.blocks):
        if (None if blk.points is None else tuple(blk.points)) == regs:
            break
    else:
        b = len(tr.blocks)
        tr.blocks.append(Block(engine.n_out, engine, None if regs is None else list(regs)))
    cols = [tr.leaf(("j", b, 0, k)) for k in range(engine.n_out)]
    return TMat(tr, cols, "value", b)


def _is_shift_of(tr: _Tracer, blk: Block, arg: TMat) -> bool:
    """``model(arg + c)`` differentiated with respect to ``arg`` (reference examples/portal_1D_1C/problem.py:36-39):
    the shift has an identity Jacobian, so the derivative is the block's own.  True when every coordinate the block
    was called with is ``arg[:, i] + const`` / ``arg[:, i] - const`` / ``const + arg[:, i]``."""
    if blk.points is None or len(arg.cols) != tr.d:
        return False
    for i, (r, c) in enumerate(zip(blk.points, arg.cols)):
        if not isinstance(c, TVar):
            return False
        op, a, b, _ = tr.instrs[r]
        is_const = lambda q: tr.instrs[q][0] == OP_CONST   # noqa: E731
        if op == OP_ADD and ((a == c.reg and is_const(b)) or (b == c.reg and is_const(a))):
            continue
        if op == OP_SUB and a == c.reg and is_const(b):
            continue
        return False
    return True


def traced_grad(u, arg) -> TMat:
    """``grad(u, arg)`` on traced values (reference condition/diff.py:25-47 semantics, served from jet channels)."""
    from .jet import UnsupportedDerivative
    if isinstance(u, _Poison):
        u.touch()
    if not isinstance(u, TVar) or not isinstance(arg, TMat):
        raise TraceError("grad() mixes traced and real tensors")
    tr = u.tr
    if u.leaf is None or u.leaf[0] != "j":
        raise UnsupportedDerivative(
            "grad(u, arg): u is not a view of a model output. Derivatives of derived expressions "
            "(e.g. grad(u*v, arg)) are not supported by the jet path")
    _, b, j, k = u.leaf
    d = tr.d
    blk = tr.blocks[b]
    try:
        same = _point_regs(tr, arg) == (None if blk.points is None else tuple(blk.points))
    except TraceError:
        same = False
    if not same and not _is_shift_of(tr, blk, arg):
        raise UnsupportedDerivative("grad(u, arg): arg is not the tensor the model was called with")
    if j == 0:
        blk.order = max(blk.order, 1)
        return TMat(tr, [tr.leaf(("j", b, 1 + i, k)) for i in range(d)], "grad", b)
    if j <= d:
        blk.order = max(blk.order, 2)
        if d == 1:
            return TMat(tr, [tr.leaf(("j", b, 2, k))], "grad", b)
        i0 = j - 1
        # off-diagonal columns are mixed second derivatives: real leaves; a block whose live program reads one runs
        # in the full-Hessian mode (order 4), otherwise they are dead code
        cols = [tr.leaf(("j", b, 1 + d + i0, k)) if i == i0 else tr.leaf(("j", b, mixed_index(d, i0, i), k))
                for i in range(d)]
        return TMat(tr, cols, "grad", b)
    raise UnsupportedDerivative("third and higher derivatives are not supported by the jet path")


def traced_unpack(m) -> tuple:
    if isinstance(m, TMat):
        return tuple(m.cols)
    raise TraceError("unpack() of a traced column")


# ---------------------------------------------------------------------------------------------- driver
def data_columns(item) -> int:
    """Columns one item of a ConditionExtra data tuple occupies in the packed [N, n_data] block."""
    if not isinstance(item, torch.Tensor) or item.dim() not in (1, 2) or not item.is_floating_point():
        raise TraceError("per-point data must be floating-point [N] or [N, c] tensors")
    return 1 if item.dim() == 1 else int(item.shape[1])


def pack_data(items, n: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The data tuple of a ConditionExtra as one contiguous float32 [N, n_data] block (the layout LOADD reads)."""
    cols = [t.detach().reshape(n, -1).to(torch.float32) for t in items]
    if out is None:
        return torch.cat(cols, dim=1).contiguous()
    torch.cat(cols, dim=1, out=out)
    return out


def trace_residual(function, model, d: int, like: torch.Tensor, data=None) -> ResidualProgram:
    """Run ``function(model, arg)`` -- ``function(model, arg, data)`` for a ConditionExtra, ``data`` being a sample
    of its per-point data tuple that fixes the column layout -- on symbolic points and return the optimised program."""
    tr = _Tracer(d, like)
    pts = TMat(tr, [tr.leaf(("x", i)) for i in range(d)], "points")
    tr.points = pts
    from . import jet
    extra = ()
    if data is not None:
        sym = []
        for item in data:
            c = data_columns(item)
            leaves = [tr.leaf(("d", tr.n_data + i)) for i in range(c)]
            tr.n_data += c
            sym.append(leaves[0] if item.dim() == 1 else TMat(tr, leaves, "data"))
        extra = (tuple(sym),)
    # a residual that draws random numbers (e.g. num_diff_random's per-point step, reference diff.py:84-111) cannot
    # be replayed as a fixed program: detected by the generator state moving while it is recorded
    import random as _pyrandom
    rng0 = torch.get_rng_state()
    rng0_cuda = torch.cuda.get_rng_state(like.device) if like.is_cuda else None
    rng0_np, rng0_py = np.random.get_state(), _pyrandom.getstate()
    with jet.use_plan(None):
        res = function(model, pts, *extra)
    rng1_np = np.random.get_state()
    if not torch.equal(rng0, torch.get_rng_state()) or \
            (rng0_cuda is not None and not torch.equal(rng0_cuda, torch.cuda.get_rng_state(like.device))) or \
            rng0_np[2:] != rng1_np[2:] or not np.array_equal(rng0_np[1], rng1_np[1]) or \
            rng0_py != _pyrandom.getstate():
        raise TraceError("residual draws random numbers")
    if isinstance(res, (TVar, TMat)) or not isinstance(res, (list, tuple)):
        raise TraceError("residual function must return a list of columns")
    outs = []
    for r in res:
        if isinstance(r, _Poison):
            r.touch()
        if not isinstance(r, TVar):
            if isinstance(r, (TMat, _Detached)):
                raise TraceError("residual component is not a traced per-point column")
            r = tr.lift(r)      # a constant component (a number or a one-element tensor broadcast over the points)
        outs.append(r.reg)
    if not tr.blocks:
        raise TraceError("residual never calls the model")
    prog = ResidualProgram(d, tr.instrs, tr.consts, outs, tr.blocks, tr.n_data)
    for b in range(len(tr.blocks)):
        prog.points_program(b)      # raises if a block's points depend on a model output
    prog = eliminate_dead(prog)        # (requested-but-unused mixed derivatives must not block the Laplace mode)
    prog = laplace_rewrite(prog)
    prog = simplify(prog)
    prog = eliminate_dead(prog)
    prog = renumber_jets(prog)
    prog.check_limits()
    return prog


def simplify(prog: ResidualProgram) -> ResidualProgram:
    """Peephole: x + 0, x - 0, x * 1, x / 1 -> x ;  x * 0 -> 0 ;  0 + x -> x ;  1 * x -> x.
    (IEEE caveat accepted: 0 * x is folded although x could be inf/nan; residual inputs are finite.)"""
    cval: Dict[int, float] = {}
    alias: Dict[int, int] = {}
    instrs = list(prog.instrs)

    def res(r):
        while r in alias:
            r = alias[r]
        return r

    for i, (op, a, b, c) in enumerate(instrs):
        if op == OP_CONST:
            cval[i] = prog.consts[a]
            continue
        a, b, c = _map_operands(op, a, b, c, res)
        instrs[i] = (op, a, b, c)
        ca, cb = cval.get(a) if op in _BINARY else None, cval.get(b) if op in _BINARY else None
        if op == OP_ADD:
            if cb == 0.0:
                alias[i] = a
            elif ca == 0.0:
                alias[i] = b
        elif op == OP_SUB:
            if cb == 0.0:
                alias[i] = a
        elif op == OP_MUL:
            if cb == 1.0:
                alias[i] = a
            elif ca == 1.0:
                alias[i] = b
            elif cb == 0.0:
                alias[i] = b
            elif ca == 0.0:
                alias[i] = a
        elif op == OP_DIV:
            if cb == 1.0:
                alias[i] = a
        if i in alias and alias[i] in cval:
            cval[i] = cval[alias[i]]
    for blk in prog.blocks:
        if blk.points is not None:
            blk.points = [res(r) for r in blk.points]
    return ResidualProgram(prog.d, instrs, prog.consts, [res(r) for r in prog.outputs], prog.blocks, prog.data_width)


def eliminate_dead(prog: ResidualProgram) -> ResidualProgram:
    live = [False] * len(prog.instrs)
    stack = list(prog.outputs)
    for blk in prog.blocks:
        if blk.points is not None:
            stack += list(blk.points)
    while stack:
        r = stack.pop()
        if live[r]:
            continue
        live[r] = True
        stack += _register_operands(*prog.instrs[r])
    remap, instrs, consts, cmap = {}, [], [], {}
    for i, (op, a, b, c) in enumerate(prog.instrs):
        if not live[i]:
            continue
        if op == OP_CONST:
            if a not in cmap:
                cmap[a] = len(consts)
                consts.append(prog.consts[a])
            a = cmap[a]
        else:
            a, b, c = _map_operands(op, a, b, c, remap.__getitem__)
        remap[i] = len(instrs)
        instrs.append((op, a, b, c))
    for blk in prog.blocks:
        if blk.points is not None:
            blk.points = [remap[r] for r in blk.points]
    return ResidualProgram(prog.d, instrs, consts, [remap[r] for r in prog.outputs], prog.blocks, prog.data_width)


def renumber_jets(prog: ResidualProgram) -> ResidualProgram:
    """LOADJ (block, j, k) in the full numbering -> (block, j' * o + k) with j' the channel index of the
    block's actual order.  The derivative order of a block is lowered to what the live program still reads."""
    d = prog.d
    used = [0] * len(prog.blocks)
    for op, a, b, c in prog.instrs:
        if op == OP_LOADJ:
            lvl = 0 if b == 0 else (1 if b <= d else (2 if b <= 2 * d else ORDER_HESSIAN))
            if prog.blocks[a].order == ORDER_LAPLACE and lvl == 2:
                lvl = ORDER_LAPLACE
            used[a] = max(used[a], lvl)
    for b, blk in enumerate(prog.blocks):
        blk.order = used[b]
        if blk.order != ORDER_LAPLACE:
            blk.lap = None
    instrs = []
    for op, a, b, c in prog.instrs:
        if op == OP_LOADJ:
            blk = prog.blocks[a]
            j = b
            if blk.order == ORDER_LAPLACE and j > d:
                j = d + 1   # the single second-order channel (laplace_rewrite left exactly one kind of read)
            if j >= n_jets(d, blk.order):
                raise TraceError("internal: jet channel beyond the block's order")
            b, c = j * blk.o + c, 0
        instrs.append((op, a, b, c))
    return ResidualProgram(prog.d, instrs, prog.consts, prog.outputs, prog.blocks, prog.data_width)


# ---------------------------------------------------------------------------------------------- Laplace mode
def laplace_rewrite(prog: ResidualProgram) -> ResidualProgram:
    """If every second-derivative leaf of a block is used only inside constant-coefficient linear combinations, and
    for every output k those combinations are multiples of ONE direction-weight vector c (the same for the whole
    block), replace them by reads of a single channel  L_k = sum_i c_i d2 u_k / dx_i2.

    Each register is analysed as  value = sum_{(k,i)} coef[(k,i)] * u_k,ii  +  rest  with python-float
    coefficients; any non-linear use of a register with a non-empty linear part disqualifies the block."""
    if not ENABLE_LAPLACE:
        return prog
    d = prog.d
    n = len(prog.instrs)
    nb = len(prog.blocks)
    lin: List[Optional[Dict[tuple, float]]] = [None] * n   # linear part in second-derivative leaves (per block key)
    cval: List[Optional[float]] = [None] * n               # value if the register is a compile-time constant
    bad = [False] * nb
    consumers_nonlinear = set()

    def blocks_of(m):
        return {key[0] for key in m} if m else set()

    for i, (op, a, b, c) in enumerate(prog.instrs):
        if op == OP_CONST:
            cval[i] = prog.consts[a]
        elif op == OP_LOADJ and b > 2 * d:
            bad[a] = True                         # a live mixed second derivative: full-Hessian mode, no Laplace channel
        elif op == OP_LOADJ and b > d:
            lin[i] = {(a, c, b - 1 - d): 1.0}    # (block, output k, direction)
        elif op in (OP_ADD, OP_SUB):
            la, lb = lin[a], lin[b]
            if la or lb:
                m = dict(la or {})
                s = 1.0 if op == OP_ADD else -1.0
                for key, v in (lb or {}).items():
                    m[key] = m.get(key, 0.0) + s * v
                lin[i] = m
            if cval[a] is not None and cval[b] is not None:
                cval[i] = cval[a] + cval[b] if op == OP_ADD else cval[a] - cval[b]
        elif op == OP_MUL:
            la, lb = lin[a], lin[b]
            if la and lb:
                for blk in blocks_of(la) | blocks_of(lb):
                    bad[blk] = True
            elif la or lb:
                l, other = (la, b) if la else (lb, a)
                if cval[other] is not None:
                    lin[i] = {key: v * cval[other] for key, v in l.items()}
                else:
                    for blk in blocks_of(l):
                        bad[blk] = True
            if cval[a] is not None and cval[b] is not None:
                cval[i] = cval[a] * cval[b]
        elif op == OP_DIV:
            if lin[b]:
                for blk in blocks_of(lin[b]):
                    bad[blk] = True
            if lin[a]:
                if cval[b] is not None and cval[b] != 0.0:
                    lin[i] = {key: v / cval[b] for key, v in lin[a].items()}
                else:
                    for blk in blocks_of(lin[a]):
                        bad[blk] = True
            if cval[a] is not None and cval[b] not in (None, 0.0):
                cval[i] = cval[a] / cval[b]
        elif op == OP_NEG:
            if lin[a]:
                lin[i] = {key: -v for key, v in lin[a].items()}
            if cval[a] is not None:
                cval[i] = -cval[a]
        elif op in _UNARY_OPS or op in (OP_POWI, OP_POWF, OP_MAX, OP_MIN, OP_GT, OP_GE, OP_SEL):
            for r in _register_operands(op, a, b, c):
                if lin[r]:
                    for blk in blocks_of(lin[r]):
                        bad[blk] = True

    # A linear part only matters where it is finally consumed: outputs, or non-linear consumers (already marked bad).
    # Collect, per block, the linear parts reaching the outputs.
    sinks: List[Dict[tuple, float]] = [lin[r] for r in prog.outputs if lin[r]]
    cand = [blk for blk in range(nb) if not bad[blk] and prog.blocks[blk].order == 2]
    if not cand or not sinks:
        return prog
    weights: Dict[int, Tuple[float, ...]] = {}
    for blk in cand:
        c_vec: Optional[np.ndarray] = None
        ok = True
        for m in sinks:
            per_k: Dict[int, np.ndarray] = {}
            for (bb, k, i), v in m.items():
                if bb == blk:
                    per_k.setdefault(k, np.zeros(d))[i] += v
            for k, vec in per_k.items():
                if not np.any(vec != 0.0):
                    continue
                if c_vec is None:
                    c_vec = vec / np.abs(vec).max() + 0.0
                    if c_vec[np.argmax(np.abs(c_vec))] < 0:
                        c_vec = -c_vec + 0.0
                scale = float(vec[np.argmax(np.abs(c_vec))] / c_vec[np.argmax(np.abs(c_vec))])
                if not np.allclose(vec, scale * c_vec, rtol=1e-12, atol=0.0):
                    ok = False
        if ok and c_vec is not None:
            weights[blk] = tuple(float(v) for v in c_vec)
    if not weights:
        return prog

    # Rewrite: every second-derivative leaf u_k,ii of a Laplace block becomes  w_i * L_k.  For a sink whose
    # coefficients are coef_i = alpha * c_i:  sum_i coef_i (w_i L) = alpha L sum_i c_i w_i, so any w with
    # sum_i c_i w_i = 1 reproduces every sink exactly; all the weight goes to the direction with the largest |c_i|
    # (w = e_i0 / c_i0), the other leaves become the constant 0 and fold away in simplify().  Intermediate registers
    # then hold partial sums that are not individually meaningful, but they are consumed linearly only (checked).
    instrs = list(prog.instrs)
    consts = list(prog.consts)
    out_map = list(range(len(instrs)))
    new_instrs: List[Tuple[int, int, int, int]] = []
    remap: Dict[int, int] = {}
    const_reg: Dict[float, int] = {}
    lap_leaf: Dict[tuple, int] = {}

    def emit(t):
        new_instrs.append(t)
        return len(new_instrs) - 1

    def const_of(v: float) -> int:
        v = float(np.float32(v))
        if v not in const_reg:
            consts.append(v)
            const_reg[v] = emit((OP_CONST, len(consts) - 1, 0, 0))
        return const_reg[v]

    for i, (op, a, b, c) in enumerate(instrs):
        if op == OP_LOADJ and b > d and a in weights:
            cv = np.asarray(weights[a])
            i0 = int(np.argmax(np.abs(cv)))
            w = 1.0 / cv[i0] if (b - 1 - d) == i0 else 0.0
            key = (a, c)
            if key not in lap_leaf:
                lap_leaf[key] = emit((OP_LOADJ, a, d + 1, c))   # full-numbering slot d+1 stands for the L channel
            if w == 0.0:
                remap[i] = const_of(0.0)
            elif w == 1.0:
                remap[i] = lap_leaf[key]
            else:
                remap[i] = emit((OP_MUL, const_of(w), lap_leaf[key], 0))
            continue
        a, b, c = _map_operands(op, a, b, c, remap.__getitem__)
        remap[i] = emit((op, a, b, c))
    for blk, cv in weights.items():
        prog.blocks[blk].order = ORDER_LAPLACE
        prog.blocks[blk].lap = cv
    for blk in prog.blocks:
        if blk.points is not None:
            blk.points = [remap[r] for r in blk.points]
    return ResidualProgram(prog.d, new_instrs, consts, [remap[r] for r in prog.outputs], prog.blocks, prog.data_width)
