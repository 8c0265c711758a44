"""Host side of the jet engine: the CUDA backend (ctypes -> libmpinn_b200.so) and the per-model
``JetEngine`` that owns the flat parameter / gradient buffers and the packed kernel operands.

PyTorch is used for device memory, streams and autograd bookkeeping only; every number on the
hot path is produced by the kernels in ``multipinn_b200/csrc``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib

_backend = None


class CudaBackend:
    """Thin typed wrapper over the C ABI.  All tensors must be CUDA fp32 and contiguous."""

    name = "cuda"
    gpu_launches = 0  # kernels launched through this backend (reported by bench.py)
    timings = None    # bench.py sets this to a list to get (kind, n, order, start_evt, end_evt) per launch

    def _timed(self, kind, n, order):
        if CudaBackend.timings is None:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (kind, n, order, e0)

    def _timed_end(self, tok):
        if tok is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            CudaBackend.timings.append(tok + (e1,))

    def __init__(self):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("multipinn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _ops(ops: np.ndarray):
        return ops.ctypes.data_as(ctypes.c_void_p), int(ops.shape[0])

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        if t is None:
            return ctypes.c_void_p(0)
        assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
        return ctypes.c_void_p(t.data_ptr())

    def stash_bytes(self, ops, d, order, n):
        p, k = self._ops(ops)
        return int(self.lib.mpinn_stash_bytes(p, k, d, order, n))

    def workspace_bytes(self, ops, n_params):
        p, k = self._ops(ops)
        return int(self.lib.mpinn_backward_workspace_bytes(p, k, n_params))

    def packed_bytes(self, ops):
        p, k = self._ops(ops)
        return int(self.lib.mpinn_packed_bytes(p, k))

    def pack_params(self, ops, params, packed):
        p, k = self._ops(ops)
        _lib.check(self.lib.mpinn_pack_params(p, k, self._ptr(params), self._ptr(packed), self._stream()),
                   "mpinn_pack_params")
        CudaBackend.gpu_launches += 1

    @staticmethod
    def _lap(lap, d):
        """Host array of the d Laplace-mode channel weights (order 3), or NULL."""
        if lap is None:
            return None
        assert len(lap) == d
        return (ctypes.c_float * d)(*[float(v) for v in lap])

    def engine_supports(self, ops, d, order, engine):
        p, k = self._ops(ops)
        return bool(self.lib.mpinn_engine_supports(p, k, int(d), int(order), int(engine)))

    def jet_forward(self, ops, params, packed, x, order, out, stash, engine=_lib.ENGINE_AUTO, lap=None):
        p, k = self._ops(ops)
        n, d = x.shape
        tok = self._timed("fwd", n, order)
        lap_h = self._lap(lap, d)
        _lib.check(self.lib.mpinn_jet_forward(p, k, self._ptr(params), self._ptr(packed), self._ptr(x),
                                              n, d, order, lap_h, self._ptr(out), self._ptr(stash), engine,
                                              self._stream()), "mpinn_jet_forward")
        self._timed_end(tok)
        CudaBackend.gpu_launches += 1

    def jet_backward(self, ops, params, packed, x, order, dout, stash, dparams, workspace,
                     engine=_lib.ENGINE_AUTO, dx=None, lap=None):
        p, k = self._ops(ops)
        n, d = x.shape
        tok = self._timed("bwd", n, order)
        lap_h = self._lap(lap, d)
        _lib.check(self.lib.mpinn_jet_backward(p, k, self._ptr(params), self._ptr(packed), self._ptr(x),
                                               n, d, order, lap_h, self._ptr(dout), self._ptr(stash),
                                               self._ptr(dparams), dparams.numel(), self._ptr(dx),
                                               self._ptr(workspace),
                                               workspace.numel() * workspace.element_size(), engine,
                                               self._stream()), "mpinn_jet_backward")
        self._timed_end(tok)
        CudaBackend.gpu_launches += 2

    def residual_program(self, prog, x, jets, djets=None, weights=None, sums=None, values=None,
                         point_major=False, jet_offset=0, jet_stride=0, data=None, point_weights=None, wsums=None):
        """One launch: residual components of a traced program at the points ``x``; ``sums[k] += sum r_k^2``,
        ``djets[b]`` = seeds of the reverse pass for ``sum_k weights[k] sum_n r_k^2``, ``values[k, n] = r_k``
        (``values[n, k]`` with ``point_major``).  ``jet_offset`` / ``jet_stride``: the jet blocks hold ``jet_stride``
        points and this launch works on the slice starting at point ``jet_offset`` (conditions batched into one
        forward / reverse launch)."""
        st = prog.struct()
        nb = len(jets)
        if isinstance(jet_offset, int):
            jet_offset = [jet_offset] * nb
        offs = [o * b.o * 4 for o, b in zip(jet_offset, prog.blocks)]
        jp = (ctypes.c_void_p * nb)(*[t.data_ptr() + o for t, o in zip(jets, offs)]) if nb else None
        dp = None
        if djets is not None:
            dp = (ctypes.c_void_p * nb)(*[(t.data_ptr() + o if t is not None else 0) for t, o in zip(djets, offs)])
        for t in list(jets) + [t for t in (djets or []) if t is not None]:
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
        if point_weights is not None:
            assert tuple(point_weights.shape) == (len(prog.outputs), x.shape[0]) and point_weights.dtype == torch.float32
        if prog.data_width:
            assert data is not None and tuple(data.shape) == (x.shape[0], prog.data_width) \
                and data.dtype == torch.float32, "per-point data block must be float32 [N, n_data]"
        tok = self._timed("resid", x.shape[0], 0)
        _lib.check(self.lib.mpinn_residual_program(ctypes.byref(st), self._ptr(x),
                                                   self._ptr(data if prog.data_width else None), x.shape[0], jp, dp,
                                                   self._ptr(weights), self._ptr(sums), self._ptr(values),
                                                   1 if point_major else 0, int(jet_stride),
                                                   self._ptr(point_weights), self._ptr(wsums), self._stream()),
                   "mpinn_residual_program")
        self._timed_end(tok)
        CudaBackend.gpu_launches += 1

    def adam_tick(self, state, beta1, beta2):
        _lib.check(self.lib.mpinn_adam_tick(self._ptr(state), float(beta1), float(beta2), self._stream()),
                   "mpinn_adam_tick")
        CudaBackend.gpu_launches += 1

    def adam_step_dev(self, p, g, m, v, state, beta1, beta2, eps, grad_scale=1.0):
        _lib.check(self.lib.mpinn_adam_step_dev(self._ptr(p), self._ptr(g), self._ptr(m), self._ptr(v),
                                                p.numel(), self._ptr(state), float(beta1), float(beta2),
                                                float(eps), float(grad_scale), self._stream()),
                   "mpinn_adam_step_dev")
        CudaBackend.gpu_launches += 1

    def selftest_umma(self, mode, a, b, c):
        _lib.check(self.lib.mpinn_selftest_umma(int(mode), self._ptr(a), self._ptr(b), self._ptr(c),
                                                self._stream()), "mpinn_selftest_umma")

    def sumsq(self, r, accum):
        """accum (0-d / 1-elem fp32) += sum(r**2); r is a 1-D fp32 view with any stride."""
        assert r.dim() == 1 and r.is_cuda and r.dtype == torch.float32
        _lib.check(self.lib.mpinn_sumsq(ctypes.c_void_p(r.data_ptr()), r.shape[0],
                                        r.stride(0) if r.shape[0] > 1 else 1,
                                        ctypes.c_void_p(accum.data_ptr()), self._stream()), "mpinn_sumsq")
        CudaBackend.gpu_launches += 1

    def adam_step(self, p, g, m, v, step, lr, beta1, beta2, eps, grad_scale=1.0):
        _lib.check(self.lib.mpinn_adam_step(self._ptr(p), self._ptr(g), self._ptr(m), self._ptr(v),
                                            p.numel(), int(step), float(lr), float(beta1), float(beta2),
                                            float(eps), float(grad_scale), self._stream()),
                   "mpinn_adam_step")
        CudaBackend.gpu_launches += 1


def backend():
    """The process-wide backend.  Created on first use; raises if the CUDA path is unavailable."""
    global _backend
    if _backend is None:
        _backend = CudaBackend()
    return _backend


def _set_backend_for_tests(b):
    """Test hook: swap the kernel backend (used by the CPU test-suite to exercise host logic with
    the oracle standing in for the GPU).  Never called by product code."""
    global _backend
    old = _backend
    _backend = b
    return old


class JetEngine:
    """Per-model state: layer program, flat parameter/gradient buffers (the module's parameters are
    re-seated as views of them, so ``state_dict``/pickling/optimizers keep working), packed
    operands, and the forward/backward entry points."""

    COMM_TAIL = 128

    def __init__(self, model: torch.nn.Module, program: np.ndarray):
        self.model = model
        self.ops = _lib.ops_array(program)
        self.d = int(self.ops[0, 1])
        self.n_out = int(self.ops[-1, 2])
        self.kernel_engine = _lib.ENGINE_AUTO
        # opt-in: also produce dL/d(points) in the reverse pass, so that `points.grad` is populated like the
        # reference's backward() does (only its adversarial point generator reads it)
        self.points_grad = False
        self._packed_version = -1
        self.flat: Optional[torch.Tensor] = None
        self.flat_grad: Optional[torch.Tensor] = None
        self.packed: Optional[torch.Tensor] = None
        self.workspace: Optional[torch.Tensor] = None
        self.packed_valid = False
        self.managed = False  # a trainer promises to call invalidate() after each parameter update
        self.generation = 0   # bumped by invalidate(): jet blocks of older parameter values are never reused
        self._params = []
        self._offsets = []
        self._tc_ok = {}
        self.has_skip = bool(((self.ops[:, 6] & (_lib.FLAG_SAVE_SKIP | _lib.FLAG_ADD_SKIP)) != 0).any())

    # ---------------------------------------------------------------- flat buffers
    def _is_flat(self) -> bool:
        if self.flat is None:
            return False
        base = self.flat.data_ptr()
        esz = self.flat.element_size()
        params = list(self.model.parameters())
        if len(params) != len(self._params):
            return False
        for p, off in zip(params, self._offsets):
            if p.data_ptr() != base + off * esz or p.dtype != torch.float32:
                return False
        return True

    def ensure_flat(self):
        """(Re)build the flat buffers if the module's parameters moved (``.to(device)``, reload...)."""
        if self._is_flat():
            return
        params = list(self.model.parameters())
        if not params:
            raise RuntimeError("model has no parameters")
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32:
                raise TypeError("multipinn_b200 runs fp32 parameters only (the reference parity contract)")
            if p.device != dev:
                raise RuntimeError("all parameters must live on one device")
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        # the gradient buffer has a tail so the data-parallel sync can append the loss scalars and
        # send gradient + metrics as one message
        self._grad_storage = torch.zeros(total + self.COMM_TAIL, dtype=torch.float32, device=dev)
        grad = self._grad_storage[:total]
        offs, off = [], 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            offs.append(off)
            off += n
        flat.requires_grad_(any(p.requires_grad for p in params))
        flat.grad = grad if flat.requires_grad else None
        self.flat, self.flat_grad = flat, grad
        self._params, self._offsets = params, offs
        self.attach_grads()
        b = backend()
        self.packed = torch.empty(max(b.packed_bytes(self.ops), 4) // 4 + 64, dtype=torch.float32, device=dev)
        self.workspace = None
        self.packed_valid = False

    def comm_buffer(self, extra: int) -> torch.Tensor:
        """``[P + extra]`` view: the flat gradient followed by ``extra`` scratch floats."""
        if extra > self.COMM_TAIL:
            raise ValueError(f"at most {self.COMM_TAIL} loss components fit the comm tail")
        return self._grad_storage[: self.flat_grad.numel() + extra]

    def attach_grads(self):
        """Point every trainable parameter's ``.grad`` at its slice of the flat gradient buffer."""
        for p, off in zip(self._params, self._offsets):
            if p.requires_grad:
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
            else:
                p.grad = None   # frozen after the buffers were built: optimizers must not see a gradient

    def zero_grad(self):
        self.ensure_flat()
        self.flat_grad.zero_()
        if self.flat.requires_grad:
            self.flat.grad = self.flat_grad
        self.attach_grads()

    def invalidate(self):
        """Must be called after ANY in-place change of the parameters (optimizer step, ``load_state_dict``,
        manual ``p.data.copy_``): the kernels read packed copies of the weights."""
        self.packed_valid = False
        self.generation += 1

    def n_jets(self, order: int) -> int:
        if order == _lib.ORDER_LAPLACE:
            return self.d + 2
        if order == _lib.ORDER_HESSIAN:
            return 1 + 2 * self.d + self.d * (self.d - 1) // 2
        return 1 + order * self.d

    def tc_supported(self, order: int) -> bool:
        """Whether the tcgen05 engine runs this program at ``order`` (cached host-side query)."""
        ok = self._tc_ok.get(order)
        if ok is None:
            b = backend()
            ok = bool(b.engine_supports(self.ops, self.d, order, _lib.ENGINE_TCGEN05)) \
                if hasattr(b, "engine_supports") else False
            self._tc_ok[order] = ok
        return ok

    def _param_versions(self) -> int:
        """Sum of the parameters' autograd version counters: grows with every in-place write torch knows about
        (``reset_parameters()``, ``p.copy_()``, an optimizer step ...) -- not with writes through ``.data`` or with the
        raw-pointer updates of this package's own Adam kernel, after which the operands are re-packed anyway."""
        return sum(p._version for p in self._params)

    def _ensure_packed(self):
        ver = self._param_versions()
        if ver != self._packed_version and self.packed_valid and self.managed:
            self.generation += 1            # written behind the trainer's back (e.g. MarchingTrainer.reset_weights)
            self.packed_valid = False
        if not (self.packed_valid and self.managed):
            backend().pack_params(self.ops, self.flat.detach(), self.packed)
            self.packed_valid = True
        self._packed_version = ver

    # ---------------------------------------------------------------- kernels
    def _engine_id(self) -> int:
        """Engine both passes of a step run on.  dL/d(points) through a Fourier-feature first op exists on the
        FP32-FMA engine only (and the two engines' stash layouts differ, so the forward must follow)."""
        if self.points_grad and (int(self.ops[0, 0]) == _lib.OP_FOURIER or int(self.ops[0, 2]) == 128):
            return _lib.ENGINE_FFMA      # likewise for the layer-wise 128-wide tcgen05 engine (no dx output)
        return self.kernel_engine

    def stash_floats(self, order: int, n: int) -> int:
        return max(backend().stash_bytes(self.ops, self.d, order, n), 4) // 4

    def forward(self, x: torch.Tensor, order: int, need_stash: bool, lap=None, out=None, stash=None):
        """x [N,d] fp32 contiguous -> (out [J,N,o], stash or None).  ``order`` 3 is the Laplace mode with channel
        weights ``lap``.  ``out`` / ``stash`` may be caller-owned buffers (static buffers of a captured step)."""
        self.ensure_flat()
        self._ensure_packed()
        b = backend()
        n, d = x.shape
        if out is None:
            out = torch.empty((self.n_jets(order), n, self.n_out), dtype=torch.float32, device=x.device)
        if stash is None and (need_stash or (x.is_cuda and (order > 0 or self.has_skip)
                                             and self._engine_id() != _lib.ENGINE_FFMA and self.tc_supported(order))):
            # inference (order 0) needs no stash on either engine; value-only passes WITH derivative channels and
            # residual-block programs (their skip scratch lives behind it) stay on the tensor-core engine by
            # passing one
            stash = torch.empty(self.stash_floats(order, n), dtype=torch.float32, device=x.device)
        b.jet_forward(self.ops, self.flat.detach(), self.packed, x, order, out, stash, self._engine_id(),
                      **({"lap": lap} if lap is not None else {}))
        return out, stash

    def _workspace(self, device):
        if self.workspace is None or self.workspace.device != device:
            nbytes = backend().workspace_bytes(self.ops, self.flat.numel())
            self.workspace = torch.empty(nbytes // 4, dtype=torch.float32, device=device)
        return self.workspace

    def backward_into(self, x, order, dout, stash, dflat, lap=None, dx=None):
        """dflat [P] += parameter gradient for dL/d(out jets) = dout (one reverse launch + slab fold)."""
        backend().jet_backward(self.ops, self.flat.detach(), self.packed, x, order, dout, stash, dflat,
                               self._workspace(x.device), self._engine_id(), dx,
                               **({"lap": lap} if lap is not None else {}))

    def backward(self, x: torch.Tensor, order: int, dout: torch.Tensor, stash: torch.Tensor, want_dx: bool = False,
                 lap=None):
        """Returns the flat parameter gradient [P] for dL/d(out jets) = dout -- and, with ``want_dx``, also the
        gradient w.r.t. the points [N, d] (what the reference's ``backward()`` leaves in ``points.grad``)."""
        dflat = torch.zeros_like(self.flat_grad)
        dx = torch.empty_like(x) if want_dx else None
        self.backward_into(x, order, dout, stash, dflat, lap, dx)
        return (dflat, dx) if want_dx else dflat
