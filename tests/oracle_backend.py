"""TEST-ONLY kernel backend: implements the ``multipinn_b200.engine`` backend interface with the
numpy oracle (oracle/jet_ref.py) on CPU tensors, so the host-side logic (jet-block decoding,
plans, flat buffers, trainers, gloo data-parallel) can be exercised without a GPU.  Product code
never imports this."""
import numpy as np
import torch

from oracle import jet_ref as JR
from oracle import residual_ref as RR


class OracleBackend:
    name = "oracle"

    def __init__(self, dtype=np.float64):
        self.dtype = dtype
        self.calls = []

    def stash_bytes(self, ops, d, order, n):
        return 4

    def workspace_bytes(self, ops, n_params):
        return 4

    def packed_bytes(self, ops):
        return 4

    def pack_params(self, ops, params, packed):
        self.calls.append("pack")

    def _np(self, t):
        return t.detach().cpu().numpy().astype(self.dtype)

    def engine_supports(self, ops, d, order, engine):
        return True

    def jet_forward(self, ops, params, packed, x, order, out, stash, engine=0, lap=None):
        self.calls.append(("fwd", int(order), int(x.shape[0])))
        o, _ = JR.forward(ops, self._np(params), self._np(x), order, lap)
        out.copy_(torch.from_numpy(np.ascontiguousarray(o.transpose(1, 0, 2))).to(out.dtype))

    def jet_backward(self, ops, params, packed, x, order, dout, stash, dparams, workspace, engine=0, dx=None,
                     lap=None):
        self.calls.append(("bwd", int(order), int(x.shape[0])))
        p, xx = self._np(params), self._np(x)
        _, cache = JR.forward(ops, p, xx, order, lap)
        dp, dxx = JR.backward(ops, p, xx, order, cache, self._np(dout).transpose(1, 0, 2), lap)
        if dx is not None:
            dx.copy_(torch.from_numpy(np.ascontiguousarray(dxx)).to(dx.dtype))
        # frozen Fourier encodings get no gradient from the fused reverse pass
        ops = np.asarray(ops).reshape(-1, 8)
        if ops[0, 0] == JR.OP_FOURIER and not (ops[0, 6] & 4):
            dp[ops[0, 3]: ops[0, 3] + ops[0, 1] * ops[0, 2] // 2] = 0
        dparams.add_(torch.from_numpy(dp).to(dparams.dtype))

    def residual_program(self, prog, x, jets, djets=None, weights=None, sums=None, values=None,
                         point_major=False, jet_offset=0, jet_stride=0, data=None, point_weights=None, wsums=None):
        self.calls.append(("resid", int(x.shape[0])))
        if prog.data_width:
            assert data is not None and tuple(data.shape) == (x.shape[0], prog.data_width)
        prog.struct()   # exercises the limits / ctypes packing the CUDA backend relies on
        n = int(x.shape[0])
        if jet_stride:     # batched conditions / stacked blocks: this launch works on slices of longer jet blocks
            offs = [jet_offset] * len(jets) if isinstance(jet_offset, int) else list(jet_offset)
            jets = [j[:, o:o + n] for j, o in zip(jets, offs)]
            if djets is not None:
                djets = [None if t is None else t[:, o:o + n] for t, o in zip(djets, offs)]
        vals, sm, dj = RR.evaluate(prog.instrs, prog.consts, prog.outputs, [b.o for b in prog.blocks],
                                   self._np(x), [self._np(j) for j in jets],
                                   None if weights is None else self._np(weights), want_seeds=djets is not None,
                                   data=None if data is None else self._np(data),
                                   point_weights=None if point_weights is None else self._np(point_weights))
        if wsums is not None:
            pw = 1.0 if point_weights is None else self._np(point_weights).astype(np.float64)
            wsums.add_(torch.from_numpy((pw * vals.astype(np.float64) ** 2).sum(axis=1)).to(wsums.dtype))
        if sums is not None:
            sums.add_(torch.from_numpy(sm).to(sums.dtype))
        if values is not None:
            v = torch.from_numpy(np.ascontiguousarray(vals.T if point_major else vals)).to(values.dtype)
            values.copy_(v.reshape(values.shape))
        if djets is not None:
            for t, g in zip(djets, dj):
                if t is not None:
                    t.copy_(torch.from_numpy(g).to(t.dtype))

    def adam_tick(self, state, beta1, beta2):
        import math
        step = float(state[1]) + 1.0
        state[1] = step
        state[2] = float(state[0]) / (1.0 - float(np.float32(beta1)) ** step)
        state[3] = math.sqrt(1.0 - float(np.float32(beta2)) ** step)

    def adam_step_dev(self, p, g, m, v, state, beta1, beta2, eps, grad_scale=1.0):
        self.calls.append(("adam_dev", int(state[1])))
        JR.adam_step(p.numpy(), g.numpy(), m.numpy(), v.numpy(), int(state[1]), float(state[0]), beta1, beta2, eps,
                     grad_scale)

    def sumsq(self, r, accum):
        accum.add_(torch.sum(r.detach().double() ** 2).to(accum.dtype))

    def adam_step(self, p, g, m, v, step, lr, beta1, beta2, eps, grad_scale=1.0):
        self.calls.append(("adam", int(step)))
        JR.adam_step(p.numpy(), g.numpy(), m.numpy(), v.numpy(), step, lr, beta1, beta2, eps, grad_scale)
