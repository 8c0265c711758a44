"""The fused train step: for every condition ONE forward-jet launch, ONE residual-program launch (residuals,
sum of squares, reverse-pass seeds) and ONE reverse launch that accumulates straight into the flat gradient;
then the loss bookkeeping, Adam with device-side scalars and the operand re-pack -- captured once into CUDA
graphs and replayed, so a step costs the host one or two graph launches instead of ~130 kernel launches and an
autograd graph walk.

Replaces, for conditions whose residual was traced (``trace.py``), the reference's
``Trainer._train_batches`` body: ``calc_loss`` -> ``PINN.calculate_mean_loss`` / ``calculate_loss`` ->
``batch_loss.backward()`` -> ``optimizer.step()`` (reference ``multipinn/trainer/trainer.py:119-153``,
``multipinn/PINN/PINN.py:91-125``).  Conditions that could not be traced keep the eager jet path; the trainer
then does not use this module at all.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib
from .engine import backend


class _BlockBuffers:
    __slots__ = ("order", "lap", "points_prog", "x", "out", "dout", "stash")


class ConditionRunner:
    """Static buffers and launches of one (condition, traced program) pair.  Every ``model(...)`` call of the
    residual is one *block*: a forward-jet launch before the residual program and a reverse launch after it."""

    def __init__(self, cond, prog, engine):
        self.cond = cond
        self.prog = prog
        self.engine = engine
        self.K = len(prog.outputs)
        self.n = -1
        self.x = None
        self.data = None            # [n, n_data] per-point data block of a ConditionExtra (LOADD)
        self._data_src = None
        # PointwiseLosses: lambda_k[n] (read), r_k[n] (written every step: the ascent direction r^2) and the slice of
        # the weighted-sum accumulators; all None otherwise
        self.lam = self.vals = self.wsums = None
        self.blocks: List[_BlockBuffers] = []
        for b, blk in enumerate(prog.blocks):
            bb = _BlockBuffers()
            bb.order = blk.order
            bb.lap = blk.lap if blk.order == _lib.ORDER_LAPLACE else None
            bb.points_prog = prog.points_program(b)
            bb.x = bb.out = bb.dout = bb.stash = None
            self.blocks.append(bb)
        # several model calls of the same jet order (model(arg) and model(arg + 1), a periodic boundary ...) are
        # STACKED: one forward and one reverse launch over the concatenated points of all calls, the residual
        # program reads call b at point offset b * n of the shared jet block
        keys = {(bb.order, tuple(bb.lap) if bb.lap is not None else None) for bb in self.blocks}
        self.stacked = len(self.blocks) > 1 and len(keys) == 1

    # -- a constant of the residual changed (closure_fingerprint): the re-traced program usually has the same shape
    def _shape(self, prog):
        return (len(prog.outputs), prog.data_width,
                tuple((b.o, b.order, None if b.lap is None else tuple(b.lap), b.points is None) for b in prog.blocks))

    def swap_program(self, prog) -> bool:
        """Take over a re-traced program in place (buffers, batching and policy state stay); False if its shape differs."""
        if self._shape(prog) != self._shape(self.prog):
            return False
        self.prog = prog
        for b, bb in enumerate(self.blocks):
            bb.points_prog = prog.points_program(b)
        return True

    # -- batching: conditions whose single model call has the same jet order share ONE forward and ONE reverse
    #    launch over their concatenated points (fewer tail waves and launches); each keeps its own residual program,
    #    which runs on its slice of the shared jet block
    def batch_key(self):
        if len(self.blocks) != 1 or self.blocks[0].points_prog is not None:
            return None
        bb = self.blocks[0]
        return (bb.order, tuple(bb.lap) if bb.lap is not None else None)

    def attach_shared(self, group, n0: int):
        self.group, self.n0 = group, n0
        self.n = group.counts[group.members.index(self)]
        self.x = group.x[n0:n0 + self.n]
        self._bind_data(self.n, group.x.device)

    def _bind_data(self, n: int, device):
        w = self.prog.data_width
        self.data = torch.empty((n, w), dtype=torch.float32, device=device) if w else None
        self._data_src = None

    def load_batch(self):
        """Copy the condition's current batch (points and, for a ConditionExtra, its data) into the static buffers."""
        c = self.cond
        self.x.copy_(c.batch_points.detach(), non_blocking=True)
        if self.data is not None and c.batch_data is not self._data_src:   # data change only with the points set
            from .trace import pack_data
            pack_data(c.batch_data, self.n, out=self.data)
            self._data_src = c.batch_data

    def launch_residual_shared(self, weights, sums):
        g = self.group
        backend().residual_program(self.prog, self.x, [g.out], [g.dout], weights, sums, self.vals,
                                   jet_offset=self.n0, jet_stride=g.n_total, data=self.data,
                                   point_weights=self.lam, wsums=self.wsums)

    def bind(self, n: int, device) -> bool:
        """(Re)allocate the static buffers for n points; True if they changed (a captured graph is then stale)."""
        if n == self.n and self.x is not None and self.x.device == device:
            return False
        eng = self.engine
        self.n = n
        self._bind_data(n, device)
        if self.stacked:
            B, bb0 = len(self.blocks), self.blocks[0]
            self.xcat = torch.empty((B * n, eng.d), dtype=torch.float32, device=device)
            self.out_cat = torch.empty((eng.n_jets(bb0.order), B * n, eng.n_out), dtype=torch.float32, device=device)
            self.dout_cat = torch.empty_like(self.out_cat)
            self.stash_cat = torch.empty(eng.stash_floats(bb0.order, B * n), dtype=torch.float32, device=device)
            ident = [b for b, bb in enumerate(self.blocks) if bb.points_prog is None]
            for b, bb in enumerate(self.blocks):
                bb.x = self.xcat[b * n:(b + 1) * n]
            # the condition's own points are copied straight into the slice of the call that uses them unchanged
            self.x = self.blocks[ident[0]].x if ident else torch.empty((n, eng.d), dtype=torch.float32, device=device)
            return True
        self.x = torch.empty((n, eng.d), dtype=torch.float32, device=device)
        for bb in self.blocks:
            bb.x = self.x if bb.points_prog is None else torch.empty_like(self.x)
            bb.out = torch.empty((eng.n_jets(bb.order), n, eng.n_out), dtype=torch.float32, device=device)
            bb.dout = torch.empty_like(bb.out)
            bb.stash = torch.empty(eng.stash_floats(bb.order, n), dtype=torch.float32, device=device)
        return True

    def _forward_blocks(self, x, blocks, need_stash, data=None):
        eng, be = self.engine, backend()
        for bb in blocks:
            if bb.points_prog is not None:   # e.g. model(arg + 1): the shifted points are produced on the device
                be.residual_program(bb.points_prog, x, [], None, None, None, bb.x, point_major=True, data=data)
            eng.forward(bb.x, bb.order, need_stash, lap=bb.lap, out=bb.out, stash=bb.stash)

    def launch(self, weights, sums, dflat):
        eng = self.engine
        if self.stacked:
            be, bb0, n, B = backend(), self.blocks[0], self.n, len(self.blocks)
            for bb in self.blocks:
                if bb.points_prog is not None:
                    be.residual_program(bb.points_prog, self.x, [], None, None, None, bb.x, point_major=True,
                                        data=self.data)
            eng.forward(self.xcat, bb0.order, True, lap=bb0.lap, out=self.out_cat, stash=self.stash_cat)
            be.residual_program(self.prog, self.x, [self.out_cat] * B, [self.dout_cat] * B, weights, sums, self.vals,
                                jet_offset=[b * n for b in range(B)], jet_stride=B * n, data=self.data,
                                point_weights=self.lam, wsums=self.wsums)
            eng.backward_into(self.xcat, bb0.order, self.dout_cat, self.stash_cat, dflat, lap=bb0.lap)
            return
        self._forward_blocks(self.x, self.blocks, True, self.data)
        backend().residual_program(self.prog, self.x, [bb.out for bb in self.blocks],
                                   [bb.dout for bb in self.blocks], weights, sums, self.vals, data=self.data,
                                   point_weights=self.lam, wsums=self.wsums)
        for bb in self.blocks:
            eng.backward_into(bb.x, bb.order, bb.dout, bb.stash, dflat, lap=bb.lap)

    def residual_values(self, x: torch.Tensor, data=None) -> torch.Tensor:
        """[K, N] residual components at arbitrary points, value-only (no seeds, no reverse pass): the fast path
        for residual-driven samplers and callbacks.  ``data``: the packed [N, n_data] block of a ConditionExtra
        (default: its ``data_gen`` evaluated at ``x``)."""
        eng = self.engine
        n = x.shape[0]
        if self.prog.data_width and data is None:
            from .trace import pack_data
            data = pack_data(self.cond.data_gen(x), n).to(x.device)
        tmp = []
        for bb in self.blocks:
            t = _BlockBuffers()
            t.order, t.lap, t.points_prog = bb.order, bb.lap, bb.points_prog
            t.x = x if bb.points_prog is None else torch.empty_like(x)
            t.out = torch.empty((eng.n_jets(bb.order), n, eng.n_out), dtype=torch.float32, device=x.device)
            t.stash = None
            tmp.append(t)
        self._forward_blocks(x, tmp, False, data)
        vals = torch.empty((self.K, n), dtype=torch.float32, device=x.device)
        backend().residual_program(self.prog, x, [t.out for t in tmp], None, None, None, vals, data=data)
        return vals


class BatchedGroup:
    """Conditions with the same jet order evaluated by one forward and one reverse launch."""

    def __init__(self, members: List[ConditionRunner]):
        self.members = members
        self.engine = members[0].engine
        self.order = members[0].blocks[0].order
        self.lap = members[0].blocks[0].lap
        self.counts: List[int] = []
        self.n_total = -1
        self.x = self.out = self.dout = self.stash = None

    def bind(self, device) -> bool:
        counts = [int(m.cond.batch_points.shape[0]) for m in self.members]
        if counts == self.counts and self.x is not None and self.x.device == device:
            return False
        eng = self.engine
        self.counts, self.n_total = counts, sum(counts)
        n = self.n_total
        self.x = torch.empty((n, eng.d), dtype=torch.float32, device=device)
        self.out = torch.empty((eng.n_jets(self.order), n, eng.n_out), dtype=torch.float32, device=device)
        self.dout = torch.empty_like(self.out)
        self.stash = torch.empty(eng.stash_floats(self.order, n), dtype=torch.float32, device=device)
        n0 = 0
        for m, c in zip(self.members, counts):
            m.attach_shared(self, n0)
            n0 += c
        return True

    def launch(self, weights, sums, dflat, k0s):
        eng = self.engine
        eng.forward(self.x, self.order, True, lap=self.lap, out=self.out, stash=self.stash)
        for m, k0 in zip(self.members, k0s):
            m.launch_residual_shared(weights[k0:k0 + m.K], sums[k0:k0 + m.K])
        eng.backward_into(self.x, self.order, self.dout, self.stash, dflat, lap=self.lap)


class FusedStep:
    """Whole-step executor for a trainer whose conditions are all traced."""

    def __init__(self, engine, conditions, programs, mode: str, flat_adam=None, use_graphs: bool = True, reg=None):
        self.engine = engine
        # loss policy: "mean" | "legacy" (the reference's two built-in ones) | "constant" (ConstantLosses) |
        # "gradient" / "normal" (GradientLosses / NormalLosses: statistics of K per-condition gradients)
        self.mode = mode
        self.reg = reg
        self.per_condition = mode in ("gradient", "normal")
        self.trainer = None           # set by build(): the pointwise policy follows the trainer's epoch counter
        self.flat_adam = flat_adam
        self.runners: List[ConditionRunner] = [ConditionRunner(c, p, engine) for c, p in zip(conditions, programs)]
        self.K = sum(r.K for r in self.runners)
        self.k0 = []
        k = 0
        for r in self.runners:
            self.k0.append(k)
            k += r.K
        # launch units, in condition order of their first member: single conditions or batched groups
        by_key, self.units = {}, []
        for r in self.runners:
            key = None if self.per_condition else r.batch_key()   # per-condition gradients: one reverse launch each
            if key is None:
                self.units.append(r)
            elif key in by_key:
                by_key[key].append(r)
            else:
                by_key[key] = [r]
                self.units.append(by_key[key])
        self.units = [(u if isinstance(u, ConditionRunner) else (BatchedGroup(u) if len(u) > 1 else u[0]))
                      for u in self.units]
        self.use_graphs = use_graphs
        self._g_compute: Optional[torch.cuda.CUDAGraph] = None
        self._g_compute_acc: Optional[torch.cuda.CUDAGraph] = None     # same, adding to the gradient (batches > 0)
        self._g_update: Optional[torch.cuda.CUDAGraph] = None
        self._calls = {"_g_compute": 0, "_g_compute_acc": 0, "_g_update": 0}
        self._dev = None
        self._lr_on_device = None
        self._scale = 1.0
        self.graph_replays = 0
        self._launches = {"_g_compute": 0, "_g_compute_acc": 0, "_g_update": 0}   # kernels one direct run of each part launches

    @staticmethod
    def build(trainer) -> Optional["FusedStep"]:
        """A FusedStep for ``trainer`` or None when any precondition fails (the trainer stays on the eager path)."""
        eng = trainer._engine
        if eng is None or eng.points_grad or trainer.num_batches < 1:
            return None
        multi = trainer.num_batches > 1 and not hasattr(trainer, "rank")     # (TrainerMultiGPU: one batch per rank)
        from .regularization import (AdaptiveConditionsLosses, ConstantLosses, GradientLosses, NormalLosses,
                                     PointwiseLosses)
        from .trainer.trainer import Trainer
        reg = None
        if trainer.calc_loss == Trainer.mean_calc_loss:
            mode = "mean"
        elif trainer.calc_loss == Trainer.sum_of_means_calc_loss:
            mode = "legacy"
        elif type(trainer.calc_loss) in (ConstantLosses, GradientLosses, NormalLosses):
            reg = trainer.calc_loss
            mode = {ConstantLosses: "constant", GradientLosses: "gradient", NormalLosses: "normal"}[type(reg)]
        elif type(trainer.calc_loss) is AdaptiveConditionsLosses and (
                trainer.calc_loss.adaptive_cond is None or hasattr(trainer.calc_loss.adaptive_cond, "_fused_state")):
            reg, mode = trainer.calc_loss, "adaptive"
        elif type(trainer.calc_loss) is PointwiseLosses and trainer.calc_loss.optimizer_def is torch.optim.Adam \
                and trainer.num_batches == 1 and not trainer.calc_loss.models:
            reg, mode = trainer.calc_loss, "pointwise"
        else:
            return None
        if multi and mode not in ("mean", "legacy", "constant"):
            return None        # per-condition gradient policies and the weight-training ones: one batch only
        progs = []
        for c in trainer.pinn.conditions:
            p = c.program_for(trainer.pinn.model) if hasattr(c, "program_for") else None
            if p is None:
                return None
            progs.append(p)
        K = sum(len(p.outputs) for p in progs)
        if K + 1 > eng.COMM_TAIL:
            return None
        # a ConditionExtra hands its residual the data of ALL its points whatever the batch (reference
        # condition/condition.py:186-188 slices the tuple, not the tensors): only consistent with one batch
        if any(p.data_width for p in progs) and trainer.num_batches != 1:
            return None
        # the regularizers raise the reference's own exceptions for shapes they do not support: leave those to them
        if mode == "constant" and len(reg.lambdas) != K:
            return None
        if mode in ("gradient", "normal") and (K != len(progs) or K < 2):
            return None
        fs = FusedStep(eng, trainer.pinn.conditions, progs, mode, trainer._flat_adam,
                       use_graphs=getattr(trainer, "use_cuda_graphs", True), reg=reg)
        import weakref
        fs.trainer = weakref.ref(trainer)
        return fs

    # ------------------------------------------------------------------ buffers
    def _bind(self, device):
        eng = self.engine
        eng.ensure_flat()
        changed = device != self._dev
        for u in self.units:
            if isinstance(u, BatchedGroup):
                changed |= u.bind(device)
            else:
                changed |= u.bind(int(u.cond.batch_points.shape[0]), device)
        if changed:
            self._dev = device
            self._g_compute = self._g_compute_acc = self._g_update = None
            self._calls = {"_g_compute": 0, "_g_compute_acc": 0, "_g_update": 0}
            eng.attach_grads()
            K = self.K
            counts = []
            for r in self.runners:
                counts += [float(r.n)] * r.K
            total = float(sum(r.n * r.K for r in self.runners))
            self.sums = torch.zeros(K, dtype=torch.float32, device=device)
            self.inv_counts = torch.tensor([1.0 / c for c in counts], dtype=torch.float32, device=device)
            self.inv_total = 1.0 / total
            if self.mode == "mean":      # total = sum_k sum_n r^2 / sum_k N_k   (reference PINN.py:109-125)
                self.weights = torch.full((K,), 1.0 / total, dtype=torch.float32, device=device)
            elif self.mode == "constant":  # total = sum_k lambda_k mean_n r^2     (reference regularization/constant.py)
                self.lambdas = torch.tensor([float(v) for v in self.reg.lambdas], dtype=torch.float32, device=device)
                self.weights = self.inv_counts * self.lambdas
            elif self.mode == "adaptive":    # total = sum_k softmax(theta)_k mean_n r_kn^2  (reference adaptive.py:7-153)
                from .regularization.adaptive import AdaptiveWeightSum
                self.weights = torch.zeros(K, dtype=torch.float32, device=device)
                aw = self.reg.adaptive_cond
                st = getattr(aw, "_fused_state", None)
                if st is None or st["m"].numel() != K or st["m"].device != device:
                    aw = AdaptiveWeightSum(counts, self.reg.lr).to(device)  # theta_0 = log(N_k), as the reference
                    st = {"gtheta": torch.zeros(K, dtype=torch.float32, device=device),
                          "m": torch.zeros(K, dtype=torch.float32, device=device),
                          "v": torch.zeros(K, dtype=torch.float32, device=device),
                          "state": torch.zeros(4, dtype=torch.float32, device=device), "calls": 0}
                    st["state"][0] = float(self.reg.lr)
                    aw._fused_state = st      # lives with the policy object: survives a rebuild of this executor
                    self.reg.adaptive_cond = aw
                self.ad_theta = aw.train_weight.data                        # the module's parameter IS the buffer
                self.ad_w = torch.softmax(self.ad_theta, dim=0)
                aw.weight = self.ad_w
                self.ad_gtheta, self.ad_m, self.ad_v, self.ad_state = st["gtheta"], st["m"], st["v"], st["state"]
                self.ad_g = torch.zeros(K, dtype=torch.float32, device=device)
                self._ad_st = st
            elif self.mode == "pointwise":   # total = sum_k sum_n lambda_kn r_kn^2  (reference pointwise.py:84-100)
                self.weights = torch.ones(K, dtype=torch.float32, device=device)
                self._pw_allocate(device)
            else:                        # total = sum_k mean_n r^2                (reference trainer.py:155-167)
                self.weights = self.inv_counts.clone()  # (gradient / normal: per-condition means, combined below)
            if self.per_condition:
                P = eng.flat_grad.numel()
                self.G = torch.zeros((K, P), dtype=torch.float32, device=device)   # one gradient slab per condition
                if self.reg.lambda_regularization is None or self.reg.lambda_regularization.device != device:
                    self.reg.lambda_regularization = torch.ones(K - 1, dtype=torch.float32, device=device)
                if self.mode == "gradient":
                    # w_i = 1 / (T * numel(tensor(i))): sum_i |g_i| w_i == mean over tensors of mean |g|
                    w = torch.zeros(P, dtype=torch.float32, device=device)
                    trainable = [(p, off) for p, off in zip(eng._params, eng._offsets) if p.requires_grad]
                    for p, off in trainable:
                        w[off:off + p.numel()] = 1.0 / (len(trainable) * p.numel())
                    self.tensor_mean_w = w
                else:
                    last = list(list(eng.model.children())[-1].children())[-1]
                    off = next(o for p, o in zip(eng._params, eng._offsets) if p is last.weight)
                    self.last_slice = slice(off, off + last.weight.numel())
            if self.flat_adam is not None:
                self.adam_state = torch.zeros(4, dtype=torch.float32, device=device)
                self.adam_state[1] = float(self.flat_adam.step)
                self._step_on_device = self.flat_adam.step
                self._lr_on_device = None

    def _loss_view(self):
        eng = self.engine
        P = eng.flat_grad.numel()
        return eng.comm_buffer(self.K + 1)[P:P + self.K + 1]   # [K component means, total]

    # ------------------------------------------------------------------ launches
    # ------------------------------------------------------------------ PointwiseLosses (per-point loss weights)
    def _pw_allocate(self, device):
        """Flat buffers over all (component, point) pairs: lambda, the residual values of the last step, Adam moments
        and the ascent direction; every runner works on its [K_r, n_r] views."""
        total = sum(r.K * r.n for r in self.runners)
        self.pw_lam = torch.ones(total, dtype=torch.float32, device=device)
        self.pw_vals = torch.zeros(total, dtype=torch.float32, device=device)
        self.pw_g = torch.zeros(total, dtype=torch.float32, device=device)
        self.pw_m = torch.zeros(total, dtype=torch.float32, device=device)
        self.pw_v = torch.zeros(total, dtype=torch.float32, device=device)
        self.pw_train = torch.ones(total, dtype=torch.float32, device=device)    # 0 on ignored components
        self.pw_state = torch.zeros(4, dtype=torch.float32, device=device)
        self.wsums = torch.zeros(self.K, dtype=torch.float32, device=device)
        off = 0
        for r, k0 in zip(self.runners, self.k0):
            m = r.K * r.n
            r.lam = self.pw_lam[off:off + m].view(r.K, r.n)
            r.vals = self.pw_vals[off:off + m].view(r.K, r.n)
            r.wsums = self.wsums[k0:k0 + r.K]
            off += m
        self._pw_calls = 0
        self._pw_live = False
        self.reg.models = []      # (re)built against the new buffers at the next call

    def _pw_sync_policy(self):
        """Host side of PointwiseLosses.__call__ (reference pointwise.py:36-100, 117-152) before the launches: the reset
        at grid updates, the first launch (weights = bias-free Linear(N, 1) modules filled with 1, here views of the
        flat lambda buffer), and which optimizer is the one being stepped.  Quirk kept: after a grid update the new
        weight modules get a NEW optimizer appended while optimizers[current_batch] -- the first one -- keeps being
        stepped, so only the first generation of weights is ever trained."""
        reg, tr = self.reg, self.trainer()
        if tr.current_epoch % tr.update_grid_every == 0 and tr.current_batch == 0:
            reg.ignored_mask = None
            reg.models = []
        if len(reg.models) < 1:
            reg.models.append([])
            reg.ignored_mask = []
            self.pw_lam.fill_(1.0)
            self.pw_train.fill_(1.0)
            off, index = 0, 0
            for r in self.runners:
                for k in range(r.K):
                    if index in reg.ignored_indexes:
                        reg.ignored_mask.append(True)
                        reg.models[0].append(None)
                        self.pw_train[off:off + r.n] = 0.0
                    else:
                        reg.ignored_mask.append(False)
                        lin = torch.nn.Linear(r.n, 1, bias=False, device=self.pw_lam.device)
                        lin.weight.data = self.pw_lam[off:off + r.n].view(1, r.n)
                        reg.models[0].append(lin)
                    off += r.n
                    index += 1
            reg.optimizers.append(("fused-adam-ascent", self.pw_state, self.pw_m, self.pw_v))
            live = len(reg.optimizers) == 1
            if live:
                self.pw_m.zero_()
                self.pw_v.zero_()
                self.pw_state.zero_()
                self.pw_state[0] = float(reg.lr)
            self._pw_calls = 0
            if live != self._pw_live or not live:
                self._g_compute = None                      # the captured sequence changes
                self._calls["_g_compute"] = 0
            self._pw_live = live

    def _pw_ascent(self):
        """lambda <- Adam(maximize=True, lr) step on d total / d lambda = r^2 of the previous evaluation."""
        be = backend()
        torch.mul(self.pw_vals, self.pw_vals, out=self.pw_g)
        self.pw_g.mul_(self.pw_train)
        be.adam_tick(self.pw_state, 0.9, 0.999)
        be.adam_step_dev(self.pw_lam, self.pw_g, self.pw_m, self.pw_v, self.pw_state, 0.9, 0.999, 1e-8, -1.0)

    def _compute(self, accumulate: bool = False):
        eng = self.engine
        if not accumulate:          # later batches of an epoch add to the gradient of the earlier ones
            eng.flat_grad.zero_()
        self.sums.zero_()
        if self.per_condition:
            self.G.zero_()
        if self.mode == "pointwise":
            self.wsums.zero_()
            if self._pw_live and self._pw_calls > 0:
                self._pw_ascent()
        if self.mode == "adaptive":
            # theta <- Adam(lr, maximize=True, weight_decay=1e-6) on d total / d theta of the PREVIOUS evaluation
            # (AdaptiveWeightSum.forward steps its optimizer before it uses the weights; nothing to step the first time)
            if self._ad_st["calls"] > 0:
                be = backend()
                torch.mul(self.ad_gtheta, -1.0, out=self.ad_g)
                self.ad_g.add_(self.ad_theta, alpha=1e-6)
                be.adam_tick(self.ad_state, 0.9, 0.999)
                be.adam_step_dev(self.ad_theta, self.ad_g, self.ad_m, self.ad_v, self.ad_state, 0.9, 0.999, 1e-8, 1.0)
            torch.softmax(self.ad_theta, dim=0, out=self.ad_w)
            torch.mul(self.ad_w, self.inv_counts, out=self.weights)     # seeds: 2 w_k r / N_k
        for u in self.units:
            if isinstance(u, BatchedGroup):
                u.launch(self.weights, self.sums, eng.flat_grad, [self.k0[self.runners.index(m)] for m in u.members])
            else:
                c = self.runners.index(u)
                k0 = self.k0[c]
                u.launch(self.weights[k0:k0 + u.K], self.sums[k0:k0 + u.K],
                         self.G[c] if self.per_condition else eng.flat_grad)
        lv = self._loss_view()
        K = self.K
        if self.mode == "constant":      # reported losses are the weighted ones (reference constant.py:33-36)
            torch.mul(self.sums, self.weights, out=lv[:K])
            torch.sum(lv[:K], dim=0, out=lv[K])
            return
        torch.mul(self.sums, self.inv_counts, out=lv[:K])
        if self.mode == "pointwise":     # reported: unweighted means; optimised: the weighted sums
            torch.sum(self.wsums, dim=0, out=lv[K])
            return
        if self.mode == "adaptive":      # total = w . L ;  d total / d theta = w * (L - total)   (softmax Jacobian)
            torch.dot(self.ad_w, lv[:K], out=lv[K])
            torch.mul(self.ad_w, lv[:K] - lv[K], out=self.ad_gtheta)
            return
        if self.mode == "mean":
            torch.mul(torch.sum(self.sums), self.inv_total, out=lv[K])
        elif self.mode == "legacy":
            torch.sum(lv[:K], dim=0, out=lv[K])
        else:
            # GradientLosses (reference gradient.py:35-59):  lambda_k <- (1-a) lambda_k + a max|g_0| / mean-of-means|g_k|
            # NormalLosses   (reference normal.py:27-45):    lambda_k <- (1-a) lambda_k + a std(g_0[W_last]) / std(g_k[W_last])
            # then  total = L_0 + sum_k lambda_k L_k  and its gradient  g_0 + sum_k lambda_k g_k
            G, lam, a = self.G, self.reg.lambda_regularization, float(self.reg.alpha)
            if self.mode == "gradient":
                ratio = G[0].abs().max() / (G[1:].abs() * self.tensor_mean_w).sum(dim=1)
            else:
                ratio = torch.std(G[0, self.last_slice]) / torch.std(G[1:, self.last_slice], dim=1)
            lam.mul_(1.0 - a).add_(ratio, alpha=a)
            torch.add(G[0], torch.mv(G[1:].t(), lam), out=eng.flat_grad)
            torch.add(lv[0], torch.dot(lam, lv[1:K]), out=lv[K])

    def _update(self, grad_scale: float):
        eng = self.engine
        fa = self.flat_adam
        g = fa.optimizer.param_groups[0]
        b1, b2 = g["betas"]
        be = backend()
        be.adam_tick(self.adam_state, b1, b2)
        be.adam_step_dev(eng.flat.detach(), eng.flat_grad, fa.m, fa.v, self.adam_state, b1, b2, g["eps"], grad_scale)
        be.pack_params(eng.ops, eng.flat.detach(), eng.packed)

    def _graphed(self, which: str, fn):
        """Run ``fn`` directly twice (kernel/module warm-up), then capture it once and replay."""
        from .engine import CudaBackend

        def direct():
            n0 = CudaBackend.gpu_launches
            fn()
            self._launches[which] = CudaBackend.gpu_launches - n0

        if not (self.use_graphs and self._dev is not None and self._dev.type == "cuda"):
            direct()
            return
        g = getattr(self, which)
        if g is None:
            self._calls[which] += 1
            if self._calls[which] <= 2:
                direct()
                return
            g = torch.cuda.CUDAGraph()
            n0 = CudaBackend.gpu_launches
            with torch.cuda.graph(g):
                fn()
            CudaBackend.gpu_launches = n0      # capture launches nothing
            setattr(self, which, g)
        g.replay()
        self.graph_replays += 1
        CudaBackend.gpu_launches += self._launches[which]   # the replay runs the same kernels as a direct call

    def refresh(self) -> bool:
        """Follow residuals whose closed-over constants changed since they were traced (``Condition.program_for``
        re-traces them): programs of the same shape are swapped in and the captured graph is dropped; False means the
        step has to be rebuilt (or, when a residual no longer traces, abandoned for the eager path)."""
        model, changed = self.engine.model, False
        for r in self.runners:
            p = r.cond.program_for(model)
            if p is r.prog:
                continue
            if p is None or not r.swap_program(p):
                return False
            changed = True
        if changed:
            self._g_compute = self._g_compute_acc = None
            self._calls["_g_compute"] = self._calls["_g_compute_acc"] = 0
        return True

    # ------------------------------------------------------------------ the step
    def compute(self, accumulate: bool = False):
        """Gradient of the total loss into ``engine.flat_grad`` (added to it with ``accumulate``: the second and later
        batches of an epoch, reference trainer.py:119-153) and the losses into the comm tail."""
        conds = [r.cond for r in self.runners]
        self._bind(conds[0].batch_points.device)
        for r in self.runners:
            r.load_batch()
        self.engine._ensure_packed()
        if self.mode == "pointwise":
            self._pw_sync_policy()
        if accumulate:
            if self.mode not in ("mean", "legacy", "constant"):
                raise RuntimeError("gradient accumulation over batches is fused for the mean / legacy / constant policies")
            self._graphed("_g_compute_acc", lambda: self._compute(True))
        else:
            self._graphed("_g_compute", self._compute)
        if self.mode == "pointwise":
            self._pw_calls += 1
        if self.mode == "adaptive":
            self._ad_st["calls"] += 1

    def update(self, grad_scale: float = 1.0) -> bool:
        """Optimizer step on the device; False if the optimizer is not the fused Adam (caller steps it)."""
        fa = self.flat_adam
        if fa is None:
            return False
        if fa.step != self._step_on_device:   # optimizer state was loaded from outside (checkpoint resume)
            self.adam_state[1:2].fill_(float(fa.step))
        lr = float(fa.optimizer.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.adam_state[0:1].fill_(lr)
            self._lr_on_device = lr
        if grad_scale != self._scale:   # baked into the captured launch
            self._scale = grad_scale
            self._g_update = None
        self._graphed("_g_update", lambda: self._update(self._scale))
        fa.step += 1
        self._step_on_device = fa.step
        fa.mark_stepped()
        self.engine.generation += 1      # parameters changed; operands were re-packed inside _update
        self.engine.packed_valid = True
        return True

    def losses(self):
        """(total_loss 0-d, component means [K]) as fresh tensors."""
        lv = self._loss_view().clone()
        return lv[self.K], lv[:self.K]
