"""Training loop with the reference's interface (reference ``multipinn/trainer/trainer.py:9-338``):
``Trainer`` (gradient accumulation over ``num_batches``), ``TrainerOneBatch`` and the data-parallel
``TrainerMultiGPU``.

What is different underneath:
  * the gradient lives in ONE flat buffer (the model's parameters and ``.grad`` are views of flat
    storage owned by the ``JetEngine``), so ``zero_grad`` is one memset and the data-parallel sync
    is ONE all-reduce of ``[P + K + 1]`` floats (gradient, K component losses, total loss) instead of
    one call per parameter tensor plus two for the metrics (reference trainer.py:242-258);
  * ``optimizer.step()`` for a plain ``torch.optim.Adam`` runs as one fused kernel over the flat
    buffers; its moments are exposed through ``optimizer.state`` so checkpoints keep working.
    Any other optimizer is stepped through its own ``step()``;
  * rank 0's parameters are broadcast once at start (the reference never synchronises replicas).
"""
from __future__ import annotations

from typing import Callable, List, Union

import torch
import torch.distributed as dist

from .. import jet
from ..engine import backend
from ..fused import FusedStep
from ..neural_network.base import JetModule
from ..PINN import PINN


class _FlatAdam:
    """Fused Adam over the engine's flat buffers, mirrored into ``optimizer.state``."""

    def __init__(self, engine, optimizer: torch.optim.Adam):
        self.engine = engine
        self.optimizer = optimizer
        engine.ensure_flat()
        self.m = torch.zeros_like(engine.flat_grad)
        self.v = torch.zeros_like(engine.flat_grad)
        self.step = 0
        self._step_t = torch.tensor(0.0, device="cpu")   # ONE host tensor shared by every parameter's state["step"]
        for p, off in zip(engine._params, engine._offsets):
            st = optimizer.state.get(p)
            if st:  # resume from an existing torch state
                self.m[off:off + p.numel()].copy_(st["exp_avg"].reshape(-1))
                self.v[off:off + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
                self.step = max(self.step, int(st["step"]))
        self._publish()

    @staticmethod
    def eligible(engine, optimizer) -> bool:
        if type(optimizer) is not torch.optim.Adam or len(optimizer.param_groups) != 1:
            return False
        g = optimizer.param_groups[0]
        if g.get("amsgrad") or g.get("weight_decay", 0) != 0 or g.get("maximize") or \
                g.get("capturable") or g.get("differentiable"):
            return False
        engine.ensure_flat()
        # the reverse kernels produce gradients for every layer and the flat update steps the whole buffer, so a
        # model with frozen layers must go through torch.optim.Adam (which skips parameters without .grad);
        # a frozen Fourier matrix is fine: the fused reverse pass leaves its gradient slot at zero
        ops = engine.ops
        fourier_w = int(ops[0, 3]) if int(ops[0, 0]) == 1 else -1
        for p, off in zip(engine._params, engine._offsets):
            if not p.requires_grad and off != fourier_w:
                return False
        return {id(p) for p in g["params"]} == {id(p) for p in engine._params}

    def _publish(self):
        self._step_t.fill_(float(self.step))
        for p, off in zip(self.engine._params, self.engine._offsets):
            if not p.requires_grad:
                continue
            n = p.numel()
            self.optimizer.state[p] = {
                "step": self._step_t,
                "exp_avg": self.m[off:off + n].view(p.shape),
                "exp_avg_sq": self.v[off:off + n].view(p.shape),
            }

    def apply(self, grad_scale: float = 1.0):
        g = self.optimizer.param_groups[0]
        self.step += 1
        b1, b2 = g["betas"]
        backend().adam_step(self.engine.flat.detach(), self.engine.flat_grad, self.m, self.v, self.step,
                            g["lr"], b1, b2, g["eps"], grad_scale)
        self.mark_stepped()

    def load_state(self, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int):
        """Resume from flat Adam moments (what a checkpoint of ``optimizer.state_dict()`` flattens to)."""
        self.m.copy_(exp_avg.reshape(-1))
        self.v.copy_(exp_avg_sq.reshape(-1))
        self.step = int(step)
        self._step_t.fill_(float(self.step))

    def mark_stepped(self):
        self._step_t.fill_(float(self.step))
        self.optimizer._opt_called = True  # what LRScheduler checks before its first step()


class Trainer:
    def __init__(
        self,
        pinn: PINN,
        optimizer: torch.optim.Optimizer,
        scheduler: torch.optim.lr_scheduler.LRScheduler,
        num_epochs: int = 1000,
        num_batches: int = 1,
        update_grid_every: int = 1,
        calc_loss: Union[str, Callable[[], List[torch.Tensor]]] = "mean",
        callbacks_organizer=None,
        mixed_training: bool = False,
    ) -> None:
        self.pinn = pinn
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.num_epochs = num_epochs
        self.num_batches = num_batches
        self.update_grid_every = update_grid_every if update_grid_every is not None else num_epochs + 1
        if isinstance(calc_loss, str):
            if calc_loss == "mean":
                self.calc_loss = self.mean_calc_loss
            elif calc_loss == "legacy":
                self.calc_loss = self.sum_of_means_calc_loss
            else:
                raise ValueError
        else:
            self.calc_loss = calc_loss
        self.callbacks_organizer = callbacks_organizer
        if mixed_training:
            raise NotImplementedError(
                "mixed_training (fp16 autocast) is not part of the fused fp32-class path")
        self.mixed_training = False
        self.current_epoch = 0
        self.current_batch = 0
        self.inum_batches = 1.0 / self.num_batches
        self.epoch_loss_detailed = None
        self.total_loss = None
        self.current_lr = None
        for cond in self.pinn.conditions:
            cond.set_batching(self.num_batches)
        model = self.pinn.model
        self._engine = model.jet_engine if isinstance(model, JetModule) else None
        self._flat_adam = None
        # Fused step (multipinn_b200/fused.py): used when every condition's residual was traced into a device
        # program and the loss policy is one of the reference's two; otherwise the eager jet path below runs.
        self.use_fused_step = True
        self.use_cuda_graphs = True
        self._fused = None
        if self._engine is not None:
            self._engine.ensure_flat()
            self._engine.attach_grads()
            self._engine.managed = True
            self._engine.invalidate()
            if _FlatAdam.eligible(self._engine, optimizer):
                self._flat_adam = _FlatAdam(self._engine, optimizer)

    # ---------------------------------------------------------------- loop (reference :74-117)
    def train(self):
        for self.current_epoch in range(self.num_epochs + 1):
            self._train_epoch()
            if self.scheduler is not None:
                self.scheduler.step()
                self.current_lr = self.optimizer.param_groups[0]["lr"]
            self._run_callbacks()

    def _run_callbacks(self):
        if self.callbacks_organizer is None:
            return
        self.pinn.model.eval()
        with torch.no_grad():
            for callback in self.callbacks_organizer.base_callbacks:
                callback(self)
        self.pinn.model.train()
        for callback in self.callbacks_organizer.grad_callbacks:
            callback(self)

    def _train_epoch(self):
        self.pinn.model.train()
        if self.current_epoch % self.update_grid_every == 0:
            self.pinn.update_data()
        self.epoch_loss_detailed = None
        self.total_loss = 0
        jet.clear_blocks()
        self._train_batches()

    def _accumulate(self, batch_loss, losses):
        d = losses.detach()
        self.epoch_loss_detailed = d if self.epoch_loss_detailed is None else self.epoch_loss_detailed + d
        self.total_loss = self.total_loss + batch_loss.detach()

    def _zero_grad(self):
        if self._engine is not None:
            self._engine.zero_grad()
        else:
            self.optimizer.zero_grad()

    def _optimizer_step(self, grad_scale: float = 1.0):
        if self._flat_adam is not None:
            self._flat_adam.apply(grad_scale)
        else:
            if grad_scale != 1.0:
                for p in self.pinn.model.parameters():
                    if p.grad is not None:
                        p.grad.mul_(grad_scale)
            self.optimizer.step()
        if self._engine is not None:
            self._engine.invalidate()

    def fused_step(self):
        """The FusedStep executor of this trainer, or None (built lazily at the first step)."""
        if not self.use_fused_step:
            return None
        if self._fused is None:
            self._fused = FusedStep.build(self) or False
            if self._fused is False and self._engine is not None:
                self._explain_eager_path()
        if self._fused:
            self._fused.use_graphs = self.use_cuda_graphs
            return self._fused
        return None

    def _explain_eager_path(self):
        """Say once why this trainer stays on the eager jet path (same kernels, autograd in between; slower)."""
        import logging
        model = self.pinn.model
        notes = [f"condition {i} ({getattr(c.function, '__name__', c.function)}): {c.trace_note}"
                 for i, c in enumerate(self.pinn.conditions)
                 if hasattr(c, "program_for") and c.program_for(model) is None and c.trace_note]
        why = "; ".join(notes) if notes else "the loss policy, the batching or an engine option is not covered by it"
        logging.getLogger("multipinn_b200").info("fused step not used: %s", why)
        self.fused_step_note = why

    def _current_fused_step(self):
        """``fused_step()`` after following residual constants that changed since the trace (GradualTrainer's
        ``param_setter``, reference trainer/gradual_trainer.py:30-35)."""
        fs = self.fused_step()
        if fs is not None and not fs.refresh():
            was_pointwise = fs.mode == "pointwise"
            self._fused = None
            fs = self.fused_step()
            if fs is None and was_pointwise:
                raise RuntimeError(
                    "a residual changed shape (or stopped tracing) while the per-point weights of PointwiseLosses "
                    "live on the fused step; construct the trainer with use_fused_step = False for this problem")
        return fs

    def _fused_batch(self, batch_index: int) -> bool:
        """One whole optimisation step on the fused path; False if this trainer is not eligible."""
        fs = self._current_fused_step()
        if fs is None:
            return False
        self.current_batch = batch_index
        self.pinn.select_batch(batch_index)
        fs.compute()
        if not fs.update():
            self._optimizer_step()       # any other optimizer: gradients are in the p.grad views
        self.total_loss, self.epoch_loss_detailed = fs.losses()
        return True

    def _fused_batches(self) -> bool:
        """An epoch of ``num_batches`` batches on the fused path: the gradient accumulates over the batches (the first
        compute overwrites it, the others add), then ONE optimizer step (reference trainer.py:119-153)."""
        fs = self._current_fused_step()
        if fs is None:
            return False
        for self.current_batch in range(self.num_batches):
            self.pinn.select_batch(self.current_batch)
            fs.compute(accumulate=self.current_batch > 0)
            total, losses = fs.losses()
            self._accumulate(total, losses)
        if not fs.update():
            self._optimizer_step()
        self.epoch_loss_detailed = self.epoch_loss_detailed * self.inum_batches
        self.total_loss = self.total_loss * self.inum_batches
        return True

    def _train_batches(self):
        if self.num_batches == 1 and self._fused_batch(0):
            return
        if self.num_batches > 1 and self._fused_batches():
            return
        self._zero_grad()  # outside the loop: gradients accumulate over the batches
        for self.current_batch in range(self.num_batches):
            self.pinn.select_batch(self.current_batch)
            batch_loss, losses = self.calc_loss(self)
            batch_loss.backward()
            self._accumulate(batch_loss, losses)
        self._optimizer_step()
        self.epoch_loss_detailed = self.epoch_loss_detailed * self.inum_batches
        self.total_loss = self.total_loss * self.inum_batches

    # ---------------------------------------------------------------- loss policies (:155-180)
    @staticmethod
    def sum_of_means_calc_loss(self: "Trainer"):
        losses = torch.stack(self.pinn.calculate_loss())
        return torch.sum(losses), losses

    @staticmethod
    def mean_calc_loss(self: "Trainer"):
        total_loss, losses_list = self.pinn.calculate_mean_loss()
        return total_loss, torch.stack(losses_list)


class TrainerOneBatch(Trainer):
    def __init__(self, pinn, optimizer, scheduler, num_epochs: int = 1000, update_grid_every: int = 1,
                 calc_loss="mean", callbacks_organizer=None, mixed_training: bool = False) -> None:
        super().__init__(pinn, optimizer, scheduler, num_epochs, 1, update_grid_every, calc_loss,
                         callbacks_organizer, mixed_training)

    def _train_batches(self):
        if self._fused_batch(0):
            return
        self.pinn.select_batch(0)
        self._zero_grad()
        batch_loss, losses = self.calc_loss(self)
        batch_loss.backward()
        self._optimizer_step()
        self._accumulate(batch_loss, losses)


class TrainerMultiGPU(Trainer):
    """One process per GPU; rank r trains on slice r of every condition's points
    (reference trainer.py:183-284) and the replicas exchange one flat message per step."""

    def __init__(self, pinn, optimizer, scheduler, num_epochs: int = 1000, num_batches: int = 1,
                 update_grid_every: int = 1, calc_loss="mean", callbacks_organizer=None,
                 mixed_training: bool = False, rank: int = 0, sharded_points: bool = False) -> None:
        super().__init__(pinn, optimizer, scheduler, num_epochs, num_batches, update_grid_every,
                         calc_loss, callbacks_organizer, mixed_training)
        self.rank = rank
        # sharded_points=True: every rank's conditions hold ONLY that rank's slice (generators report
        # the global n_points); the reference instead generates all points on every rank and uses 1/N
        self.sharded_points = sharded_points
        assert dist.get_world_size() == num_batches
        self._comm = None
        if self._engine is not None:  # replicas must start from identical weights
            dist.broadcast(self._engine.flat.detach(), src=0)
            self._engine.invalidate()
        else:
            for p in self.pinn.model.parameters():
                dist.broadcast(p.data, src=0)

    def _fused_dp_step(self) -> bool:
        fs = self._current_fused_step()
        if fs is None:
            return False
        self.pinn.select_batch(0 if self.sharded_points else self.rank)
        fs.compute()                                   # gradient + [K losses, total] in ONE flat message
        eng = self._engine
        msg = eng.comm_buffer(fs.K + 1)
        dist.all_reduce(msg, op=dist.ReduceOp.SUM)
        msg.mul_(1.0 / float(dist.get_world_size()))   # p.grad and the logged losses become means over ranks
        if not fs.update():
            self._optimizer_step()
        self.total_loss, self.epoch_loss_detailed = fs.losses()
        self._comm = None
        return True

    def _train_batches(self):
        if self._fused_dp_step():
            return
        self._zero_grad()
        self.pinn.select_batch(0 if self.sharded_points else self.rank)
        batch_loss, losses = self.calc_loss(self)
        batch_loss.backward()
        self._accumulate(batch_loss, losses)
        self.reduce_gradients()
        self._optimizer_step()
        self.reduce_metrics()

    def reduce_gradients(self):
        """After this call every ``p.grad`` is the mean over ranks (reference :242-251) -- done with one
        all-reduce that also carries the logged losses."""
        world = float(dist.get_world_size())
        if self._engine is None:
            for p in self.pinn.model.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad.data, op=dist.ReduceOp.SUM)
                    p.grad.data /= world
            self._comm = None
            return
        eng = self._engine
        P = eng.flat_grad.numel()
        K = self.epoch_loss_detailed.numel()
        msg = eng.comm_buffer(K + 1)
        msg[P:P + K].copy_(self.epoch_loss_detailed.reshape(-1))
        msg[P + K:P + K + 1].copy_(self.total_loss.reshape(-1))
        dist.all_reduce(msg, op=dist.ReduceOp.SUM)
        msg.mul_(1.0 / world)
        self._comm = (msg, P, K)

    def reduce_metrics(self):
        """``total_loss`` / ``epoch_loss_detailed`` become means over ranks (reference :253-258)."""
        if self._comm is not None:
            msg, P, K = self._comm
            self.epoch_loss_detailed = msg[P:P + K].clone().reshape(self.epoch_loss_detailed.shape)
            self.total_loss = msg[P + K].clone()
            return
        dist.all_reduce(self.total_loss, op=dist.ReduceOp.SUM)
        dist.all_reduce(self.epoch_loss_detailed, op=dist.ReduceOp.SUM)
        self.total_loss /= dist.get_world_size()
        self.epoch_loss_detailed /= dist.get_world_size()

    def _run_callbacks(self):
        if self.rank == 0:
            super()._run_callbacks()
