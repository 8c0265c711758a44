"""Data-parallel trainer on 2 CPU processes (gloo): rank r trains on slice r of every condition and
the replicas exchange ONE flat message per step.  Checks the reference's DP semantics
(trainer/trainer.py:219-258): gradients and logged losses are means over ranks, replicas stay
bit-identical, and the result equals single-process training on the full point set."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp_

NAME = "allen_cahn_2D_1C"
STEPS = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup_backend():
    from multipinn_b200 import engine
    from tests.oracle_backend import OracleBackend

    engine._set_backend_for_tests(OracleBackend())


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import multipinn_b200 as mp
        from tests.test_host_cpu import build_pinn
        from tests.test_oracle_golden import load

        _setup_backend()
        torch.manual_seed(100 + rank)  # replicas deliberately start from different weights...
        g = load(NAME)
        model, conds, pinn = build_pinn(NAME, g, world=world)
        if rank != 0:
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(0.01 * torch.randn_like(p))
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        calls = {"n": 0}
        real_all_reduce = dist.all_reduce

        def counting_all_reduce(*a, **k):
            calls["n"] += 1
            return real_all_reduce(*a, **k)

        tr = mp.TrainerMultiGPU(pinn, opt, None, num_epochs=STEPS, num_batches=world,
                                update_grid_every=10**9, rank=rank)  # ...and are synchronised here
        dist.all_reduce = counting_all_reduce
        losses = []
        for e in range(STEPS):
            tr.current_epoch = e + 1
            tr._train_epoch()
            losses.append([float(tr.total_loss)] + tr.epoch_loss_detailed.tolist())
        dist.all_reduce = real_all_reduce
        np.savez(os.path.join(outdir, f"rank{rank}.npz"),
                 flat=model.jet_engine.flat.detach().numpy(), losses=np.array(losses),
                 grad=model.jet_engine.flat_grad.numpy(), n_allreduce=calls["n"])
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    world = 2
    port = _free_port()
    with tempfile.TemporaryDirectory() as outdir:
        mp_.spawn(_worker, args=(world, port, outdir), nprocs=world, join=True)
        r0 = np.load(os.path.join(outdir, "rank0.npz"))
        r1 = np.load(os.path.join(outdir, "rank1.npz"))
    assert np.array_equal(r0["flat"], r1["flat"])            # replicas identical after training
    assert np.array_equal(r0["losses"], r1["losses"])        # metrics are global means on every rank
    assert int(r0["n_allreduce"]) == STEPS                   # ONE collective per step

    # single-process run on the full point set, same initial weights (rank 0's)
    import multipinn_b200 as mp
    from multipinn_b200 import engine
    from tests.oracle_backend import OracleBackend
    from tests.test_host_cpu import build_pinn
    from tests.test_oracle_golden import load

    old = engine._set_backend_for_tests(OracleBackend())
    try:
        model, conds, pinn = build_pinn(NAME, load(NAME), world=1)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        tr = mp.TrainerOneBatch(pinn, opt, None, num_epochs=STEPS, update_grid_every=10**9)
        losses = []
        for e in range(STEPS):
            tr.current_epoch = e + 1
            tr._train_epoch()
            losses.append(float(tr.total_loss))
        single = model.jet_engine.flat.detach().numpy()
    finally:
        engine._set_backend_for_tests(old)
    # equal shards => mean over ranks of per-rank mean losses == global mean loss
    assert np.allclose(r0["losses"][:, 0], losses, rtol=1e-5)
    assert np.abs(r0["flat"] - single).max() < 5e-6
