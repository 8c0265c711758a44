"""2+ GPUs under torchrun: the flow of examples/navier_stokes_2D_obstacle_multigpu/run_train.py -- every process calls
torch.set_default_device(f"cuda:{rank}") (reference utils/device.py:47), builds model / generators / TrainerMultiGPU
without naming a device, trains a few epochs on the fused step, and the replicas must end bit-identical.
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_default_device.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipinn_b200 as mp  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", rank=rank, world_size=world)
torch.set_default_device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.manual_seed(1234 + rank)           # replicas start differently; the trainer broadcasts rank 0's weights


def inner(model, arg):
    u, v, p = mp.unpack(model(arg))
    u_x, u_y = mp.unpack(mp.grad(u, arg))
    v_x, v_y = mp.unpack(mp.grad(v, arg))
    p_x, p_y = mp.unpack(mp.grad(p, arg))
    u_xx, _ = mp.unpack(mp.grad(u_x, arg))
    _, u_yy = mp.unpack(mp.grad(u_y, arg))
    v_xx, _ = mp.unpack(mp.grad(v_x, arg))
    _, v_yy = mp.unpack(mp.grad(v_y, arg))
    return [u * u_x + v * u_y + p_x - 0.02 * (u_xx + u_yy), u * v_x + v * v_y + p_y - 0.02 * (v_xx + v_yy), u_x + v_y]


def wall(model, arg):
    u, v, p = mp.unpack(model(arg))
    return [u, v]


conds = [mp.Condition(inner, mp.Hypercube([0, 0], [2, 1])), mp.Condition(wall, mp.Hypercube([0, 0], [2, 0])),
         mp.Condition(wall, mp.Hypercube([0, 1], [2, 1]))]
model = mp.ResNet(input_dim=2, output_dim=3, hidden_layers=[64, 64, 64])
mp.Generator(n_points=8192 * world, sampler="pseudo").use_for(conds[0])
mp.Generator(n_points=512 * world, sampler="pseudo").use_for(conds[1:])
pinn = mp.PINN(model=model, conditions=conds)
opt = torch.optim.Adam(params=model.parameters(), lr=1e-3)
sched = torch.optim.lr_scheduler.ExponentialLR(optimizer=opt, gamma=0.999)
tr = mp.TrainerMultiGPU(pinn=pinn, optimizer=opt, scheduler=sched, num_epochs=8, num_batches=world, update_grid_every=4,
                        calc_loss="mean", callbacks_organizer=None, rank=rank)
tr.train()
assert tr.fused_step() is not None, [c.trace_note for c in conds]
flat = model.jet_engine.flat.detach().double()
mine = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
allsums = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allsums, mine)
ok = all(bool(torch.equal(a, allsums[0])) for a in allsums)
if rank == 0:
    print(f"DP-DEFAULT-DEVICE world={world} fused=True replicas_in_sync={ok} loss={float(tr.total_loss):.4e} "
          f"graph_replays={tr.fused_step().graph_replays}")
assert ok
dist.destroy_process_group()
