"""The reference's run scripts call `set_device_and_seed` first, which ends in `torch.set_default_device("cuda")`
(reference utils/device.py:23-50): models, points and every tensor user code creates without a device land on the GPU.
The mirror must work in that regime exactly as with explicit `.cuda()` calls -- the flow of
examples/poisson_2D_1C/run_train.py:22-80 (and the regularizer / ConditionExtra / multi-batch variants) under it."""
import types

import numpy as np
import pytest
import torch

import multipinn_b200 as mp

pytestmark = pytest.mark.gpu


@pytest.fixture()
def default_cuda():
    from multipinn_b200 import engine, jet
    engine._set_backend_for_tests(None)
    jet.clear_blocks()
    torch.set_default_device("cuda")
    torch.set_default_dtype(torch.float32)
    yield
    torch.set_default_device("cpu")


def _poisson(ns=mp):
    def inner(model, arg):
        (u,) = ns.unpack(model(arg))
        x, y = ns.unpack(arg)
        u_x, u_y = ns.unpack(ns.grad(u, arg))
        u_xx, _ = ns.unpack(ns.grad(u_x, arg))
        _, u_yy = ns.unpack(ns.grad(u_y, arg))
        return [u_xx + u_yy + 2 * torch.pi**2 * torch.sin(torch.pi * x) * torch.sin(torch.pi * y)]

    def bc(model, arg):
        (u,) = ns.unpack(model(arg))
        return [u]

    domain = mp.Hypercube([0, 0], [1, 1])
    return [mp.Condition(inner, domain), mp.Condition(bc, mp.Hypercube([0, 0], [0, 1])),
            mp.Condition(bc, mp.Hypercube([1, 0], [1, 1]))], 2, 1


@pytest.mark.parametrize("policy", ["mean", "constant", "gradient", "pointwise", "adaptive", "eager"])
def test_run_train_flow_under_default_device_cuda(policy, default_cuda):
    conds, din, dout = _poisson()
    torch.manual_seed(0)
    model = mp.FNN(input_dim=din, output_dim=dout, hidden_layers=[64, 64, 64])
    assert next(model.parameters()).is_cuda
    mp.Generator(n_points=4096, sampler="pseudo").use_for(conds[0])
    mp.Generator(n_points=512, sampler="pseudo").use_for(conds[1:])
    pinn = mp.PINN(model=model, conditions=conds)
    opt = torch.optim.Adam(params=model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(optimizer=opt, gamma=0.999)
    calc = {"mean": "mean", "eager": "mean", "constant": mp.ConstantLosses([1.0, 10.0, 10.0]),
            "gradient": mp.GradientLosses(), "pointwise": mp.PointwiseLosses(),
            "adaptive": mp.AdaptiveConditionsLosses()}[policy]
    seen = []
    org = types.SimpleNamespace(base_callbacks=[lambda tr: seen.append(float(tr.total_loss))], grad_callbacks=[])
    tr = mp.Trainer(pinn=pinn, optimizer=opt, scheduler=sched, num_epochs=12, update_grid_every=5, calc_loss=calc,
                    callbacks_organizer=org)
    if policy == "eager":
        tr.use_fused_step = False
    tr.train()
    assert (tr.fused_step() is not None) == (policy != "eager")
    assert len(seen) == 13 and np.isfinite(seen).all()
    if policy in ("mean", "eager", "constant"):      # (the weight-training policies RAISE their own total by design)
        assert min(seen[1:5]) < seen[0]              # (points regenerate at epochs 5 and 10)
    # inference the way callbacks do it
    with torch.no_grad():
        y = model(torch.rand(1000, 2))
    assert y.is_cuda and y.shape == (1000, 1) and bool(torch.isfinite(y).all())
    sd = model.state_dict()
    assert all(v.is_cuda for v in sd.values())


def test_condition_extra_and_multi_batch_under_default_device_cuda(default_cuda):
    conds, din, dout = _poisson()

    def neumann(model, arg, data):
        (u,) = mp.unpack(model(arg))
        u_x, u_y = mp.unpack(mp.grad(u, arg))
        n = data[0]
        return [n[:, 0] * u_x + n[:, 1] * u_y]

    conds.append(mp.ConditionExtra(neumann, mp.Hypercube([0, 1], [1, 1]),
                                   [lambda pts: torch.tensor([[0.0, 1.0]]).repeat(pts.shape[0], 1)]))
    torch.manual_seed(0)
    model = mp.FNN(input_dim=din, output_dim=dout, hidden_layers=[64, 64])
    mp.Generator(n_points=2048, sampler="pseudo").use_for(conds[0])
    mp.Generator(n_points=256, sampler="pseudo").use_for(conds[1:])
    pinn = mp.PINN(model=model, conditions=conds)
    opt = torch.optim.Adam(params=model.parameters(), lr=1e-3)
    tr = mp.Trainer(pinn=pinn, optimizer=opt, scheduler=None, num_epochs=6, update_grid_every=3, calc_loss="mean")
    tr.train()
    assert tr.fused_step() is not None and conds[-1].program_for(model).data_width == 2
    first = float(tr.total_loss)
    # several batches per epoch: needs a ConditionExtra-free problem (the reference's data slicing is per tuple)
    conds2, _, _ = _poisson()
    mp.Generator(n_points=2048, sampler="pseudo").use_for(conds2[0])
    mp.Generator(n_points=256, sampler="pseudo").use_for(conds2[1:])
    model2 = mp.FNN(input_dim=din, output_dim=dout, hidden_layers=[64, 64])
    tr2 = mp.Trainer(pinn=mp.PINN(model=model2, conditions=conds2),
                     optimizer=torch.optim.Adam(params=model2.parameters(), lr=1e-3), scheduler=None, num_epochs=6,
                     num_batches=4, update_grid_every=3, calc_loss="legacy")
    tr2.train()
    assert tr2.fused_step() is not None
    assert np.isfinite(first) and np.isfinite(float(tr2.total_loss))
