"""The reference's OTHER trainers driving the patched `Trainer` (build container only: needs /root/reference).

`GradualTrainer` (reference trainer/gradual_trainer.py:8-52) changes a problem parameter through the example's
`nonlocal` setter between epochs; `MarchingTrainer` (reference trainer/marching.py:13-88) re-targets the conditions'
geometries, replaces a condition's function by one that also calls the previous stage's model, and re-initialises the
network in place between stages.  Both are the reference's own, unpatched classes; they call `trainer._train_epoch()` of
whatever `Trainer` they were given.  A subprocess runs each schedule twice on the same seeds and points -- unpatched
reference vs after `multipinn_b200.install()` (fused step where the residuals trace, oracle backend standing in for the
kernels) -- and compares the loss after every epoch and the final weights."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import importlib, importlib.util, os, sys, types
import numpy as np, torch
ROOT, which = sys.argv[1:3]
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
ref = mk.import_reference()
import multipinn_b200
from multipinn_b200 import engine
ob = importlib.util.spec_from_file_location("ob", os.path.join(ROOT, "tests", "oracle_backend.py"))
obm = importlib.util.module_from_spec(ob); ob.loader.exec_module(obm)
engine._set_backend_for_tests(obm.OracleBackend())

class FixedPerGeometry:
    """Deterministic points: the same for both runs, regenerated when a stage changes the geometry."""
    def __init__(self, n): self.n_points = n
    def generate(self, condition, model):
        st = np.random.get_state(); np.random.seed(1234 + int(1000 * float(np.sum(condition.geometry.bbox))))
        p = np.asarray(condition.geometry.random_points(self.n_points, "pseudo"), dtype=np.float32)
        np.random.set_state(st)
        return torch.tensor(p).requires_grad_()
    def use_for(self, c): c.generator = self

def run(patched):
    for k in [k for k in sys.modules if k.startswith("examples.")]:
        del sys.modules[k]
    if patched: multipinn_b200.install()
    else: multipinn_b200.uninstall()
    import multipinn
    from multipinn.trainer.gradual_trainer import GradualTrainer
    from multipinn.trainer.marching import MarchingTrainer
    name = "convection_1D_gradual_learning" if which == "gradual" else "convection_1D_marching"
    mod = importlib.import_module(f"examples.{name}.problem")
    conds, din, dout, hook = mod.convection_problem()
    torch.manual_seed(0)
    model = multipinn.neural_network.FNN(input_dim=din, output_dim=dout, hidden_layers=[64, 64, 64])
    for c in conds:
        FixedPerGeometry(24).use_for(c)
    pinn = multipinn.PINN(model=model, conditions=conds)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    org = types.SimpleNamespace(base_callbacks=[], grad_callbacks=[])
    losses = []
    org.grad_callbacks.append(lambda tr: losses.append(float(tr.total_loss)))
    if which == "gradual":
        trainer = multipinn.Trainer(pinn=pinn, optimizer=opt, scheduler=None, num_epochs=7, update_grid_every=100,
                                    calc_loss="mean", callbacks_organizer=org)
        GradualTrainer(trainer=trainer, param_setter=hook, list_of_params=[2.0, 5.0], g_l_schedule=[3, 5]).train()
    else:
        trainer = multipinn.Trainer(pinn=pinn, optimizer=opt, scheduler=None, num_epochs=100, update_grid_every=100,
                                    calc_loss="mean", callbacks_organizer=org)
        mt = MarchingTrainer(trainer=trainer, divide=hook, steps=[0.0, 0.5, 1.0], epochs_per_iter=[4, 9])
        torch.manual_seed(7)           # reset_weights draws from the global generator
        mt.march_trainer()
    fused = None
    if patched:
        fused = [c.program_for(model) is not None for c in conds]
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    return np.array(losses), flat, fused

l_ref, w_ref, _ = run(False)
l_new, w_new, fused = run(True)
multipinn_b200.uninstall()
assert len(l_ref) == len(l_new) and len(l_ref) >= 8, (len(l_ref), len(l_new))
err = np.abs(l_new - l_ref) / np.abs(l_ref)
assert err.max() < 5e-5, (err, l_ref, l_new)
assert np.abs(w_new - w_ref).max() < 2e-5, np.abs(w_new - w_ref).max()
print("TRAINERS-OK", which, fused, err.max())
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("which", ["gradual", "marching"])
def test_reference_gradual_and_marching_trainers_drive_the_patched_trainer(which, tmp_path):
    script = tmp_path / "run_trainers.py"
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT, which], capture_output=True, text=True, timeout=900,
                       cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-3000:]
    assert "TRAINERS-OK" in r.stdout
