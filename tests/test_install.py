"""`multipinn_b200.install()`: the UNMODIFIED reference examples run through this package (build container only:
needs /root/reference).  A subprocess imports the real `multipinn` (with stand-ins for the plotting / Hydra
packages that are not installed here and that the hot path never touches), patches it in place, imports
`examples/<name>/problem.py` untouched, and trains 3 steps through `multipinn.PINN` / `multipinn.TrainerOneBatch`
-- which now are the fused ones -- against the golden vectors of the unmodified reference."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import importlib, importlib.util, os, sys
import numpy as np, torch
ROOT, name, modname, factory, model_name = sys.argv[1:6]
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
ref = mk.import_reference()                       # the real reference package (stubs for plotly / hydra ...)
import multipinn_b200
from multipinn_b200 import engine
ob = importlib.util.spec_from_file_location("ob", os.path.join(ROOT, "tests", "oracle_backend.py"))
obm = importlib.util.module_from_spec(ob); ob.loader.exec_module(obm)
engine._set_backend_for_tests(obm.OracleBackend())   # no GPU here: the oracle stands in for the kernels
ref_cond = ref.condition.condition.Condition
patched = multipinn_b200.install()
assert "multipinn.condition.condition.Condition" in patched and "multipinn.trainer.trainer.Trainer" in patched
import multipinn
assert multipinn.condition.condition.Condition is multipinn_b200.Condition is not ref_cond
assert multipinn.PINN is multipinn_b200.PINN and multipinn.TrainerOneBatch is multipinn_b200.TrainerOneBatch
assert multipinn.neural_network.DenseNet.__module__.startswith("multipinn.")      # outside the path: untouched
mod = importlib.import_module(modname)            # the UNMODIFIED example: `from multipinn.condition... import ...`
conds, din, dout = getattr(mod, factory)()
assert type(conds[0]) is multipinn_b200.Condition
assert type(conds[0].geometry).__module__.startswith("multipinn.geometry")         # the reference's own geometry
g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
kw = dict(hidden_layers=[int(g["w::" + k].shape[0]) for k in g["state_keys"] if k.endswith("weight")][:-1])
if model_name == "FNN" and name == "poisson_2D_1C":
    kw = dict(hidden_layers=[64] * 9)
elif name == "burgers_1D":
    kw = dict(hidden_layers=[128] * 2)
model = getattr(multipinn.neural_network, model_name)(input_dim=din, output_dim=dout, **kw)
model.load_state_dict({k: torch.tensor(g["w::" + k]) for k in g["state_keys"]}, strict=True)
for i, c in enumerate(conds):
    c.generator = mk._FixedGenerator(torch.tensor(g[f"pts{i}"]).requires_grad_())
pinn = multipinn.PINN(model=model, conditions=conds)
pinn.update_data()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9999)
tr = multipinn.TrainerOneBatch(pinn, opt, sched, num_epochs=3, update_grid_every=10**9, calc_loss="mean")
assert all(c.program_for(model) is not None for c in conds), [c.trace_note for c in conds]
assert tr.fused_step() is not None
traj = []
for e in range(3):
    tr.current_epoch = e + 1
    tr._train_epoch()
    sched.step()
    traj.append(float(tr.total_loss))
    if e == 0:
        assert np.abs(tr.epoch_loss_detailed.numpy() - g["losses_f64"]).max() / np.abs(g["losses_f64"]).max() < 1e-5
        assert np.abs(model.jet_engine.flat.detach().numpy() - g["params_after1"]).max() < 5e-6
assert np.abs(np.array(traj) - g["traj_f32"]).max() / np.abs(g["traj_f32"]).max() < 2e-5
assert np.abs(model.jet_engine.flat.detach().numpy() - g["params_after3"]).max() < 1e-5
multipinn_b200.uninstall()
assert multipinn.condition.condition.Condition is ref_cond
print("INSTALL-OK", name, traj)
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name,modname,factory,model_name", [
    ("poisson_2D_1C", "examples.poisson_2D_1C.problem", "poisson_2D_1C", "FNN"),
    ("burgers_1D", "examples.burgers_1D.problem", "problem_burgers1D", "FNN"),
])
def test_unmodified_reference_examples_run_through_install(name, modname, factory, model_name, tmp_path):
    script = tmp_path / "run_install.py"
    script.write_text(SCRIPT)
    r = subprocess.run([sys.executable, str(script), ROOT, name, modname, factory, model_name],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-3000:]
    assert "INSTALL-OK" in r.stdout
