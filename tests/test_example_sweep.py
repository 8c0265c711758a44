"""Drop-in check on reference examples BEYOND the six BASELINE configs (build container only: needs /root/reference).
`scripts/example_sweep.py <name>` imports the unmodified `examples/<name>/problem.py` twice -- once against the
unpatched reference, once after `multipinn_b200.install()` -- and compares per-condition losses and the parameter
gradient at the same weights and points (oracle backend standing in for the kernels).  The examples chosen here are the
ones whose residuals needed tracer support that the six BASELINE configs do not exercise: a shifted second model call
with a derivative w.r.t. the unshifted points, batch-sized constants and row sums, masked assignment, a real tensor on
the left of a traced column, Symbols-generated residuals, constant components."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = ["portal_1D_1C", "laplace_2D_neumann", "gas_dynamics_1D", "heat_2D_simple", "wave_1D_1C",
         "navier_stokes_3D_pipe_marching", "poisson_3D"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", CASES)
def test_unmodified_example_matches_the_unpatched_reference_on_the_fused_step(name, tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "example_sweep.py"), name],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    out = r.stdout
    assert r.returncode == 0, out[-2000:] + "\n" + r.stderr[-3000:]
    assert "1 match, 0 differ, 0 skipped" in out, out[-2000:] + "\n" + r.stderr[-2000:]
    line = next(l for l in out.splitlines() if l.startswith("OK"))
    assert " fused " in line, line           # every condition traced: the whole step runs on the fused path
    assert "EAGER" not in out
