"""Run the REFERENCE's own unit tests for the hot-path interfaces against this package (build container only:
needs /root/reference).  `multipinn.*` imports are aliased to `multipinn_b200.*` by tests/reference_suite_plugin.py
and the kernels are replaced by the oracle backend, so this checks API fidelity of the host mirror:
constructors, layer structure, state, shapes, batching, Symbols code generation, activation known answers."""
import os
import subprocess
import sys

import pytest

REF_TESTS = "/root/reference/tests"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FILES = [
    "neural_network/test_activation_function.py",
    "neural_network/test_feedforward_neural_network.py",
    "neural_network/test_residual_neural_network.py",
    "neural_network/test_light_residual_block.py",
    "neural_network/test_fourier_features.py",
    "condition/test_condition.py",
    "condition/test_symbols.py",
    "condition/test_diff.py",
]
# upstream tests that cannot pass for reasons unrelated to the mirror
KNOWN = {
    # fails against the unmodified reference too: passes a Mock model to code that calls model.parameters()
    "test_init_output_len",
    # MultiScaleFFNN (several encoders feeding one shared trunk) is a family outside the accelerated path
    "test_multiscale_ffnn_forward",
}


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not present (GPU box)")
def test_reference_unit_tests_pass_against_the_mirror(tmp_path):
    files = [os.path.join(REF_TESTS, f) for f in FILES if os.path.exists(os.path.join(REF_TESTS, f))]
    assert files, "no reference test files found"
    deselect = " and ".join(f"not {k}" for k in sorted(KNOWN))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "reference_suite_plugin", "-p", "no:cacheprovider",
           "--rootdir", str(tmp_path), "-k", deselect, "-W", "ignore"] + files
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests") + os.pathsep + ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = "\n".join(r.stdout.strip().splitlines()[-25:])
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    assert " passed" in r.stdout and " failed" not in r.stdout.splitlines()[-1], tail
    print(r.stdout.strip().splitlines()[-1])
