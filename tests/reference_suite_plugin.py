"""pytest plugin used by tests/test_reference_suite.py: makes `import multipinn...` resolve to the mirror
(multipinn_b200) and routes the kernels through the oracle backend so the reference's own unit tests for the
hot-path interfaces run on CPU against this package."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import multipinn_b200  # noqa: E402
import multipinn_b200.condition.condition  # noqa: E402
import multipinn_b200.condition.diff  # noqa: E402
import multipinn_b200.condition.symbols  # noqa: E402
import multipinn_b200.neural_network.activation_function  # noqa: E402
import multipinn_b200.neural_network.feedforward_neural_network  # noqa: E402
import multipinn_b200.neural_network.fourier_features  # noqa: E402
import multipinn_b200.neural_network.light_residual_block  # noqa: E402
import multipinn_b200.neural_network.residual_neural_network  # noqa: E402
from multipinn_b200 import engine  # noqa: E402
import importlib.util  # noqa: E402

# loaded by path under a private name: the reference's tests are themselves a package called `tests`
_spec = importlib.util.spec_from_file_location("_mpinn_oracle_backend", os.path.join(ROOT, "tests", "oracle_backend.py"))
_ob = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_ob)
OracleBackend = _ob.OracleBackend

ALIASES = ["", ".condition", ".condition.condition", ".condition.diff", ".condition.symbols", ".neural_network",
           ".neural_network.activation_function", ".neural_network.feedforward_neural_network",
           ".neural_network.fourier_features", ".neural_network.light_residual_block",
           ".neural_network.residual_neural_network", ".PINN", ".trainer"]
for a in ALIASES:
    sys.modules["multipinn" + a] = sys.modules["multipinn_b200" + a]
engine._set_backend_for_tests(OracleBackend())
