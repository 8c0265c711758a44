"""In-place drop-in: ``multipinn_b200.install()`` patches an importable reference ``multipinn`` package so that its
hot path -- the three MLP families, ``Condition`` / ``grad`` / ``unpack`` / ``Symbols``, ``PINN``, the trainers and
the mirrored regularizers -- resolves to this package, while everything outside the path (geometry, samplers,
callbacks, plotting, metrics, the other network families, Hydra glue) stays the reference's own code.

    import multipinn_b200
    multipinn_b200.install()          # before the example / training script imports multipinn.*
    runpy.run_module("examples.poisson_2D_1C.run_train", run_name="__main__")

so that e.g. ``examples/poisson_2D_1C/run_train.py:22-80`` (``from multipinn import *``; ``poisson_2D_1C()``;
``initialize_model``; ``PINN``; ``Trainer(...).train()``) runs unmodified on the fused kernels.

The patch is by attribute replacement on the reference's own modules (not by swapping whole modules), so names the
mirror does not provide (``DenseNet``, ``GNN``, ``GPINN``, ``MarchingTrainer`` ...) keep working; ``uninstall()``
restores every replaced attribute.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Tuple

_saved: List[Tuple[object, str, object, bool]] = []

# reference module -> {attribute: (mirror module, attribute)}
_TABLE: Dict[str, Dict[str, Tuple[str, str]]] = {
    "multipinn.neural_network.activation_function": {
        n: ("multipinn_b200.neural_network.activation_function", n) for n in ("GELU", "Sine", "fused_gelu", "fused_sin")},
    "multipinn.neural_network.feedforward_neural_network": {
        n: ("multipinn_b200.neural_network.feedforward_neural_network", n) for n in ("FNN", "XavierFNN")},
    "multipinn.neural_network.light_residual_block": {
        "LightResidualBlock": ("multipinn_b200.neural_network.light_residual_block", "LightResidualBlock")},
    "multipinn.neural_network.residual_neural_network": {
        "ResNet": ("multipinn_b200.neural_network.residual_neural_network", "ResNet")},
    "multipinn.neural_network.fourier_features": {
        n: ("multipinn_b200.neural_network.fourier_features", n) for n in ("FourierEncoding", "FourierFeatureNetwork")},
    "multipinn.neural_network.inference_wrapper": {
        "Inference": ("multipinn_b200.neural_network.inference_wrapper", "Inference")},
    "multipinn.neural_network": {
        **{n: ("multipinn_b200.neural_network", n) for n in ("FNN", "XavierFNN", "ResNet", "FourierFeatureNetwork",
                                                             "Inference")},
        "FFNN": ("multipinn_b200.neural_network", "FourierFeatureNetwork")},
    "multipinn.condition.condition": {
        n: ("multipinn_b200.condition.condition", n) for n in ("Condition", "ConditionExtra")},
    "multipinn.condition.diff": {
        n: ("multipinn_b200.condition.diff", n) for n in (
            "grad", "unpack", "num_diff", "num_diff_random", "num_diff_second_same", "num_diff_second_cross",
            "num_laplace")},
    "multipinn.condition.symbols": {"Symbols": ("multipinn_b200.condition.symbols", "Symbols")},
    "multipinn.condition": {
        n: ("multipinn_b200.condition", n) for n in (
            "Condition", "ConditionExtra", "Symbols", "grad", "unpack", "num_diff", "num_diff_random",
            "num_diff_second_same", "num_diff_second_cross", "num_laplace")},
    "multipinn.PINN.PINN": {"PINN": ("multipinn_b200.PINN.PINN", "PINN")},
    "multipinn.PINN": {"PINN": ("multipinn_b200.PINN", "PINN")},
    "multipinn.trainer.trainer": {
        n: ("multipinn_b200.trainer.trainer", n) for n in ("Trainer", "TrainerOneBatch", "TrainerMultiGPU")},
    "multipinn.trainer": {
        n: ("multipinn_b200.trainer", n) for n in ("Trainer", "TrainerOneBatch", "TrainerMultiGPU")},
    "multipinn.regularization": {
        n: ("multipinn_b200.regularization", n) for n in (
            "ConstantLosses", "GradientLosses", "NormalLosses", "PointwiseLosses", "AdaptiveConditionsLosses",
            "AdaptiveConditionsAndPointsLosses")},
    "multipinn": {
        **{n: ("multipinn_b200.neural_network", n) for n in ("FNN", "XavierFNN", "ResNet", "FourierFeatureNetwork")},
        "FFNN": ("multipinn_b200.neural_network", "FourierFeatureNetwork"),
        "PINN": ("multipinn_b200.PINN", "PINN"),
        **{n: ("multipinn_b200.trainer", n) for n in ("Trainer", "TrainerOneBatch", "TrainerMultiGPU")},
        **{n: ("multipinn_b200.regularization", n) for n in (
            "ConstantLosses", "GradientLosses", "NormalLosses", "PointwiseLosses", "AdaptiveConditionsLosses",
            "AdaptiveConditionsAndPointsLosses")},
    },
}


def installed() -> bool:
    return bool(_saved)


def install(strict: bool = False) -> List[str]:
    """Patch the importable ``multipinn`` package in place; returns the list of ``module.attr`` names replaced.
    Modules of the reference that cannot be imported (missing optional dependencies) are skipped unless
    ``strict``.  Idempotent."""
    if _saved:
        return [f"{m.__name__}.{a}" for m, a, _, _ in _saved]
    done = []
    for ref_name, attrs in _TABLE.items():
        try:
            ref_mod = importlib.import_module(ref_name)
        except Exception:
            if strict:
                raise
            continue
        if ref_mod.__name__.startswith("multipinn_b200"):
            continue   # `multipinn` already IS the mirror (alias set up by a test harness)
        for attr, (mir_name, mir_attr) in attrs.items():
            mir_mod = importlib.import_module(mir_name)
            if not hasattr(mir_mod, mir_attr):
                continue
            had = hasattr(ref_mod, attr)
            _saved.append((ref_mod, attr, getattr(ref_mod, attr, None), had))
            setattr(ref_mod, attr, getattr(mir_mod, mir_attr))
            done.append(f"{ref_name}.{attr}")
    # modules that did `from multipinn... import X` before install() hold stale references: refresh the common
    # ones (the reference's own utils resolve model classes through multipinn.neural_network at call time)
    for mod in list(sys.modules.values()):
        name = getattr(mod, "__name__", "")
        if not name.startswith("multipinn.") or name.startswith("multipinn_b200"):
            continue
        for ref_mod, attr, old, had in list(_saved):
            if had and getattr(mod, attr, None) is old and mod is not ref_mod:
                _saved.append((mod, attr, old, True))
                setattr(mod, attr, getattr(ref_mod, attr))
    return done


def uninstall() -> None:
    while _saved:
        mod, attr, old, had = _saved.pop()
        if had:
            setattr(mod, attr, old)
        else:
            try:
                delattr(mod, attr)
            except AttributeError:
                pass
