"""multipinn_b200 -- a B200-native fused PINN train step behind multipinn's Python API.

Mirrors the reference's hot-path surface (``multipinn/__init__.py``): the MLP families of
``neural_network``, ``Condition``/``grad``/``unpack``/``Symbols``, ``PINN`` and the trainers.
All numerics run in hand-written sm_100a kernels (``csrc/``) reached through a C ABI
(``include/mpinn_b200.h``); there is no CPU or eager fallback.
"""
from . import neural_network
from .condition import (Condition, ConditionExtra, Symbols, UnsupportedDerivative, grad, unpack)
from .generation import DeviceGenerator, Generator, ResidualAdaptiveGenerator
from .geometry import Hypercube
from .jet import jet_order
from .neural_network import FFNN, FNN, FourierFeatureNetwork, ResNet, XavierFNN
from .PINN import PINN
from .install import install, uninstall
from .regularization import (AdaptiveConditionsAndPointsLosses, AdaptiveConditionsLosses, ConstantLosses,
                             GradientLosses, NormalLosses, PointwiseLosses)
from .trainer import Trainer, TrainerMultiGPU, TrainerOneBatch

__version__ = "0.1.0"
__all__ = ["neural_network", "Condition", "ConditionExtra", "Symbols", "grad", "unpack", "Generator", "DeviceGenerator", "ResidualAdaptiveGenerator",
           "Hypercube", "jet_order", "FNN", "XavierFNN", "ResNet", "FourierFeatureNetwork", "FFNN",
           "PINN", "ConstantLosses", "GradientLosses", "NormalLosses", "PointwiseLosses", "AdaptiveConditionsLosses",
           "AdaptiveConditionsAndPointsLosses", "install", "uninstall", "Trainer", "TrainerOneBatch", "TrainerMultiGPU", "UnsupportedDerivative"]
