"""Loss-weighting callables plugged into the trainer as ``calc_loss`` (reference ``multipinn/regularization``).
Mirrored so far: ``ConstantLosses`` and ``GradientLosses`` (the shipped default of poisson_2D_1C)."""
from .basic import BasicLosses
from .constant import ConstantLosses
from .gradient import GradientLosses

__all__ = ["BasicLosses", "ConstantLosses", "GradientLosses"]
