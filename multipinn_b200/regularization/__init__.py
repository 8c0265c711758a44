"""Loss-weighting callables plugged into the trainer as ``calc_loss`` (reference ``multipinn/regularization``).
Mirrored: ``ConstantLosses``, ``GradientLosses`` (the shipped default of poisson_2D_1C), ``NormalLosses`` and
``PointwiseLosses`` (the shipped default of navier_stokes_2D_obstacle_multigpu) and the two self-adaptive ones."""
from .adaptive import AdaptiveConditionsAndPointsLosses, AdaptiveConditionsLosses, AdaptiveWeightSum
from .basic import BasicLosses
from .constant import ConstantLosses
from .gradient import GradientLosses
from .normal import NormalLosses
from .pointwise import PointwiseLosses

__all__ = ["BasicLosses", "ConstantLosses", "GradientLosses", "NormalLosses", "PointwiseLosses",
           "AdaptiveConditionsLosses", "AdaptiveConditionsAndPointsLosses", "AdaptiveWeightSum"]
