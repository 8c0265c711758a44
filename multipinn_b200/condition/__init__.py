"""Mirror of ``multipinn.condition`` (reference ``multipinn/condition/__init__.py``)."""
from .condition import Condition, ConditionExtra
from .diff import (UnsupportedDerivative, grad, num_diff, num_diff_random, num_diff_second_cross,
                   num_diff_second_same, num_laplace, unpack)
from .symbols import Symbols

__all__ = ["Condition", "ConditionExtra", "Symbols", "grad", "unpack", "num_diff", "num_diff_random",
           "num_diff_second_same", "num_diff_second_cross", "num_laplace", "UnsupportedDerivative"]
