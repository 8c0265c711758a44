"""Mirror of ``multipinn.neural_network`` for the three MLP families on the fused path
(reference ``multipinn/neural_network/__init__.py``)."""
from .activation_function import GELU, Sine, fused_gelu, fused_sin
from .feedforward_neural_network import FNN, XavierFNN
from .fourier_features import FourierEncoding, FourierFeatureNetwork
from .fourier_features import FourierFeatureNetwork as FFNN
from .inference_wrapper import Inference
from .light_residual_block import LightResidualBlock
from .residual_neural_network import ResNet

__all__ = ["FNN", "XavierFNN", "ResNet", "LightResidualBlock", "FourierEncoding",
           "FourierFeatureNetwork", "FFNN", "GELU", "Sine", "fused_gelu", "fused_sin", "Inference"]
