from .domain_nd import Hypercube
from .geometry import Geometry

__all__ = ["Geometry", "Hypercube"]
