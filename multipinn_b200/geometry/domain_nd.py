"""Minimal axis-aligned box geometry: just enough of the reference's ``Hypercube``
(reference ``multipinn/geometry/domain_nd.py:12-126``) to feed conditions with collocation points
(``bbox``, ``random_points``).  Geometry/CSG is outside the accelerated path."""
import numpy as np

from .geometry import Geometry


class Hypercube(Geometry):
    def __init__(self, low, high):
        if len(low) != len(high):
            raise ValueError("Dimensions of low and high do not match.")
        self.low = np.array(low, dtype="float32")
        self.high = np.array(high, dtype="float32")
        if np.any(self.low > self.high):
            raise ValueError("low > high")
        self.side_length = self.high - self.low
        super().__init__(len(low), (self.low, self.high), float(np.linalg.norm(self.side_length)))

    def random_points(self, n, random="pseudo"):
        if random != "pseudo":
            raise NotImplementedError("only the pseudo-random sampler is provided here")
        x = np.random.random(size=(n, self.dim)).astype("float32")
        return self.side_length * x + self.low
