"""Abstract geometry interface the conditions are written against (reference
``multipinn/geometry/geometry.py:20-57``): a dimension, a bounding box, a characteristic size and a point
sampler.  Only the interface lives here; shapes/CSG are outside the accelerated path."""
import abc

import numpy as np


class Geometry(abc.ABC):
    def __init__(self, dim, bbox, diam):
        self.dim = dim
        self.bbox = bbox
        # never larger than the bounding box diagonal
        self.diam = min(diam, float(np.linalg.norm(np.asarray(bbox[1]) - np.asarray(bbox[0]))))

    @abc.abstractmethod
    def random_points(self, n, random="pseudo"):
        """``n`` points of shape ``[n, dim]`` inside (or on) the geometry."""
        raise NotImplementedError
