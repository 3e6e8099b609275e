"""stub of pyclustering.cluster.center_initializer (see ../__init__.py)."""
import numpy as np


class kmeans_plusplus_initializer:
    def __init__(self, data, amount_centers, *a, **kw):
        self.data, self.k = np.asarray(data, dtype=np.float64), int(amount_centers)

    def initialize(self):
        uniq = np.unique(self.data, axis=0)
        idx = np.linspace(0, len(uniq) - 1, min(self.k, len(uniq))).astype(int)
        return uniq[idx].tolist()
