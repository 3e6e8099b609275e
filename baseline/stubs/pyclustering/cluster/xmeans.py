"""Deterministic stand-in for pyclustering's x-means on the learner's similarity features (PARITY UNPINNED).

The reference clusters rows of ``[rewards_t, clean_num_t]`` with both columns in {0, 1}
(src/learners/homophily_learner.py:184-203), i.e. at most four distinct points, with ``kmax = 4``.  On such data any
x-means run ends with clusters that are unions of identical points; this stand-in returns one cluster per distinct
point (the finest such partition, <= kmax).  pyclustering's k-means++ seeding is random and the package is absent, so
no bit-level claim is made for the learner's similarity loss -- the env path does not depend on it."""
import numpy as np


class xmeans:
    def __init__(self, data, initial_centers=None, kmax=20, *a, **kw):
        self.data, self.kmax = np.asarray(data, dtype=np.float64), int(kmax)
        self._clusters = []

    def process(self):
        d = self.data.reshape(len(self.data), -1)
        uniq, inv = np.unique(d, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        if len(uniq) > self.kmax:                     # not reachable for {0,1}^2 data; keep kmax groups
            inv = np.minimum(inv, self.kmax - 1)
        self._clusters = [np.nonzero(inv == k)[0].tolist() for k in range(min(len(uniq), self.kmax))]
        return self

    def get_clusters(self):
        return self._clusters
