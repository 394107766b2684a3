"""AssignCenters: assignment to user-supplied centers.

Mirrors pyemma/coordinates/clustering/assign.py:33-108 (reference @ 3327f28): centers from an array
or a file (:71-76), 2-D check (:79-80), dimension check against the data producer (:90-98),
_estimate just assigns (:100-108).
"""
import numpy as np

from .interface import AbstractClustering

__all__ = ["AssignCenters"]


class AssignCenters(AbstractClustering):
    def __init__(self, clustercenters, metric="euclidean", stride=1, n_jobs=None, skip=0):
        super().__init__(metric=metric, n_jobs=n_jobs)
        if isinstance(clustercenters, str):
            # the reference goes through create_file_reader (csv / npy); those two formats are kept
            if clustercenters.endswith(".npy"):
                clustercenters = np.load(clustercenters)
            else:
                clustercenters = np.loadtxt(clustercenters, ndmin=2)
        clustercenters = np.array(clustercenters, dtype=np.float32, order="C")
        if not clustercenters.ndim == 2:
            raise ValueError("cluster centers have to be 2d")
        self.set_params(clustercenters=clustercenters, metric=metric, stride=stride, skip=skip)
        self._estimated = True  # centers are given: no estimation required

    def describe(self):
        return "[{name} centers shape={shape}]".format(name=type(self).__name__, shape=self.clustercenters.shape)

    @property
    def n_clusters(self):
        return len(self.clustercenters)

    @property
    def data_producer(self):
        return self._data_producer

    @data_producer.setter
    def data_producer(self, dp):
        if dp is not None:
            dim = self.clustercenters.shape[1]
            if not dim == dp.dimension():
                raise ValueError("cluster centers have wrong dimension. Have dim=%i"
                                 ", but input has %i" % (dim, dp.dimension()))
        AbstractClustering.data_producer.fset(self, dp)

    def _estimate(self, iterable, **kw):
        self.assign(None, self.stride)
        return self
