"""UniformTimeClustering: centers picked uniformly in time, then the usual GPU assignment.

Mirrors pyemma/coordinates/clustering/uniform_time.py:33-106 (reference @ 3327f28): n_clusters default
min(sqrt(N), 5000) (:69-74), clipping to the number of frames (:77-83), the index formula (:85-90) and the mapping of a
global frame index to (trajectory, frame) (:98-106).  No arithmetic of its own; `dtrajs` goes through the shared
assignment path (interface.py).
"""
import math

import numpy as np

from .interface import AbstractClustering

__all__ = ["UniformTimeClustering"]


class UniformTimeClustering(AbstractClustering):
    def __init__(self, n_clusters=2, metric="euclidean", stride=1, n_jobs=None, skip=0):
        super().__init__(metric=metric, n_jobs=n_jobs)
        self.set_params(n_clusters=n_clusters, metric=metric, stride=stride, skip=skip)

    def describe(self):
        return "[Uniform time clustering, k = %i, inp_dim=%i]" % (self.n_clusters, self.data_producer.dimension())

    @staticmethod
    def _idx_to_traj_idx(idx, cumsum):
        prev_len = 0
        for traj_idx, length in enumerate(cumsum):
            if prev_len <= idx < length:
                return traj_idx, idx - prev_len
            prev_len = length
        raise ValueError("Requested index %s was out of bounds [0,%s)" % (idx, cumsum[-1]))

    def _estimate(self, iterable, **kw):
        if not hasattr(iterable, "ra_gather"):
            raise NotImplementedError("uniform time clustering needs a random-access data source (DataInMemory)")
        if self.n_clusters is None:
            total_length = int(sum(iterable.trajectory_lengths(stride=self.stride, skip=self.skip)))
            self.n_clusters = min(int(math.sqrt(total_length)), 5000)
            self.logger.info("The number of cluster centers was not specified, "
                             "using min(sqrt(N), 5000)=%s as n_clusters." % self.n_clusters)
        T = iterable.n_frames_total(stride=self.stride, skip=self.skip)
        if self.n_clusters > T:
            self.n_clusters = T
            self.logger.info("Requested more clusters than there are total data points %i. "
                             "Will do clustering with k = %i" % (T, T))
        next_t = (T // self.n_clusters) // 2                       # first point in the middle of its time segment
        cumsum = np.cumsum(iterable.trajectory_lengths(skip=self.skip))
        linspace = self.stride * np.arange(next_t, T - next_t + 1, (T - 2 * next_t + 1) // self.n_clusters)[:self.n_clusters]
        ra_stride = np.array([self._idx_to_traj_idx(x, cumsum) for x in linspace])
        self.clustercenters = np.asarray(iterable.ra_gather(ra_stride, skip=self.skip), dtype=np.float32)
        assert len(self.clustercenters) == self.n_clusters
        return self
