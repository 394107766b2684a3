"""RegularSpaceClustering on the B200 backend.

Mirrors pyemma/coordinates/clustering/regspace.py:40-187 (reference @ 3327f28): dmin >= 0
(:103-109), max_centers >= 0 (:119-125), n_clusters alias of max_centers (:127-133), streaming
partial_fit over chunks (:144-151), MaxCentersReachedException -> warning + NotConvergedWarning
(:153-163) with the centers found so far kept (:166-181).
What changes: deeptime RegularSpace.partial_fit -> libb2k's regspace handle (b2k_regspace_*).
Center discovery is sequential in frame order, so it is NOT sharded: "replicas only" (SURVEY 8e);
the subsequent assignment shards like any other.
"""
import warnings

import numpy as np

from .. import _lib
from .interface import AbstractClustering, NotConvergedWarning

__all__ = ["RegularSpaceClustering"]


class RegularSpaceClustering(AbstractClustering):
    def __init__(self, dmin, max_centers=1000, metric="euclidean", stride=1, n_jobs=None, skip=0):
        super().__init__(metric=metric, n_jobs=n_jobs)
        self._converged = False
        self.set_params(dmin=dmin, metric=metric, max_centers=max_centers, stride=stride, skip=skip)

    def describe(self):
        return "[RegularSpaceClustering dmin=%f, inp_dim=%i]" % (self._dmin, self.data_producer.dimension())

    @property
    def dmin(self):
        """Minimum distance between cluster centers."""
        return self._dmin

    @dmin.setter
    def dmin(self, d):
        d = float(d)
        if d < 0:
            raise ValueError("d has to be positive")
        self._dmin = d

    @property
    def max_centers(self):
        """Cutoff during clustering. If reached no more data is taken into account."""
        return self._max_centers

    @max_centers.setter
    def max_centers(self, value):
        value = int(value)
        if value < 0:
            raise ValueError("max_centers has to be positive")
        self._max_centers = value

    @property
    def n_clusters(self):
        return self.max_centers

    @n_clusters.setter
    def n_clusters(self, val):
        self.max_centers = val

    @property
    def converged(self):
        return self._converged

    def _estimate(self, iterable, **kwargs):
        used_frames = 0
        d = iterable.dimension()
        handle = _lib.RegspaceHandle(d, self.dmin, self.max_centers, self.metric)
        it = iterable.iterator(return_trajindex=False, stride=self.stride, chunk=self.chunksize, skip=self.skip)
        n_frames_total = int(np.sum(iterable.trajectory_lengths(stride=self.stride, skip=self.skip)))
        try:
            with it:
                for X in it:
                    handle.partial_fit(X.astype(np.float32, order="C", copy=False))
                    used_frames += len(X)
            self._converged = True
        except _lib.MaxCentersReachedException:
            self._converged = False
            msg = ("Maximum number of cluster centers reached."
                   " Consider increasing max_centers or choose"
                   " a larger minimum distance, dmin.")
            self.logger.warning(msg)
            warnings.warn(msg)
            used_data = used_frames / float(max(n_frames_total, 1)) * 100.0
            raise NotConvergedWarning("Used data for centers: %.2f%%" % used_data)
        finally:
            # even if not converged, we store the found centers (regspace.py:166-181); note that
            # n_clusters (== max_centers) is overwritten with the number found, like upstream.
            clustercenters = handle.centers().reshape(-1, d)
            handle.close()
            self.clustercenters = clustercenters
            self.n_clusters = len(clustercenters)
            self._estimated = True
            if len(clustercenters) == 1:
                self.logger.warning("Have found only one center according to "
                                    "minimum distance requirement of %f" % self.dmin)
        return self
