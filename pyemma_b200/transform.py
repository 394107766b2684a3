"""LinearProjection: the TICA / PCA output stage in front of the clustering path.

Mirrors `_transform_array` of pyemma/coordinates/transform/_tica_base.py:117-133 and pca.py:257-265,
    Y = (X - mean) . eigenvectors[:, :dim]      (fp64 model, result cast to float32),
for a model estimated elsewhere (the estimation itself -- covariances, eigen-decomposition -- is out of scope).
It is a data source: wrapping a source with it and handing it to `cluster_kmeans` makes `staging.gather_frames` project
every chunk on the device while it is staged, so the raw features never become resident in HBM and never round-trip
through the host as an (N, dim) array.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, staging
from .data import as_source

__all__ = ["LinearProjection"]


class LinearProjection:
    def __init__(self, data, mean, eigenvectors, dim=None):
        self.raw = as_source(data)
        W = np.asarray(eigenvectors, dtype=np.float64)
        if W.ndim != 2 or W.shape[0] != self.raw.dimension():
            raise ValueError("eigenvectors must be (input dimension, n) with input dimension %d" % self.raw.dimension())
        self._dim = int(W.shape[1] if dim is None else dim)
        if not 1 <= self._dim <= W.shape[1]:
            raise ValueError("dim must be in [1, %d]" % W.shape[1])
        self.eigenvectors = np.ascontiguousarray(W)
        self.mean = None if mean is None else np.ascontiguousarray(np.asarray(mean, dtype=np.float64).reshape(-1))
        if self.mean is not None and len(self.mean) != self.raw.dimension():
            raise ValueError("mean has wrong length")
        self._dev = None
        self._chunksize = None
        self.check_output = getattr(self.raw, "check_output", True)

    # ---- device model -------------------------------------------------------------------------
    def device_model(self, dev):
        if self._dev is None or self._dev[0] != dev:
            Wd = torch.from_numpy(self.eigenvectors).to(dev)
            md = None if self.mean is None else torch.from_numpy(self.mean).to(dev)
            self._dev = (dev, md, Wd)
        return self._dev[1], self._dev[2]

    def project_host_chunk(self, X, out, ctx=None):
        """(n, din) host chunk -> rows of the CUDA tensor `out` (n, dim), through the pinned staging of libb2k"""
        ctx = ctx or _lib.context()
        md, Wd = self.device_model(out.device)
        X = np.asarray(X)
        if X.dtype != np.float32 and md is not None:
            # fp64 (or integer) features: the reference subtracts the mean in fp64 BEFORE anything is rounded
            # (_tica_base.py:130-133).  Casting first would lose ~6e-8 |mean| per value, which whitening eigenvectors
            # amplify for features with a large mean and a small variance -- so the mean leaves on the host in fp64
            # and the device multiplies the fp32 image of the (small) difference.
            X = X.astype(np.float64, copy=False) - self.mean
            md = None
        X = np.require(X, dtype=np.float32, requirements=["C", "A"])
        _lib.check(ctx.lib.b2k_stage_project(ctx.handle, C.c_void_p(X.ctypes.data), X.shape[0], X.shape[1],
                                             C.c_void_p(md.data_ptr()) if md is not None else None,
                                             C.c_void_p(Wd.data_ptr()), Wd.shape[1], self._dim,
                                             C.c_void_p(out.data_ptr())))

    # ---- DataSource surface -------------------------------------------------------------------
    def dimension(self):
        return self._dim

    ndim = property(dimension)

    def output_type(self):
        return np.float32()

    def number_of_trajectories(self, stride=None):
        return self.raw.number_of_trajectories(stride)

    def trajectory_length(self, itraj, stride=1, skip=0):
        return self.raw.trajectory_length(itraj, stride, skip)

    def trajectory_lengths(self, stride=1, skip=0):
        return self.raw.trajectory_lengths(stride, skip)

    def n_frames_total(self, stride=1, skip=0):
        return self.raw.n_frames_total(stride, skip)

    @property
    def chunksize(self):
        return self.raw.chunksize if self._chunksize is None else self._chunksize

    @chunksize.setter
    def chunksize(self, value):
        self._chunksize = None if value is None else int(value)

    def n_chunks(self, chunksize, stride=1, skip=0):
        return self.raw.n_chunks(chunksize, stride, skip)

    def transform(self, X):
        """(n, din) array -> (n, dim) float32 (the reference's `transform`)"""
        X = np.asarray(X)
        dev = staging.device()
        out = torch.empty((len(X), self._dim), dtype=torch.float32, device=dev)
        if len(X):
            self.project_host_chunk(X, out)
        return out.cpu().numpy()

    def get_output(self, stride=1, skip=0, chunk=None):
        return [self.transform(x) for x in self.raw.get_output(stride=stride, skip=skip)]

    def iterator(self, stride=1, skip=0, chunk=None, return_trajindex=True):
        inner = self.raw.iterator(stride=stride, skip=skip, chunk=self.chunksize if chunk is None else chunk,
                                  return_trajindex=True)
        outer = self

        class _It:
            def __init__(s):
                s.pos, s.last_chunk, s.last_chunk_in_traj, s.current_trajindex = 0, False, False, 0

            def __enter__(s):
                return s

            def __exit__(s, *e):
                return False

            def __iter__(s):
                return s

            def __next__(s):
                itraj, X = next(inner)
                s.pos, s.last_chunk, s.last_chunk_in_traj = inner.pos, inner.last_chunk, inner.last_chunk_in_traj
                s.current_trajindex = itraj
                Y = outer.transform(X)
                return (itraj, Y) if return_trajindex else Y

            def n_chunks(s):
                return inner.n_chunks()

        return _It()
