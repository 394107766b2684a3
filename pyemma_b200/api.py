"""User entry points with the reference's signatures.

Mirrors pyemma/coordinates/api.py (reference @ 3327f28): cluster_kmeans :1727-1866,
cluster_regspace :1953-2055, assign_to_centers :2060-2156, _check_old_chunksize_arg :74-99.
"""
import warnings

from .clustering import (AssignCenters, KmeansClustering, MiniBatchKmeansClustering, RegularSpaceClustering,
                         UniformTimeClustering)

__all__ = ["cluster_kmeans", "cluster_mini_batch_kmeans", "cluster_uniform_time", "cluster_regspace",
           "assign_to_centers"]

_NOTSET = object()


def _check_old_chunksize_arg(chunksize, chunk_size_default, **kw):
    """api.py:74-99: the deprecated spelling `chunk_size` still wins when given."""
    chosen = None
    if "chunk_size" in kw:
        chosen = kw.pop("chunk_size")
        warnings.warn('Passing deprecated setting "chunk_size", please use "chunksize" instead.',
                      DeprecationWarning)
    elif chunksize is not chunk_size_default:
        chosen = chunksize
    if kw:
        raise TypeError("unexpected keyword arguments: %s" % sorted(kw))
    return chosen


def cluster_kmeans(data=None, k=None, max_iter=10, tolerance=1e-5, stride=1, metric="euclidean",
                   init_strategy="kmeans++", fixed_seed=False, n_jobs=None, chunksize=None, skip=0, keep_data=False,
                   clustercenters=None, **kwargs):
    """k-means clustering (api.py:1727).  Returns the estimator; if data is given it is estimated."""
    kmpp_scan = kwargs.pop("kmpp_scan", "auto")
    cs = _check_old_chunksize_arg(chunksize, None, **kwargs)
    res = KmeansClustering(n_clusters=k, max_iter=max_iter, metric=metric, tolerance=tolerance,
                           init_strategy=init_strategy, fixed_seed=fixed_seed, n_jobs=n_jobs, skip=skip,
                           keep_data=keep_data, clustercenters=clustercenters, stride=stride, kmpp_scan=kmpp_scan)
    if data is not None:
        res.estimate(data, chunksize=cs)
    else:
        res.chunksize = cs
    return res


def cluster_mini_batch_kmeans(data=None, k=100, max_iter=10, batch_size=0.2, metric="euclidean",
                              init_strategy="kmeans++", n_jobs=None, chunksize=None, skip=0, clustercenters=None,
                              **kwargs):
    """k-means with the mini-batch strategy (api.py:1671-1723)."""
    cs = _check_old_chunksize_arg(chunksize, None, **kwargs)
    res = MiniBatchKmeansClustering(n_clusters=k, max_iter=max_iter, metric=metric, init_strategy=init_strategy,
                                    batch_size=batch_size, n_jobs=n_jobs, skip=skip, clustercenters=clustercenters)
    if data is not None:
        res.estimate(data, chunksize=cs)
    else:
        res.chunksize = cs
    return res


def cluster_uniform_time(data=None, k=None, stride=1, metric="euclidean", n_jobs=None, chunksize=None, skip=0,
                         **kwargs):
    """uniform time clustering (api.py:1870-1949)."""
    cs = _check_old_chunksize_arg(chunksize, None, **kwargs)
    res = UniformTimeClustering(k, metric=metric, n_jobs=n_jobs, skip=skip, stride=stride)
    if data is not None:
        res.estimate(data, chunksize=cs)
    else:
        res.chunksize = cs
    return res


def cluster_regspace(data=None, dmin=-1, max_centers=1000, stride=1, metric="euclidean", n_jobs=None,
                     chunksize=None, skip=0, **kwargs):
    """regular space clustering (api.py:1953)."""
    if dmin == -1:
        raise ValueError("provide a minimum distance for clustering, e.g. 2.0")
    cs = _check_old_chunksize_arg(chunksize, None, **kwargs)
    res = RegularSpaceClustering(dmin, max_centers=max_centers, metric=metric, n_jobs=n_jobs, stride=stride,
                                 skip=skip)
    if data is not None:
        res.estimate(data, chunksize=cs)
    else:
        res.chunksize = cs
    return res


def assign_to_centers(data=None, centers=None, stride=1, return_dtrajs=True, metric="euclidean", n_jobs=None,
                      chunksize=None, skip=0, **kwargs):
    """assign data to given centers (api.py:2060)."""
    if centers is None:
        raise ValueError("You have to provide centers in form of a filename or NumPy array or a reader created "
                         "by source function")
    cs = _check_old_chunksize_arg(chunksize, None, **kwargs)
    res = AssignCenters(centers, metric=metric, n_jobs=n_jobs, skip=skip, stride=stride)
    if data is not None:
        res.estimate(data, chunksize=cs)
        if return_dtrajs:
            return res.dtrajs
    else:
        res.chunksize = cs
    return res
