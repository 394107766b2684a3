"""Consumers of discrete trajectories on the GPU.

count_states / visited_set / number_of_states mirror pyemma/util/discrete_trajectories.py:146-225;
count_matrix mirrors deeptime.markov.tools.estimation.count_matrix as the MSM estimators call it
(pyemma/msm/estimators/_msm_estimator_base.py:3,229,332: `count_matrix(dtrajs, lag, sliding=..., sparse_return=...)`).
dtrajs may be numpy arrays (uploaded once, 4 bytes per frame) or int32 CUDA tensors (e.g. the labels a fit left in
HBM); the counting itself runs in libb2k (csrc/dtraj.cu), integer exact.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, staging

__all__ = ["count_states", "visited_set", "number_of_states", "count_matrix"]


def _as_device_list(dtrajs, dev):
    if isinstance(dtrajs, (np.ndarray, torch.Tensor)) and dtrajs.ndim == 1:
        dtrajs = [dtrajs]
    out = []
    for dt in dtrajs:
        if isinstance(dt, torch.Tensor):
            if dt.ndim != 1:
                raise ValueError("discrete trajectories must be 1-dimensional")
            out.append(dt.to(device=dev, dtype=torch.int32).contiguous())
        else:
            a = np.asarray(dt)
            if a.ndim != 1:
                raise ValueError("discrete trajectories must be 1-dimensional")
            if a.size and not np.issubdtype(a.dtype, np.integer):
                raise TypeError("discrete trajectories must hold integers")
            out.append(torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev))
    return out


def _max_state(dl):
    m = -1
    for t in dl:
        if t.numel():
            m = max(m, int(t.max().item()))
    return m


def count_states(dtrajs, ignore_negative=False):
    """histogram of state occurrences, length max+1 (discrete_trajectories.py:146-181)"""
    ctx = _lib.context()
    dev = staging.device(ctx)
    dl = _as_device_list(dtrajs, dev)
    if not ignore_negative:
        for t in dl:
            if t.numel() and int(t.min().item()) < 0:
                raise ValueError("'list' argument must have no negative elements")  # np.bincount's message
    ns = _max_state(dl) + 1
    if ns <= 0:
        return np.zeros(0, dtype=int)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    counts = torch.zeros(ns, dtype=torch.int64, device=dev)
    for t in dl:
        _lib.check(ctx.lib.b2k_dev_count_states(ctx.handle, C.c_void_p(t.data_ptr()), t.numel(), ns,
                                                C.c_void_p(counts.data_ptr())))
    return counts.cpu().numpy().astype(int)


def visited_set(dtrajs):
    """states with at least one count (discrete_trajectories.py:184-199)"""
    hist = count_states(dtrajs)
    return np.argwhere(hist > 0)[:, 0]


def number_of_states(dtrajs, only_used=False):
    """largest state index + 1, or the number of visited states (discrete_trajectories.py:202-225)"""
    if only_used:
        return int((count_states(dtrajs) > 0).sum())
    dev = staging.device()
    return _max_state(_as_device_list(dtrajs, dev)) + 1


def count_matrix(dtrajs, lag, sliding=True, sparse_return=True, nstates=None, return_device=False):
    """C[i, j] = number of transitions i -> j at lag time `lag`, summed over the trajectories.

    sliding=True counts every t, sliding=False only t = 0, lag, 2 lag, ... (deeptime/msmtools count_matrix).
    Returns a scipy csr_matrix (sparse_return=True) or a dense float64 array like the reference, or the int64 CUDA
    tensor itself with return_device=True."""
    lag = int(lag)
    if lag < 1:
        raise ValueError("lag must be a positive integer")
    ctx = _lib.context()
    dev = staging.device(ctx)
    dl = _as_device_list(dtrajs, dev)
    ns = _max_state(dl) + 1
    if nstates is not None:
        if nstates < ns:
            raise ValueError("nstates=%d is smaller than the number of states in the dtrajs (%d)" % (nstates, ns))
        ns = int(nstates)
    ns = max(ns, 1)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    Cm = torch.zeros((ns, ns), dtype=torch.int64, device=dev)
    for t in dl:
        _lib.check(ctx.lib.b2k_dev_count_matrix(ctx.handle, C.c_void_p(t.data_ptr()), t.numel(), ns, lag,
                                                1 if sliding else 0, C.c_void_p(Cm.data_ptr())))
    if return_device:
        return Cm
    dense = Cm.cpu().numpy().astype(np.float64)
    if sparse_return:
        import scipy.sparse
        return scipy.sparse.csr_matrix(dense)
    return dense
