"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this module.  pyemma_b200/ never does (tests/test_cabi.py::test_product_never_references_oracle
enforces it).  See oracle.cpp for what is restated and its parity status
("parity unpinned upstream": deeptime / mdtraj are absent offline).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

EUCLIDEAN, MINRMSD = 0, 1
_METRICS = {"euclidean": EUCLIDEAN, "minRMSD": MINRMSD}


def build(force=False):
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_CB = C.CFUNCTYPE(None, C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        f32p, i32p, i64p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_euclid_sq.restype = C.c_float
        L.orc_euclid_sq.argtypes = [f32p, f32p, C.c_int64]
        L.orc_euclid_sq_pragma.restype = C.c_float
        L.orc_euclid_sq_pragma.argtypes = [f32p, f32p, C.c_int64]
        L.orc_euclid_sq_seq.restype = C.c_float
        L.orc_euclid_sq_seq.argtypes = [f32p, f32p, C.c_int64]
        L.orc_compute_metric.restype = C.c_float
        L.orc_compute_metric.argtypes = [f32p, f32p, C.c_int64, C.c_int, C.POINTER(C.c_int)]
        L.orc_center_and_trace.argtypes = [f32p, C.c_int, f32p]
        L.orc_assign.argtypes = [f32p, C.c_int64, C.c_int64, f32p, C.c_int64, C.c_int, C.c_int, i32p]
        L.orc_kmeans_cluster.argtypes = [f32p, C.c_int64, C.c_int64, f32p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                         f32p, i32p]
        L.orc_cost.restype = C.c_float
        L.orc_cost.argtypes = [f32p, C.c_int64, C.c_int64, f32p, C.c_int64, i32p, C.c_int, C.c_int, C.c_int]
        L.orc_cluster_loop.argtypes = [f32p, C.c_int64, C.c_int64, f32p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_int, _CB, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       f32p, C.c_int, f32p, i32p]
        L.orc_kmpp_init.argtypes = [f32p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int,
                                    f32p, i64p, _CB, C.c_void_p]
        L.orc_rng_stream.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_uint64), f32p, C.c_int]
        L.orc_regspace.argtypes = [f32p, C.c_int64, C.c_int64, f32p, i64p, C.c_float, C.c_int64, C.c_int, C.c_int,
                                   i64p]
        L.orc_pairwise.argtypes = [f32p, C.c_int64, C.c_int64, f32p, C.c_int64, C.c_int, f32p]
        L.orc_build_info.restype = C.c_char_p
        _lib = L
    return _lib


def _f32(a):
    return np.require(a, dtype=np.float32, requirements=["C", "A"])


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _metric(m):
    return _METRICS[m] if isinstance(m, str) else int(m)


def _check(rc):
    if rc == 3:
        raise ValueError("RMSDMetric is only implemented for input data with a dimension divisible by 3.")
    if rc == 2:
        raise ValueError("invalid argument")
    if rc not in (0, 4):
        raise RuntimeError("oracle rc=%d" % rc)
    return rc


def euclid_sq(x, y, variant="omp4"):
    x, y = _f32(x), _f32(y)
    f = {"omp4": lib().orc_euclid_sq, "pragma": lib().orc_euclid_sq_pragma, "seq": lib().orc_euclid_sq_seq}[variant]
    return np.float32(f(_p(x), _p(y), x.size))


def compute_metric(x, y, metric="euclidean"):
    x, y = _f32(x).ravel(), _f32(y).ravel()
    err = C.c_int(0)
    v = lib().orc_compute_metric(_p(x), _p(y), x.size, _metric(metric), C.byref(err))
    if err.value:
        _check(3)
    return np.float32(v)


def center_and_trace(c):
    c = _f32(c).ravel().copy()
    tr = np.zeros(1, np.float32)
    lib().orc_center_and_trace(_p(c), c.size // 3, _p(tr))
    return c, tr[0]


def assign(X, centers, metric="euclidean", n_threads=1):
    X, centers = _f32(X), _f32(centers)
    n, d = X.shape
    out = np.empty(n, np.int32)
    _check(lib().orc_assign(_p(X), n, d, _p(centers), centers.shape[0], _metric(metric), n_threads,
                            _p(out, C.c_int32)))
    return out


def kmeans_cluster(X, centers, metric="euclidean", n_threads=1, acc="f32seq"):
    X, centers = _f32(X), _f32(centers)
    n, d = X.shape
    newc = np.empty_like(centers)
    labels = np.empty(n, np.int32)
    _check(lib().orc_kmeans_cluster(_p(X), n, d, _p(centers), centers.shape[0], _metric(metric), n_threads,
                                    0 if acc == "f32seq" else 1, _p(newc), _p(labels, C.c_int32)))
    return newc, labels


def cost(X, centers, labels, metric="euclidean", n_threads=1, acc="f32seq"):
    X, centers = _f32(X), _f32(centers)
    labels = np.require(labels, np.int32, ["C"])
    n, d = X.shape
    return np.float32(lib().orc_cost(_p(X), n, d, _p(centers), centers.shape[0], _p(labels, C.c_int32),
                                     _metric(metric), n_threads, 0 if acc == "f32seq" else 1))


def cluster_loop(X, centers, max_iter, tolerance, metric="euclidean", n_threads=1, acc="f32seq", callback=None,
                 history=False):
    """-> (centers, code, iters, inertias[, centers_hist, labels])"""
    X = _f32(X)
    cen = _f32(centers).copy()
    n, d = X.shape
    k = cen.shape[0]
    cap = max(int(max_iter), 1)
    inert = np.zeros(cap, np.float32)
    hist = np.zeros((cap, k, d), np.float32) if history else None
    labels = np.empty(n, np.int32)
    code, iters = C.c_int(0), C.c_int(0)
    cb = _CB(lambda _u: callback()) if callback is not None else _CB(0)
    _check(lib().orc_cluster_loop(_p(X), n, d, _p(cen), k, _metric(metric), n_threads, int(max_iter),
                                  C.c_float(tolerance), 0 if acc == "f32seq" else 1, cb, None, C.byref(code),
                                  C.byref(iters), _p(inert), cap, _p(hist) if history else None,
                                  _p(labels, C.c_int32)))
    res = (cen, code.value, iters.value, inert[:iters.value].copy())
    if history:
        res = res + (hist[:iters.value].copy(), labels)
    return res


def kmpp_init(X, k, seed, metric="euclidean", n_threads=1, scan="serial", callback=None, return_indices=False):
    X = _f32(X)
    n, d = X.shape
    cen = np.zeros((k, d), np.float32)
    chosen = np.full(k, -1, np.int64)
    cb = _CB(lambda _u: callback()) if callback is not None else _CB(0)
    _check(lib().orc_kmpp_init(_p(X), n, d, k, _metric(metric), int(seed), n_threads, 0 if scan == "serial" else 1,
                               _p(cen), _p(chosen, C.c_int64), cb, None))
    return (cen, chosen) if return_indices else cen


def rng_stream(seed, n, count):
    first = C.c_uint64(0)
    u = np.zeros(count, np.float32)
    lib().orc_rng_stream(int(seed), int(n), C.byref(first), _p(u), count)
    return first.value, u


def regspace(X, dmin, max_centers, metric="euclidean", n_threads=1, centers=None):
    """-> (centers, frame_indices, max_centers_reached)"""
    X = _f32(X)
    n, d = X.shape
    buf = np.zeros((max_centers, d), np.float32)
    nc = C.c_int64(0)
    if centers is not None and len(centers):
        centers = _f32(centers)
        buf[:len(centers)] = centers
        nc = C.c_int64(len(centers))
    idx = np.full(max_centers, -1, np.int64)
    rc = _check(lib().orc_regspace(_p(X), n, d, _p(buf), C.byref(nc), C.c_float(dmin), max_centers, _metric(metric),
                                   n_threads, _p(idx, C.c_int64)))
    return buf[:nc.value].copy(), idx[:nc.value].copy(), rc == 4


def pairwise(X, centers, metric="euclidean"):
    X, centers = _f32(X), _f32(centers)
    out = np.empty((X.shape[0], centers.shape[0]), np.float32)
    _check(lib().orc_pairwise(_p(X), X.shape[0], X.shape[1], _p(centers), centers.shape[0], _metric(metric),
                              _p(out)))
    return out


def build_info():
    return lib().orc_build_info().decode()
