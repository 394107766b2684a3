"""ctypes binding of libb2k.so (C ABI declared in include/b2k.h).

This is the only place the Python layer touches native code.  There is no CPU fallback:
if the library is missing or no sm_100 device is usable, every call raises.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2k.so")

OK, ERR_INVALID_ARG, ERR_DIM_NOT_MULT3, ERR_MAX_CENTERS, ERR_CUDA, ERR_NOMEM, ERR_NONFINITE = 0, 2, 3, 4, 5, 6, 7
EUCLIDEAN, MINRMSD = 0, 1
KMPP_SERIAL, KMPP_BLOCKED = 0, 1
ENGINE_AUTO, ENGINE_DIRECT, ENGINE_SCREEN = 0, 1, 2
METRICS = {"euclidean": EUCLIDEAN, "minRMSD": MINRMSD}

CALLBACK = C.CFUNCTYPE(None, C.c_void_p)
EXCHANGE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_int)  # b2k_exchange_fn

# every symbol include/b2k.h declares (tests/test_cabi_symbols.py checks the .so exports all of them)
SYMBOLS = [
    "b2k_last_error", "b2k_version", "b2k_launch_count", "b2k_ctx_create", "b2k_ctx_destroy", "b2k_ctx_set_stream",
    "b2k_ctx_sync", "b2k_ctx_set_option", "b2k_ctx_get_stat", "b2k_compute_metric", "b2k_assign", "b2k_dev_assign",
    "b2k_stage_assign", "b2k_dev_lloyd_accumulate",
    "b2k_kmeans_cluster", "b2k_kmeans_cost", "b2k_kmeans_cluster_loop", "b2k_kmeans_init_centers_kmpp",
    "b2k_dev_lloyd_create", "b2k_dev_lloyd_destroy", "b2k_dev_lloyd_acc_len", "b2k_dev_lloyd_assign_accumulate",
    "b2k_dev_lloyd_finalize", "b2k_dev_lloyd_cost", "b2k_dev_lloyd_decode_cost", "b2k_dev_absmax",
    "b2k_dev_all_finite", "b2k_dev_kmeans_cluster_loop", "b2k_dev_kmeans_init_centers_kmpp", "b2k_regspace_create",
    "b2k_regspace_destroy", "b2k_regspace_partial_fit", "b2k_dev_regspace_partial_fit", "b2k_regspace_n_centers",
    "b2k_regspace_get_centers", "b2k_regspace_cluster", "b2k_kmpp_exchange_floats",
    "b2k_dev_kmeans_init_centers_kmpp_sharded", "b2k_dev_count_states", "b2k_dev_count_matrix",
    "b2k_stage_lloyd_assign_accumulate", "b2k_dev_project", "b2k_stage_project", "b2k_upload",
    "b2k_stage_lloyd_pass", "b2k_dev_lloyd_get_labels",
]


class MaxCentersReachedException(Exception):
    """Same name as the exception deeptime raises (matched by class NAME at
    pyemma/coordinates/clustering/regspace.py:154)."""


class InvalidDataInStreamException(Exception):
    """pyemma/coordinates/data/_base/datasource.py:1175"""


_lib = None
_lock = threading.Lock()


def load():
    """Load libb2k.so (no GPU needed for loading; compute calls need one)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "pyemma_b200: native library %s is missing -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
        ip = C.POINTER(C.c_int)
        L.b2k_last_error.restype = C.c_char_p
        L.b2k_launch_count.restype = i64
        L.b2k_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.b2k_ctx_destroy.argtypes = [vp]
        L.b2k_ctx_set_stream.argtypes = [vp, vp]
        L.b2k_ctx_sync.argtypes = [vp]
        L.b2k_ctx_set_option.argtypes = [vp, C.c_char_p, i64]
        L.b2k_ctx_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
        L.b2k_compute_metric.argtypes = [vp, vp, vp, i64, C.c_int, C.POINTER(f32)]
        L.b2k_assign.argtypes = [vp, vp, i64, i32, vp, i32, C.c_int, vp]
        L.b2k_dev_assign.argtypes = [vp, vp, i64, i32, vp, i32, C.c_int, vp, vp]
        L.b2k_kmeans_cluster.argtypes = [vp, vp, i64, i32, vp, i32, C.c_int, vp, vp]
        L.b2k_kmeans_cost.argtypes = [vp, vp, i64, i32, vp, i32, vp, C.c_int, C.POINTER(f32)]
        L.b2k_kmeans_cluster_loop.argtypes = [vp, vp, i64, i32, vp, i32, C.c_int, i32, f32, CALLBACK, vp, ip, ip, vp,
                                              i32]
        L.b2k_kmeans_init_centers_kmpp.argtypes = [vp, vp, i64, i32, i32, C.c_int, i64, C.c_int, CALLBACK, vp, vp, vp]
        L.b2k_dev_lloyd_create.argtypes = [vp, vp, i64, i32, i32, C.c_int, i64, f32, C.POINTER(vp)]
        L.b2k_dev_lloyd_destroy.argtypes = [vp]
        L.b2k_dev_lloyd_acc_len.argtypes = [vp]
        L.b2k_dev_lloyd_acc_len.restype = i64
        L.b2k_dev_lloyd_assign_accumulate.argtypes = [vp, vp, vp, vp]
        L.b2k_dev_lloyd_finalize.argtypes = [vp, vp, vp, vp]
        L.b2k_dev_lloyd_get_labels.argtypes = [vp, vp]
        L.b2k_dev_lloyd_accumulate.argtypes = [vp, vp, vp]
        L.b2k_stage_assign.argtypes = [vp, vp, i64, i32, vp, i32, C.c_int, C.c_int, vp, vp, vp]
        L.b2k_dev_lloyd_cost.argtypes = [vp, vp, vp, vp]
        L.b2k_dev_lloyd_decode_cost.argtypes = [vp, i64]
        L.b2k_dev_lloyd_decode_cost.restype = C.c_double
        L.b2k_dev_absmax.argtypes = [vp, vp, i64, C.POINTER(f32)]
        L.b2k_dev_all_finite.argtypes = [vp, vp, i64, ip]
        L.b2k_dev_kmeans_cluster_loop.argtypes = [vp, vp, i64, i32, vp, i32, C.c_int, i32, f32, CALLBACK, vp, ip, ip,
                                                  vp, i32, vp]
        L.b2k_dev_kmeans_init_centers_kmpp.argtypes = [vp, vp, i64, i32, i32, C.c_int, i64, C.c_int, CALLBACK, vp, vp,
                                                       vp]
        L.b2k_kmpp_exchange_floats.argtypes = [i64, i32, i32]
        L.b2k_kmpp_exchange_floats.restype = i64
        L.b2k_dev_kmeans_init_centers_kmpp_sharded.argtypes = [vp, vp, i64, i32, i32, C.c_int, i64, i64, i64, vp, i64, vp,
                                                               EXCHANGE, vp, CALLBACK, vp, vp, vp]
        L.b2k_stage_lloyd_assign_accumulate.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b2k_stage_lloyd_pass.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
        L.b2k_upload.argtypes = [vp, vp, vp, i64]
        L.b2k_dev_project.argtypes = [vp, vp, i64, i32, vp, vp, i32, i32, vp]
        L.b2k_stage_project.argtypes = [vp, vp, i64, i32, vp, vp, i32, i32, vp]
        L.b2k_dev_count_states.argtypes = [vp, vp, i64, i32, vp]
        L.b2k_dev_count_matrix.argtypes = [vp, vp, i64, i32, i64, C.c_int, vp]
        L.b2k_regspace_create.argtypes = [vp, i32, f32, i64, C.c_int, C.POINTER(vp)]
        L.b2k_regspace_destroy.argtypes = [vp]
        L.b2k_regspace_partial_fit.argtypes = [vp, vp, i64]
        L.b2k_dev_regspace_partial_fit.argtypes = [vp, vp, i64]
        L.b2k_regspace_n_centers.argtypes = [vp]
        L.b2k_regspace_n_centers.restype = i64
        L.b2k_regspace_get_centers.argtypes = [vp, vp]
        L.b2k_regspace_cluster.argtypes = [vp, vp, i64, i32, vp, C.POINTER(i64), f32, i64, C.c_int]
        _lib = L
        return L


def last_error():
    return load().b2k_last_error().decode(errors="replace")


def check(rc):
    """Map a status code to the exception class the reference raises for the same condition."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_INVALID_ARG:
        raise ValueError(msg)
    if rc == ERR_DIM_NOT_MULT3:
        raise ValueError(msg)  # pybind11 translates std::range_error (clustering_module.cpp:13) to ValueError
    if rc == ERR_MAX_CENTERS:
        raise MaxCentersReachedException(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ERR_NONFINITE:
        raise InvalidDataInStreamException(msg)
    raise RuntimeError("libb2k error %d: %s" % (rc, msg))


def metric_id(metric):
    try:
        return METRICS[metric]
    except KeyError:
        raise ValueError("metric '%s' is not registered; available: %s" % (metric, sorted(METRICS)))


def launch_count():
    return int(load().b2k_launch_count())


class Context:
    """Owns one b2k_ctx (streams, pinned staging slots, scratch) on one device."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.b2k_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.b2k_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(self.lib.b2k_ctx_set_stream(self.handle, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        check(self.lib.b2k_ctx_sync(self.handle))

    def set_option(self, name, value):
        check(self.lib.b2k_ctx_set_option(self.handle, name.encode(), int(value)))

    def get_stat(self, name):
        v = C.c_double(0)
        check(self.lib.b2k_ctx_get_stat(self.handle, name.encode(), C.byref(v)))
        return v.value


_contexts = {}


def context(device=None):
    """Process-wide context per device (LOCAL_RANK picks the device under torchrun)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = _contexts.get(device)
    if ctx is None or ctx.handle is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def _f32c(a):
    return np.require(a, dtype=np.float32, requirements=["C", "A"])


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


# ---- host-array convenience wrappers (numpy in / numpy out) ---------------------------------------
def compute_metric(x, y, metric="euclidean", ctx=None):
    ctx = ctx or context()
    x, y = _f32c(x).ravel(), _f32c(y).ravel()
    if x.size != y.size:
        raise ValueError("compute_metric: size mismatch")
    out = C.c_float(0)
    check(ctx.lib.b2k_compute_metric(ctx.handle, _ptr(x), _ptr(y), x.size, metric_id(metric), C.byref(out)))
    return np.float32(out.value)


def _check_2d(X, centers):
    if X.ndim != 2 or centers.ndim != 2:
        raise ValueError("input data and centers must be 2-dimensional")
    if X.shape[1] != centers.shape[1]:
        raise ValueError("dimension mismatch: data has %d, centers have %d" % (X.shape[1], centers.shape[1]))


def assign(X, centers, metric="euclidean", ctx=None, out=None):
    X, centers = _f32c(X), _f32c(centers)
    _check_2d(X, centers)
    metric_id(metric)
    ctx = ctx or context()
    n, d = X.shape
    labels = out if out is not None else np.empty(n, np.int32)
    check(ctx.lib.b2k_assign(ctx.handle, _ptr(X), n, d, _ptr(centers), centers.shape[0], metric_id(metric),
                             _ptr(labels)))
    return labels


def kmeans_cluster(X, centers, metric="euclidean", ctx=None):
    X, centers = _f32c(X), _f32c(centers)
    _check_2d(X, centers)
    ctx = ctx or context()
    n, d = X.shape
    newc = np.empty_like(centers)
    labels = np.empty(n, np.int32)
    check(ctx.lib.b2k_kmeans_cluster(ctx.handle, _ptr(X), n, d, _ptr(centers), centers.shape[0], metric_id(metric),
                                     _ptr(newc), _ptr(labels)))
    return newc, labels


def kmeans_cost(X, centers, labels, metric="euclidean", ctx=None):
    ctx = ctx or context()
    X, centers = _f32c(X), _f32c(centers)
    _check_2d(X, centers)
    labels = np.require(labels, np.int32, ["C"])
    out = C.c_float(0)
    check(ctx.lib.b2k_kmeans_cost(ctx.handle, _ptr(X), X.shape[0], X.shape[1], _ptr(centers), centers.shape[0],
                                  _ptr(labels), metric_id(metric), C.byref(out)))
    return np.float32(out.value)


def _cb(callback):
    if callback is None:
        return CALLBACK(0), None
    fn = CALLBACK(lambda _u: callback())
    return fn, fn


def kmeans_cluster_loop(X, centers, max_iter, tolerance, metric="euclidean", callback=None, ctx=None):
    """-> (centers, code, iterations, inertias)   code 0 == converged"""
    ctx = ctx or context()
    X = _f32c(X)
    cen = _f32c(centers).copy()
    _check_2d(X, cen)
    cap = max(int(max_iter), 1)
    inert = np.zeros(cap, np.float32)
    code, iters = C.c_int(0), C.c_int(0)
    fn, keep = _cb(callback)
    check(ctx.lib.b2k_kmeans_cluster_loop(ctx.handle, _ptr(X), X.shape[0], X.shape[1], _ptr(cen), cen.shape[0],
                                          metric_id(metric), int(max_iter), C.c_float(tolerance), fn, None,
                                          C.byref(code), C.byref(iters), _ptr(inert), cap))
    return cen, code.value, iters.value, inert[:iters.value].copy()


def kmeans_init_centers_kmpp(X, k, random_seed, metric="euclidean", scan="blocked", callback=None, ctx=None,
                             return_indices=False):
    ctx = ctx or context()
    X = _f32c(X)
    if X.ndim != 2:
        raise ValueError("input data must be 2-dimensional")
    n, d = X.shape
    if k > n:
        raise ValueError("k=%d larger than number of frames %d" % (k, n))
    cen = np.zeros((k, d), np.float32)
    chosen = np.full(k, -1, np.int64)
    fn, keep = _cb(callback)
    mode = KMPP_SERIAL if scan == "serial" else KMPP_BLOCKED
    check(ctx.lib.b2k_kmeans_init_centers_kmpp(ctx.handle, _ptr(X), n, d, int(k), metric_id(metric), int(random_seed),
                                               mode, fn, None, _ptr(cen), _ptr(chosen)))
    return (cen, chosen) if return_indices else cen


class RegspaceHandle:
    def __init__(self, d, dmin, max_centers, metric="euclidean", ctx=None):
        self.ctx = ctx or context()
        self.d = int(d)
        h = C.c_void_p()
        check(self.ctx.lib.b2k_regspace_create(self.ctx.handle, self.d, C.c_float(dmin), int(max_centers),
                                               metric_id(metric), C.byref(h)))
        self.handle = h

    def partial_fit(self, X):
        X = _f32c(X)
        if X.ndim != 2 or X.shape[1] != self.d:
            raise ValueError("regspace: chunk has wrong shape %s" % (X.shape,))
        check(self.ctx.lib.b2k_regspace_partial_fit(self.handle, _ptr(X), X.shape[0]))

    def partial_fit_dev(self, ptr, n):
        check(self.ctx.lib.b2k_dev_regspace_partial_fit(self.handle, C.c_void_p(ptr), int(n)))

    @property
    def n_centers(self):
        return int(self.ctx.lib.b2k_regspace_n_centers(self.handle))

    def centers(self):
        out = np.empty((self.n_centers, self.d), np.float32)
        if out.size:
            check(self.ctx.lib.b2k_regspace_get_centers(self.handle, _ptr(out)))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.b2k_regspace_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
