#!/usr/bin/env python
"""k-means++ seeding time on resident frames: asynchronous rounds against the synchronous loop."""
import ctypes as C
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402

n, d, k = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (100000, 2, 100)))
dev = torch.device("cuda", 0)
ctx = _lib.context(0)
lib = ctx.lib
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
if d == 2:
    X = torch.from_numpy(bench.three_well(n, 1)).to(dev)
else:
    bench.W = dict(bench.WORKLOADS["cfg4"], name="cfg4", d=d)
    X = bench.synth_device(n, 0, dev)
cen = torch.empty((k, d), dtype=torch.float32, device=dev)
chosen = (C.c_int64 * k)()
res = {}
for mode in (2, 1, 0, 2, 1, 0):
    ctx.set_option("kmpp_async", mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(ctx.handle, C.c_void_p(X.data_ptr()), n, d, k, 0, 42, _lib.KMPP_BLOCKED,
                                                        _lib.CALLBACK(0), None, C.c_void_p(cen.data_ptr()), chosen))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    res[mode] = list(chosen)
    print("async=%d  %.2f ms per seeding (%.1f us per round)  fallbacks so far %d" % (mode, dt * 1e3, dt / k * 1e6, ctx.get_stat("kmpp_async_fallbacks")))
print("picks identical:", res[0] == res[1] == res[2])
# the Lloyd loop that follows the seeding in a fit (cfg1: 10 iterations)
code, iters = C.c_int(0), C.c_int(0)
inert = (C.c_float * 16)()
for mode in (2, 0):
    ctx.set_option("kmpp_async", mode)
    for what in ("seed+loop", "loop only"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            if what == "seed+loop":
                _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(ctx.handle, C.c_void_p(X.data_ptr()), n, d, k, 0, 42, _lib.KMPP_BLOCKED,
                                                                _lib.CALLBACK(0), None, C.c_void_p(cen.data_ptr()), None))
            _lib.check(lib.b2k_dev_kmeans_cluster_loop(ctx.handle, C.c_void_p(X.data_ptr()), n, d, C.c_void_p(cen.data_ptr()), k, 0, 10,
                                                       C.c_float(1e-5), _lib.CALLBACK(0), None, C.byref(code), C.byref(iters), inert, 16, None))
        torch.cuda.synchronize()
        print("async=%d  %-10s %.2f ms (%d iterations)" % (mode, what, (time.perf_counter() - t0) / 3 * 1e3, iters.value))
