"""dtrajs from pageable host arrays (what a PyEMMA user passes): b2k_assign frames/s vs the bounce-copy thread count."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pyemma_b200 import _lib
X = bench.synth_host(10_000_000, 3)
C = X[:1000].copy()
ctx = _lib.context(0)
out = np.empty(len(X), np.int32)
for nt in (1, 2, 4, 6, 8, 12):
    ctx.set_option("host_copy_threads", nt)
    _lib.assign(X[:1_000_000], C, out=out[:1_000_000])
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); _lib.assign(X, C, out=out); ts.append(time.perf_counter() - t0)
    print("host_copy_threads=%d: %.1f ms per 1e7 x 10 pageable frames = %.3g frames/s" % (nt, min(ts) * 1e3, len(X) / min(ts)), flush=True)
