#!/usr/bin/env python
"""Where the wall time of cluster_kmeans(X, k, max_iter=10) goes (cfg2 shape, pageable host array): every libb2k call
of the estimator timed with a device synchronize on both sides, per call name, with and without the pruned session."""
import collections
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pyemma_b200 as coor  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402


class TimedLib:
    def __init__(self, lib):
        self._lib = lib
        self.t = collections.OrderedDict()
        self.sync = True

    def __getattr__(self, name):
        f = getattr(self._lib, name)
        if not name.startswith("b2k_") or not callable(f):
            return f

        def g(*a):
            if self.sync:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f(*a)
            if self.sync:
                torch.cuda.synchronize()
            e = self.t.setdefault(name, [])
            e.append((time.perf_counter() - t0) * 1e3)
            return r
        return g


n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
bench.select_workload(wl)
k = bench.W["k"]
X = bench.synth_host(n, 3)
C0 = X[:k].copy()
ctx = _lib.context()
raw = ctx.lib
for mode, timed in ((1, False), (0, False), (1, False), (0, False), (1, True), (0, True)):
    ctx.set_option("prune_mode", mode)
    tl = TimedLib(raw)
    ctx.lib = tl if timed else raw
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    km = coor.cluster_kmeans(X, k=k, max_iter=iters, clustercenters=C0.copy(), tolerance=0.0)
    t1 = time.perf_counter()
    dt = km.dtrajs
    t2 = time.perf_counter()
    print("prune_mode=%d timed=%d fit %.1f ms (%d iterations)  dtrajs %.1f ms  inertia %.6g" % (
        mode, timed, (t1 - t0) * 1e3, len(km.inertias_), (t2 - t1) * 1e3, km.inertias_[-1]), flush=True)
    if timed:
        for name, v in tl.t.items():
            print("    %-40s n=%3d total %8.2f ms   %s" % (name, len(v), sum(v), " ".join("%.2f" % x for x in v[:14])))
ctx.lib = raw
