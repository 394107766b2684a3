#!/usr/bin/env python
"""minRMSD assign timing: one launch over all frames against pieces of `piece` frames (same kernel, same data)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402

n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000, 1000
templates = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
dev = torch.device("cuda", 0)
ctx = _lib.context(0)
lib = ctx.lib
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
Xh = bench.conformations(n, 300, templates, 5)
X = torch.from_numpy(Xh).to(dev)
cen = X[torch.randperm(n, device=dev)[:k]].clone()
lab = torch.empty(n, dtype=torch.int32, device=dev)


def run(piece):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a in range(0, n, piece):
        m = min(piece, n - a)
        _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr() + a * 3600), m, 900, C.c_void_p(cen.data_ptr()), k, 1,
                                      C.c_void_p(lab.data_ptr() + a * 4), None))
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


pieces = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [n, 100_000, 18641, 9472]
for piece in pieces:
    print("templates %d  n %d  piece %7d  ms per pass:" % (templates, n, piece), " ".join("%.1f" % run(piece) for _ in range(6)), flush=True)
# regspace centers (what the cfg5 workload assigns against) with and without the early abandon
h = _lib.RegspaceHandle(900, 2.5, k, "minRMSD", ctx)
try:
    h.partial_fit_dev(X.data_ptr(), n)
except _lib.MaxCentersReachedException:
    pass
cen = torch.from_numpy(h.centers()).to(dev)
k = cen.shape[0]
h.close()
for ab in (0, 1):
    ctx.set_option("rmsd_abandon", ab)
    run(n)
    print("regspace centers k=%d abandon=%d  %.1f ms" % (k, ab, run(n)))
    ref = lab.clone() if ab == 0 else ref
print("labels identical:", bool(torch.equal(ref, lab)))
