#!/usr/bin/env python
"""k-means++ timing probe: python tools/kmpp_probe.py N D K [serial|blocked]  (device-resident frames, one GPU)."""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemma_b200 import _lib  # noqa: E402

n, d, k = int(float(sys.argv[1])), int(sys.argv[2]), int(sys.argv[3])
scan = sys.argv[4] if len(sys.argv) > 4 else "blocked"
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(4)
means = torch.randn((200, d), generator=g, device=dev) * 5
X = torch.randn((n, d), generator=g, device=dev) + means[torch.randint(0, 200, (n,), generator=g, device=dev)]
ctx = _lib.context(0)
ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
cen = torch.empty((k, d), dtype=torch.float32, device=dev)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(ctx.lib.b2k_dev_kmeans_init_centers_kmpp(ctx.handle, C.c_void_p(X.data_ptr()), n, d, k, 0, 42,
                                                        _lib.KMPP_SERIAL if scan == "serial" else _lib.KMPP_BLOCKED,
                                                        _lib.CALLBACK(0), None, C.c_void_p(cen.data_ptr()), None))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("kmpp n=%d d=%d k=%d %s: %.3f s, %.3f ms/round, hbm-algorithmic %.1f GB/s" %
          (n, d, k, scan, dt, dt / max(k - 1, 1) * 1e3, (k - 1) * n * (4 * d + 8) / dt / 1e9), flush=True)
