#!/usr/bin/env python
"""Host-pointer entry points vs raw PCIe copies (development aid)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyemma_b200 import _lib
n, d, k = 10_000_000, 10, 1000
dev = torch.device("cuda", 0)
ctx = _lib.context(0); lib = ctx.lib
import bench
hx = torch.from_numpy(bench.synth_host(n, 7)).pin_memory()
hl = torch.empty(n, dtype=torch.int32).pin_memory()
hc = hx[:k].numpy().copy(); hn = np.empty_like(hc)
dx = torch.empty((n, d), dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D 400MB pinned ms", t(lambda: dx.copy_(hx, non_blocking=True)))
dl = torch.empty(n, dtype=torch.int32, device=dev)
print("D2H 40MB pinned ms", t(lambda: hl.copy_(dl, non_blocking=True)))
import sys as _s
if len(_s.argv) > 1 and _s.argv[1] == "torchstream":
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    print("using torch current stream", torch.cuda.current_stream(dev).cuda_stream)
if len(_s.argv) > 2:  # centers after a few Lloyd steps
    for _ in range(int(_s.argv[2])):
        _lib.check(lib.b2k_kmeans_cluster(ctx.handle, C.c_void_p(hx.data_ptr()), n, d, C.c_void_p(hc.ctypes.data), k, 0, C.c_void_p(hn.ctypes.data), C.c_void_p(hl.data_ptr())))
        hc, hn = hn, hc
for sb in (64,):
    ctx.set_option("stage_bytes", sb << 20)
    ms = t(lambda: _lib.check(lib.b2k_assign(ctx.handle, C.c_void_p(hx.data_ptr()), n, d, C.c_void_p(hc.ctypes.data), k, 0, C.c_void_p(hl.data_ptr()))))
    ms2 = t(lambda: _lib.check(lib.b2k_kmeans_cluster(ctx.handle, C.c_void_p(hx.data_ptr()), n, d, C.c_void_p(hc.ctypes.data), k, 0, C.c_void_p(hn.ctypes.data), C.c_void_p(hl.data_ptr()))))
    print("stage %d MB: b2k_assign %.2f ms, b2k_kmeans_cluster %.2f ms" % (sb, ms, ms2), "groups/frame(last chunk)", ctx.get_stat("screen_cand_chunks") / max(ctx.get_stat("screen_frames"), 1), "fallback", ctx.get_stat("screen_fallback_frames"))
