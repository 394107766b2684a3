#!/usr/bin/env python
"""Per-iteration timing of a Lloyd session on a bench workload: step ms, listed-screen kernel ms, mean list length.
    python tools/prune_probe.py [--workload cfg2|cfg3|cfg4] [--frames N] [--steps S] [--option name=value ...]"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--option", action="append", default=[])
args = ap.parse_args()
bench.select_workload(args.workload)
n = args.frames or bench.W["n"]
dev = torch.device("cuda", 0)
ctx = _lib.context(0)
lib = ctx.lib
stream = torch.cuda.current_stream(dev)
ctx.set_stream(stream.cuda_stream)
for opt in args.option:
    name, _, val = opt.partition("=")
    ctx.set_option(name, int(val))
X = bench.synth_device(n, 0, dev)
D, K = bench.W["d"], bench.W["k"]
cur = X[:K].clone()
nxt = torch.empty_like(cur)
absmax = C.c_float(0)
_lib.check(lib.b2k_dev_absmax(ctx.handle, C.c_void_p(X.data_ptr()), n * D, C.byref(absmax)))
sess = C.c_void_p()
_lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), n, D, K, 0, n, C.c_float(absmax.value), C.byref(sess)))
acc = torch.zeros(int(lib.b2k_dev_lloyd_acc_len(sess)), dtype=torch.int64, device=dev)
labels = torch.empty(n, dtype=torch.int32, device=dev)
ctx.set_option("profile", 1)
for it in range(args.steps):
    g0 = ctx.get_stat("screen_gemm_ms_total")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), C.c_void_p(labels.data_ptr()), C.c_void_p(acc.data_ptr())))
    _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()), C.c_void_p(nxt.data_ptr())))
    _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), C.c_void_p(labels.data_ptr()), C.c_void_p(acc.data_ptr())))
    cost = lib.b2k_dev_lloyd_decode_cost(sess, int(acc[-1].item()))
    e1.record(stream)
    torch.cuda.synchronize()
    g1 = ctx.get_stat("screen_gemm_ms_total")
    print("iter %2d  step %7.3f ms  screen kernel %7.3f ms  mean list %7.1f  pruned steps %d  sorts %d  groups/frame %.3f  fallback %d  cost %.6g"
          % (it + 1, e0.elapsed_time(e1), g1 - g0, ctx.get_stat("prune_mean_list"), ctx.get_stat("prune_steps"),
             ctx.get_stat("prune_sorts"), ctx.get_stat("screen_cand_chunks") / n, ctx.get_stat("screen_fallback_frames"), cost), flush=True)
    cur, nxt = nxt, cur
lib.b2k_dev_lloyd_destroy(sess)
