#!/usr/bin/env python
"""Device-resident timing of the assignment engines on one shape (development aid, not the driver bench).

python tools/shape_bench.py --n 10000000 --d 10 --k 1000 [--engine screen|direct|auto] [--blobs 20] [--check]
Prints the CUDA-event time of a Lloyd step (assign+accumulate, finalize, cost) and of a bare assign,
the screen's candidate statistics and (with --check) compares labels of the two engines.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--blobs", type=int, default=20)
    ap.add_argument("--sigma", type=float, default=0.6)
    ap.add_argument("--spread", type=float, default=1.5)
    ap.add_argument("--engine", default="screen")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--lloyd", type=int, default=3, help="Lloyd steps to run first (centers move towards means)")
    a = ap.parse_args()
    import torch
    from pyemma_b200 import _lib
    dev = torch.device("cuda", 0)
    ctx = _lib.context(0)
    lib = ctx.lib
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    means = torch.randn((a.blobs, a.d), generator=g, device=dev) * a.spread
    lab = torch.randint(0, a.blobs, (a.n,), generator=g, device=dev)
    X = torch.randn((a.n, a.d), generator=g, device=dev, dtype=torch.float32)
    X.mul_(a.sigma).add_(means[lab])
    del lab
    cur = X[torch.randperm(a.n, generator=g, device=dev)[:a.k]].clone()
    nxt = torch.empty_like(cur)
    labels = torch.empty(a.n, dtype=torch.int32, device=dev)
    eng = {"auto": 0, "direct": 1, "screen": 2}

    def evt():
        return torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = evt(), evt()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ctx.set_option("assign_engine", eng[a.engine])
    absmax = C.c_float(0)
    _lib.check(lib.b2k_dev_absmax(ctx.handle, C.c_void_p(X.data_ptr()), a.n * a.d, C.byref(absmax)))
    sess = C.c_void_p()
    _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), a.n, a.d, a.k, 0, a.n,
                                        C.c_float(absmax.value), C.byref(sess)))
    acc = torch.zeros(int(lib.b2k_dev_lloyd_acc_len(sess)), dtype=torch.int64, device=dev)

    def lloyd_step():
        nonlocal cur, nxt
        _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), C.c_void_p(labels.data_ptr()),
                                                       C.c_void_p(acc.data_ptr())))
        _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                              C.c_void_p(nxt.data_ptr())))
        _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), C.c_void_p(labels.data_ptr()),
                                          C.c_void_p(acc.data_ptr())))
        cur, nxt = nxt, cur

    for _ in range(a.lloyd):
        lloyd_step()
    torch.cuda.synchronize()
    out = {"n": a.n, "d": a.d, "k": a.k, "engine": a.engine}
    out["lloyd_step_ms"] = timeit(lloyd_step, a.reps)
    lib.b2k_dev_lloyd_destroy(sess)

    def assign():
        _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), a.n, a.d, C.c_void_p(cur.data_ptr()), a.k, 0,
                                      C.c_void_p(labels.data_ptr()), None))
    out["assign_ms"] = timeit(assign, a.reps)
    if a.engine != "direct":
        fr = ctx.get_stat("screen_frames")
        if fr:
            out["cand_chunks_per_frame"] = ctx.get_stat("screen_cand_chunks") / fr
            out["fallback_frac"] = ctx.get_stat("screen_fallback_frames") / fr
    out["frames_per_s_lloyd"] = a.n / (out["lloyd_step_ms"] * 1e-3)
    out["tflops_alg_assign"] = 2.0 * a.k * a.d * a.n / (out["assign_ms"] * 1e-3) / 1e12
    if a.check:
        ref = torch.empty_like(labels)
        ctx.set_option("assign_engine", 1)
        _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), a.n, a.d, C.c_void_p(cur.data_ptr()), a.k, 0,
                                      C.c_void_p(ref.data_ptr()), None))
        torch.cuda.synchronize()
        out["mismatch_vs_direct"] = int((ref != labels).sum().item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
