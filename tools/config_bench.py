#!/usr/bin/env python
"""Per-config timing of the hot path on ONE GPU (BASELINE.json configs 1-5; SURVEY.md 8d shapes).

bench.py carries the driver contract for configs[1]; this tool measures the other shapes the same way
(device-resident inputs, CUDA events, warm-up first) so DESIGN.md / profiles/ can quote them.
    python tools/config_bench.py --cfg 1,2,3,4,5 [--out gpurun_out/configs.json] [--kmpp-full]
Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel split of each config.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyemma_b200 as coor  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = PEAKS.get("hbm_gbs", 6500.0)
TF = PEAKS.get("bf16_tflops_sustained", 1400.0)


def device_blobs(n, d, nb, spread, sigma, seed, positive=False):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    means = torch.randn((nb, d), generator=g, device=DEV) * spread
    if positive:
        means = means.abs() + 0.3
    X = torch.randn((n, d), generator=g, device=DEV, dtype=torch.float32)
    step = 1 << 22
    for a in range(0, n, step):
        lab = torch.randint(0, nb, (min(step, n - a),), generator=g, device=DEV)
        X[a:a + step].mul_(sigma).add_(means[lab])
    return X, g


def timed(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def lloyd_and_assign(ctx, X, k, reps, label):
    """ms per Lloyd iteration (session API, what KmeansClustering runs) and per assign pass."""
    lib = ctx.lib
    n, d = X.shape
    cur = X[torch.randperm(n, device=DEV)[:k]].clone()
    nxt = torch.empty_like(cur)
    absmax = C.c_float(0)
    _lib.check(lib.b2k_dev_absmax(ctx.handle, C.c_void_p(X.data_ptr()), n * d, C.byref(absmax)))
    sess = C.c_void_p()
    _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), n, d, k, 0, n, C.c_float(absmax.value),
                                        C.byref(sess)))
    acc = torch.zeros(int(lib.b2k_dev_lloyd_acc_len(sess)), dtype=torch.int64, device=DEV)
    labels = torch.empty(n, dtype=torch.int32, device=DEV)
    costs = []
    state = {"cur": cur, "nxt": nxt}

    def step():
        c, nx = state["cur"], state["nxt"]
        _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(c.data_ptr()), C.c_void_p(labels.data_ptr()),
                                                       C.c_void_p(acc.data_ptr())))
        _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(c.data_ptr()),
                                              C.c_void_p(nx.data_ptr())))
        _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nx.data_ptr()), C.c_void_p(labels.data_ptr()),
                                          C.c_void_p(acc.data_ptr())))
        costs.append(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[-1].item())))
        state["cur"], state["nxt"] = nx, c

    ms_step = timed(step, reps, warm=2)
    ctx.set_option("profile", 1)
    step()
    torch.cuda.synchronize()
    gl = ctx.get_stat("screen_gemm_launches")
    gemm_ms = ctx.get_stat("screen_gemm_ms_total") if gl else None
    groups = ctx.get_stat("screen_cand_chunks")
    fb = ctx.get_stat("screen_fallback_frames")
    terms_used = ctx.get_stat("screen_terms_used")
    probe = {t: [ctx.get_stat("probe_centers_%d" % t), ctx.get_stat("probe_fallback_%d" % t)] for t in (1, 2)}
    ctx.set_option("profile", 0)
    lib.b2k_dev_lloyd_destroy(sess)

    def assign():
        _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, d, C.c_void_p(state["cur"].data_ptr()), k, 0,
                                      C.c_void_p(labels.data_ptr()), None))
    ms_assign = timed(assign, max(1, reps // 2), warm=1)
    flops = 2.0 * k * d * n
    out = {"cfg": label, "n": n, "d": d, "k": k, "lloyd_ms_per_iter": ms_step, "lloyd_frames_per_s": n / ms_step * 1e3,
           "assign_ms_cold_plan": ms_assign, "assign_frames_per_s": n / ms_assign * 1e3,
           "screen_gemm_ms_per_iter": gemm_ms, "cand_groups_per_frame": groups / n if gl else None,
           "fallback_frames": fb, "terms_used": terms_used, "probe_centers_fallback": probe,
           "tensor_frac_step": flops / (ms_step * 1e-3) / 1e12 / TF,
           "tensor_frac_gemm": (flops / (gemm_ms * 1e-3) / 1e12 / TF) if gemm_ms else None,
           "cost_first_last": [costs[0], costs[-1]]}
    return out


def cfg1(ctx, args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_configs import three_well
    X = three_well(100_000, 1)
    res = {"cfg": "cfg1 1e5x2 k=100 kmeans++ + 10 Lloyd + dtrajs (estimator API, host arrays)"}
    for scan in ("blocked", "serial"):
        coor.cluster_kmeans(X[:20000], k=100, max_iter=2, fixed_seed=42, kmpp_scan=scan)  # warm-up (plans, pinned slots)
        t0 = time.perf_counter()
        km = coor.cluster_kmeans(X, k=100, max_iter=10, fixed_seed=42, kmpp_scan=scan, tolerance=1e-5)
        t1 = time.perf_counter()
        dt = km.dtrajs
        t2 = time.perf_counter()
        res["fit_s_" + scan] = t1 - t0
        res["dtrajs_s_" + scan] = t2 - t1
        res["iters_" + scan] = int(len(km.inertias_))
    Xd = torch.from_numpy(X).to(DEV)
    r = lloyd_and_assign(ctx, Xd, 100, 50, "cfg1 device-resident")
    res.update({"lloyd_ms_per_iter": r["lloyd_ms_per_iter"], "assign_ms": r["assign_ms_cold_plan"],
                "assign_frames_per_s": r["assign_frames_per_s"]})
    return res


def cfg2(ctx, args):
    X, _ = device_blobs(10_000_000, 10, 20, 1.5, 0.6, 2)
    return lloyd_and_assign(ctx, X, 1000, 10, "cfg2 1e7x10 k=1000")


def cfg3(ctx, args):
    X, _ = device_blobs(12_500_000, 64, 50, 1.0, 0.3, 3, positive=True)
    return lloyd_and_assign(ctx, X, 2000, 5, "cfg3 per-GPU shard 1.25e7x64 k=2000")


def cfg4(ctx, args):
    lib = ctx.lib
    n = args.cfg4_frames
    X, _ = device_blobs(n, 256, 200, 5.0, 1.0, 4)
    res = lloyd_and_assign(ctx, X, 5000, 2, "cfg4 %gx256 k=5000" % n)
    # k-means++ (HBM-bound: k rounds over X)
    for nk, kk in (() if args.no_kmpp else ((min(n, 2_000_000), 5000),)) + (((n, 5000),) if args.kmpp_full else ()):
        cen = torch.empty((kk, 256), dtype=torch.float32, device=DEV)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(ctx.handle, C.c_void_p(X.data_ptr()), nk, 256, kk, 0, 42,
                                                        _lib.KMPP_BLOCKED, _lib.CALLBACK(0), None,
                                                        C.c_void_p(cen.data_ptr()), None))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        bytes_alg = float(kk) * nk * (4 * 256 + 8)
        res["kmpp_n%d_k%d_s" % (nk, kk)] = dt
        res["kmpp_n%d_k%d_hbm_frac" % (nk, kk)] = bytes_alg / dt / 1e9 / HBM
    return res


def conformations_device(n, n_atoms, n_templates, seed):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    T = (torch.rand((n_templates, n_atoms, 3), generator=g, device=DEV) * 4 - 2)
    out = torch.empty((n, n_atoms * 3), dtype=torch.float32, device=DEV)
    step = 1 << 16
    for a in range(0, n, step):
        m = min(step, n - a)
        q = torch.randn((m, 4), generator=g, device=DEV)
        q = q / q.norm(dim=1, keepdim=True)
        qa, qb, qc, qd = q.unbind(1)
        R = torch.stack([qa * qa + qb * qb - qc * qc - qd * qd, 2 * (qb * qc - qa * qd), 2 * (qb * qd + qa * qc),
                         2 * (qb * qc + qa * qd), qa * qa - qb * qb + qc * qc - qd * qd, 2 * (qc * qd - qa * qb),
                         2 * (qb * qd - qa * qc), 2 * (qc * qd + qa * qb), qa * qa - qb * qb - qc * qc + qd * qd],
                        dim=1).reshape(m, 3, 3)
        t = torch.randint(0, n_templates, (m,), generator=g, device=DEV)
        conf = T[t] + 0.05 * torch.randn((m, n_atoms, 3), generator=g, device=DEV)
        conf = torch.bmm(conf, R.transpose(1, 2)) + (torch.rand((m, 1, 3), generator=g, device=DEV) * 10 - 5)
        out[a:a + m] = conf.reshape(m, -1)
    return out


def cfg5(ctx, args):
    lib = ctx.lib
    n = args.cfg5_frames
    X = conformations_device(n, 300, 30, 5)
    res = {"cfg": "cfg5 %gx(300 atoms) regspace minRMSD dmin sweep + assign" % n, "sweep": []}
    for dmin in (0.8, 0.4, 0.2, 0.1, 0.05):
        h = _lib.RegspaceHandle(900, dmin, 1000, "minRMSD", ctx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hit_max = False
        try:
            h.partial_fit_dev(X.data_ptr(), n)
        except _lib.MaxCentersReachedException:
            hit_max = True
        ctx.sync()
        dt = time.perf_counter() - t0
        kc = h.n_centers
        cen = torch.from_numpy(h.centers()).to(DEV)
        h.close()
        labels = torch.empty(n, dtype=torch.int32, device=DEV)

        def assign():
            _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, 900, C.c_void_p(cen.data_ptr()), kc, 1,
                                          C.c_void_p(labels.data_ptr()), None))
        ms = timed(assign, 1, warm=1)
        res["sweep"].append({"dmin": dmin, "centers": kc, "max_centers_hit": hit_max, "regspace_s": dt,
                             "assign_ms": ms, "assign_frames_per_s": n / ms * 1e3,
                             "assign_pairs_per_s": n * kc / ms * 1e3,
                             "assign_fp32_tflops": n * kc * 18.0 * 300 / (ms * 1e-3) / 1e12})
    return res


def cfg6(ctx, args):
    """dtraj consumers: count matrix at lag 10 and state histogram of 1e7 metastable labels, k=1000"""
    n, k = 10_000_000, 1000
    g = torch.Generator(device=DEV)
    g.manual_seed(6)
    lab = torch.randint(0, k, (n // 20 + 1,), generator=g, device=DEV, dtype=torch.int32).repeat_interleave(20)[:n].contiguous()
    rnd = torch.randint(0, k, (n,), generator=g, device=DEV, dtype=torch.int32)
    lib = ctx.lib
    res = {"cfg": "dtraj consumers 1e7 labels k=1000"}
    for name, l in (("metastable(dwell 20)", lab), ("uniform random", rnd)):
        Cm = torch.zeros((k, k), dtype=torch.int64, device=DEV)
        hist = torch.zeros(k, dtype=torch.int64, device=DEV)
        ms_c = timed(lambda: _lib.check(lib.b2k_dev_count_matrix(ctx.handle, C.c_void_p(l.data_ptr()), n, k, 10, 1,
                                                                 C.c_void_p(Cm.data_ptr()))), 5, warm=1)
        ms_h = timed(lambda: _lib.check(lib.b2k_dev_count_states(ctx.handle, C.c_void_p(l.data_ptr()), n, k,
                                                                 C.c_void_p(hist.data_ptr()))), 5, warm=1)
        res[name] = {"count_matrix_ms": ms_c, "count_matrix_pairs_per_s": n / ms_c * 1e3,
                     "count_matrix_hbm_frac": 8.0 * n / (ms_c * 1e-3) / 1e9 / HBM,
                     "count_states_ms": ms_h, "count_states_hbm_frac": 4.0 * n / (ms_h * 1e-3) / 1e9 / HBM}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="1,2,3,4,5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--cfg4-frames", type=int, default=20_000_000)
    ap.add_argument("--cfg5-frames", type=int, default=1_000_000)
    ap.add_argument("--kmpp-full", action="store_true")
    ap.add_argument("--no-kmpp", action="store_true")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="b2k_ctx_set_option before the run (experiments; repeatable)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = _lib.context(0)
    ctx.set_stream(torch.cuda.current_stream(DEV).cuda_stream)
    for opt in args.option:
        name, _, val = opt.partition("=")
        ctx.set_option(name, int(val))
    fns = {"1": cfg1, "2": cfg2, "3": cfg3, "4": cfg4, "5": cfg5, "6": cfg6}
    results = []
    for c in args.cfg.split(","):
        t0 = time.perf_counter()
        try:
            r = fns[c](ctx, args)
        except Exception as e:  # keep the other configs
            r = {"cfg": c, "error": repr(e)}
        r["wall_s"] = time.perf_counter() - t0
        print(json.dumps(r), flush=True)
        results.append(r)
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peaks": PEAKS, "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
