#!/usr/bin/env python
"""CPU baseline table of BASELINE.md section 3: the oracle restatement (oracle/, test infrastructure) timed on the host
cores of the GPU box for the five configs -- cfg1/cfg2 at full size, cfg3/4/5 on stated sub-samples with full k and d,
extrapolated linearly in N (and in rounds x trials for k-means++).  Best of 3 wall-clock runs.
    python tools/cpu_baseline.py [--out gpurun_out/cpu_baseline.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def blobs(rng, n, d, nb, spread, sigma):
    cen = rng.uniform(-spread, spread, size=(nb, d))
    return (cen[rng.randint(0, nb, n)] + sigma * rng.randn(n, d)).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cpu_baseline.json"))
    args = ap.parse_args()
    O.build()
    T = os.cpu_count() or 1
    rows = []
    rng = np.random.RandomState(0)

    def lloyd_row(name, X, k, full_n):
        C0 = X[rng.choice(len(X), k, replace=False)].copy()
        t_assign = best(lambda: O.assign(X, C0, n_threads=T))
        def it():
            newc, lab = O.kmeans_cluster(X, C0, n_threads=T)
            O.cost(X, newc, lab, n_threads=T)
        t_it = best(it)
        f = full_n / len(X)
        rows.append({"config": name, "sample_frames": len(X), "full_frames": full_n, "extrapolated": f != 1.0,
                     "assign_s_full": t_assign * f, "lloyd_iter_s_full": t_it * f,
                     "assign_frames_per_s": len(X) / t_assign, "lloyd_frames_per_s": len(X) / t_it})
        print(json.dumps(rows[-1]), flush=True)

    # cfg1 (full size): k-means++ + 10 Lloyd iterations + assign
    from test_gpu_configs import three_well
    X1 = three_well(100_000, 1)
    t_pp = best(lambda: O.kmpp_init(X1, 100, 42, n_threads=T, scan="blocked"))
    c0 = O.kmpp_init(X1, 100, 42, n_threads=T, scan="blocked")
    t_loop = best(lambda: O.cluster_loop(X1, c0, 10, 1e-5, n_threads=T))
    rows.append({"config": "cfg1 1e5x2 k=100", "kmpp_s": t_pp, "cluster_loop_10_iters_s": t_loop,
                 "fit_s": t_pp + t_loop, "extrapolated": False})
    print(json.dumps(rows[-1]), flush=True)
    lloyd_row("cfg1 1e5x2 k=100", X1, 100, 100_000)
    # cfg2 (full size)
    import bench
    lloyd_row("cfg2 1e7x10 k=1000", bench.synth_host(10_000_000, 99), 1000, 10_000_000)
    # cfg3: N/1000 sub-sample of the 1e8 frames (full k, d)
    lloyd_row("cfg3 1e8x64 k=2000", blobs(rng, 100_000, 64, 50, 1.0, 0.3), 2000, 100_000_000)
    # cfg4: N/1000 sub-sample for assign; k-means++ timed at k=200 on it and scaled by rounds x trials x frames
    X4 = blobs(rng, 20_000, 256, 200, 5.0, 1.0)
    lloyd_row("cfg4 2e7x256 k=5000", X4, 5000, 20_000_000)
    kk = 200
    t = best(lambda: O.kmpp_init(X4, kk, 42, n_threads=T, scan="blocked"), reps=2)
    m_small, m_full = 2 + int(np.log(kk)), 2 + int(np.log(5000))
    scale = (5000 * m_full * 20_000_000) / (kk * m_small * len(X4))
    rows.append({"config": "cfg4 k-means++ k=5000", "sample": "k=%d over %d frames: %.2f s" % (kk, len(X4), t),
                 "kmpp_s_full": t * scale, "extrapolated": True})
    print(json.dumps(rows[-1]), flush=True)
    # cfg5: regspace + minRMSD assign on N/100 frames
    from test_gpu_configs import _conformations
    X5 = _conformations(np.random.RandomState(5), 10_000, 300, 30)
    t_rs = best(lambda: O.regspace(X5, 0.4, 1000, "minRMSD", n_threads=T), reps=2)
    cen = X5[:1000].copy()
    t_as = best(lambda: O.assign(X5, cen, "minRMSD", n_threads=T), reps=2)
    rows.append({"config": "cfg5 1e6x300 atoms", "sample_frames": len(X5), "extrapolated": True,
                 "regspace_dmin0.4_s_full": t_rs * 100, "assign_k1000_s_full": t_as * 100,
                 "assign_pairs_per_s": len(X5) * 1000 / t_as})
    print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"threads": T, "build": O.build_info(), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
