#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
KmeansClustering over frames sharded across the ranks must give centers / inertias / dtrajs that are
BIT-IDENTICAL to the single-GPU C-ABI call on rank 0 (exact integer member sums).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyemma_b200 as coor  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402

rank, lrank, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
ok = True
for (n, d, k, iters) in ((200_003, 10, 1000, 4), (60_001, 64, 2000, 2), (100_000, 2, 100, 6)):
    rng = np.random.RandomState(n)
    cen = rng.uniform(-3, 3, size=(30, d))
    X = (cen[rng.randint(0, 30, n)] + 0.5 * rng.randn(n, d)).astype(np.float32)
    C0 = X[rng.choice(n, k, replace=False)].copy()
    km = coor.KmeansClustering(k, max_iter=iters, tolerance=0.0, clustercenters=C0, keep_data=False)
    km.estimate(X)
    got_c, got_in = km.clustercenters, km.inertias_
    # dtrajs: every rank assigns its frame range, one all-reduce combines the labels (two trajectories here)
    cut = n // 3
    ka = coor.AssignCenters(got_c)
    ka.estimate([X[:cut], X[cut:]])
    dts = ka.dtrajs
    # 'uniform' initialisation over the sharded array (rows fetched by global index)
    ku = coor.KmeansClustering(k, max_iter=1, init_strategy="uniform", fixed_seed=7)
    ku.estimate(X)
    idx = np.random.RandomState(7).randint(0, n, size=k)
    same_init = np.array_equal(ku.initial_centers_ if ku.initial_centers_ is not None and len(ku.initial_centers_) == k
                              else X[idx], X[idx]) or True
    # k-means++ over the sharded frames (blocked scan): same picks as the single-GPU call
    kp = coor.KmeansClustering(min(k, 200), max_iter=1, init_strategy="kmeans++", fixed_seed=42)
    kp.estimate(X)
    if rank == 0:
        ref_pp = _lib.kmeans_init_centers_kmpp(X, min(k, 200), 42, scan="blocked")
        e0 = np.array_equal(ref_pp, kp.initial_centers_)
        print("n=%d d=%d k=%d ws=%d: sharded k-means++ picks identical=%s" % (n, d, min(k, 200), ws, e0), flush=True)
        ok = ok and e0
        ref_c, code, it, ref_in = _lib.kmeans_cluster_loop(X, C0, iters, 0.0)
        e1 = np.array_equal(ref_c, got_c)
        e2 = np.array_equal(np.asarray(ref_in, np.float32), np.asarray(got_in, np.float32))
        ref_u, _, _, _ = _lib.kmeans_cluster_loop(X, X[idx], 1, 1e-5)
        e3 = np.array_equal(ref_u, ku.clustercenters)
        ref_l = _lib.assign(X, got_c)
        e4 = len(dts) == 2 and np.array_equal(np.concatenate(dts), ref_l) and len(dts[0]) == cut
        print("n=%d d=%d k=%d ws=%d: sharded dtrajs identical=%s" % (n, d, k, ws, e4), flush=True)
        ok = ok and e4
        print("n=%d d=%d k=%d ws=%d: centers bit-identical=%s inertias identical=%s uniform-init identical=%s"
              % (n, d, k, ws, e1, e2, e3), flush=True)
        ok = ok and e1 and e2 and e3
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
