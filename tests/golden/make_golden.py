"""Generates tests/golden/*.npz from the CPU oracle (oracle/oracle.cpp, n_jobs=1, order=omp4).

The reference's own clustering tests hold no golden vectors and deeptime/mdtraj are
not importable offline, so these fixtures pin the ORACLE (against regressions and against a
different compiler) and give the GPU tests fixed byte-exact targets.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def three_well(n, seed):
    """2-D three-well mixture with Markov switching."""
    rng = np.random.RandomState(seed)
    cen = np.array([[-1.5, 0.0], [0.0, 1.2], [1.5, 0.0]])
    s = np.zeros(n, dtype=int)
    jump = rng.rand(n) < 0.01
    tgt = rng.randint(0, 3, n)
    for i in range(1, n):
        s[i] = tgt[i] if jump[i] else s[i - 1]
    return (cen[s] + 0.35 * rng.randn(n, 2)).astype(np.float32)


def main():
    out = {}
    # cfg1-like (reduced N): kmeans++ (both scan modes) + Lloyd + assign
    X = three_well(20000, 1)
    c_ser, i_ser = O.kmpp_init(X, 100, 42, scan="serial", return_indices=True)
    c_blk, i_blk = O.kmpp_init(X, 100, 42, scan="blocked", return_indices=True)
    cen, code, iters, inert, hist, labels = O.cluster_loop(X, c_ser, 10, 1e-5, acc="f32seq", history=True)
    cen64, code64, iters64, inert64 = O.cluster_loop(X, c_ser, 10, 1e-5, acc="f64")
    np.savez_compressed(os.path.join(HERE, "cfg1_small.npz"), X=X, kmpp_serial_idx=i_ser, kmpp_blocked_idx=i_blk,
                        centers_f32seq=cen, inertias_f32seq=inert, code=code, iters=iters, centers_f64=cen64,
                        inertias_f64=inert64, dtraj=O.assign(X, cen))
    # d=10, k=50 assign (cfg2 shape, reduced)
    rng = np.random.RandomState(2)
    Y = (rng.randn(4000, 10) * np.sqrt(np.maximum(1 - 0.2 * np.arange(10), 0.05))).astype(np.float32)
    C = Y[rng.choice(4000, 50, replace=False)]
    np.savez_compressed(os.path.join(HERE, "cfg2_small.npz"), X=Y, C=C, dtraj=O.assign(Y, C),
                        newC_f32seq=O.kmeans_cluster(Y, C)[0], newC_f64=O.kmeans_cluster(Y, C, acc="f64")[0])
    # minRMSD: reference test shape (tests/test_kmeans.py:235-252): 500 x 45 uniform(-50,50), seed 123
    Z = np.random.RandomState(123).uniform(-50, 50, size=(500, 45)).astype(np.float32)
    CZ = O.kmpp_init(Z, 15, 32, "minRMSD")
    np.savez_compressed(os.path.join(HERE, "minrmsd_small.npz"), X=Z, C=CZ, dtraj=O.assign(Z, CZ, "minRMSD"),
                        dist=O.pairwise(Z[:20], CZ, "minRMSD"))
    # regspace both metrics
    R = (np.random.RandomState(7).randn(3000, 6) * 2).astype(np.float32)
    ce, ie, _ = O.regspace(R, 3.0, 500)
    cr, ir, _ = O.regspace(R, 1.2, 500, "minRMSD")
    np.savez_compressed(os.path.join(HERE, "regspace_small.npz"), X=R, idx_euclid=ie, idx_rmsd=ir)
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write(O.build_info() + "\n")
    print("golden fixtures written")


if __name__ == "__main__":
    main()
