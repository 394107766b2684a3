"""GPU: the SURVEY 8(d) parity protocol at its stated sizes, CUDA path vs the CPU oracle, through the C ABI.

  cfg1  full size (tests/test_gpu_configs.py::test_cfg1_full_size_end_to_end)
  cfg2  FULL 1e7 x 10, k=1000: assign bit-exact vs oracle.assign on all host threads; one full-N Lloyd step vs acc="f64"
  cfg3  1e6-frame contiguous slice at full k=2000, d=64: assign bit-exact + one Lloyd step
  cfg4  1e6-frame slice at full k=5000, d=256: assign bit-exact; k-means++ at FULL k=5000 on 2e5 frames, blocked scan
  cfg5  1e5 frames x 300 atoms: regspace (max_centers=1000 reached) + minRMSD assign, bit-exact
Both sides' times are printed (run with -s); the oracle legs need ~2 minutes on the 16 host threads of the GPU box
(recorded in profiles/r02_fullsize_parity.log).  B2K_TEST_FULL=0 skips this file.
"""
import os
import time

import numpy as np
import pytest

import pyemma_b200 as coor

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B2K_TEST_FULL", "1") == "0", reason="B2K_TEST_FULL=0")]
THREADS = os.cpu_count() or 1


def _timed(label, fn, *a, **kw):
    t0 = time.perf_counter()
    out = fn(*a, **kw)
    dt = time.perf_counter() - t0
    print("[fullsize] %-46s %7.2f s" % (label, dt))
    return out


def _blobs(seed, n, d, nb, spread, sigma, scale=None, positive=False):
    """host-side synthetic mixture, generated in 1e6-frame pieces (fp32 throughout)"""
    rng = np.random.RandomState(seed)
    cen = rng.uniform(-spread, spread, size=(nb, d)).astype(np.float32)
    if positive:
        cen = np.abs(cen) + 0.3
    X = np.empty((n, d), np.float32)
    for a in range(0, n, 1_000_000):
        b = min(n, a + 1_000_000)
        X[a:b] = cen[rng.randint(0, nb, b - a)] + sigma * rng.standard_normal((b - a, d)).astype(np.float32)
    if scale is not None:
        X *= scale.astype(np.float32)
    return X, rng


def _centers_close(a, b, rtol=1e-5):
    assert np.abs(a - b).max() <= rtol * np.abs(b).max(), np.abs(a - b).max() / np.abs(b).max()


def test_cfg2_full_size_vs_oracle(oracle, b2k):
    scale = np.sqrt(np.maximum(1.0 - 0.2 * np.arange(10), 0.05))
    X, rng = _blobs(2, 10_000_000, 10, 20, 1.5, 0.6, scale=scale)
    C0 = X[rng.choice(len(X), 1000, replace=False)].copy()
    got = _timed("cfg2 b2k_assign 1e7 x 10, k=1000", b2k.assign, X, C0)
    ref = _timed("cfg2 oracle.assign (%d threads)" % THREADS, oracle.assign, X, C0, n_threads=THREADS)
    np.testing.assert_array_equal(got, ref)                                    # all 1e7 labels bit-exact
    newc, lab = _timed("cfg2 b2k_kmeans_cluster (Lloyd step)", b2k.kmeans_cluster, X, C0)
    onew, olab = _timed("cfg2 oracle.kmeans_cluster acc=f64", oracle.kmeans_cluster, X, C0, n_threads=THREADS, acc="f64")
    np.testing.assert_array_equal(lab, olab)
    _centers_close(newc, onew)                                                  # <= 1e-5 relative (north star)
    # the estimator's own full-size dtrajs (computed from the resident frames) are the same labels
    km = coor.cluster_kmeans(X, k=1000, max_iter=1, clustercenters=C0.copy(), keep_data=False)
    np.testing.assert_array_equal(km.dtrajs[0], oracle.assign(X, km.clustercenters, n_threads=THREADS))


def test_cfg3_slice_vs_oracle(oracle, b2k):
    X, rng = _blobs(3, 1_000_000, 64, 50, 1.0, 0.3, positive=True)
    C0 = X[rng.choice(len(X), 2000, replace=False)].copy()
    got = _timed("cfg3 b2k_assign 1e6 x 64, k=2000", b2k.assign, X, C0)
    ref = _timed("cfg3 oracle.assign", oracle.assign, X, C0, n_threads=THREADS)
    np.testing.assert_array_equal(got, ref)
    newc, lab = b2k.kmeans_cluster(X, C0)
    onew, olab = _timed("cfg3 oracle.kmeans_cluster acc=f64", oracle.kmeans_cluster, X, C0, n_threads=THREADS, acc="f64")
    np.testing.assert_array_equal(lab, olab)
    _centers_close(newc, onew)
    # second iteration from the updated centers (centers no longer coincide with frames)
    np.testing.assert_array_equal(b2k.assign(X, onew), oracle.assign(X, onew, n_threads=THREADS))


def test_cfg4_slice_assign_vs_oracle(oracle, b2k):
    X, rng = _blobs(4, 1_000_000, 256, 200, 10.0, 1.0)
    C5k = X[rng.choice(len(X), 5000, replace=False)].copy()
    C5k[:2500] += (0.05 * rng.standard_normal((2500, 256))).astype(np.float32)  # near-duplicates: tiny gaps
    got = _timed("cfg4 b2k_assign 1e6 x 256, k=5000", b2k.assign, X, C5k)
    ref = _timed("cfg4 oracle.assign", oracle.assign, X, C5k, n_threads=THREADS)
    np.testing.assert_array_equal(got, ref)


def test_cfg4_kmpp_full_k_vs_oracle(oracle, b2k):
    X, rng = _blobs(44, 200_000, 256, 200, 10.0, 1.0)
    got, gi = _timed("cfg4 b2k k-means++ k=5000 on 2e5 x 256 (blocked)", b2k.kmeans_init_centers_kmpp, X, 5000, 42,
                     scan="blocked", return_indices=True)
    ref, ri = _timed("cfg4 oracle k-means++ k=5000 (blocked)", oracle.kmpp_init, X, 5000, 42, scan="blocked",
                     n_threads=THREADS, return_indices=True)
    np.testing.assert_array_equal(gi, ri)                                       # all 5000 picks identical
    np.testing.assert_array_equal(got, ref)


def _conformations(seed, n, n_atoms, n_templates, noise):
    rng = np.random.RandomState(seed)
    T = rng.uniform(-2, 2, size=(n_templates, n_atoms, 3))
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
                  np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
                  np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1)], 1)
    out = np.empty((n, n_atoms * 3), np.float32)
    for s in range(0, n, 10000):
        e = min(n, s + 10000)
        # the visited part of conformation space grows along the trajectory (frame i draws from the first
        # 1 + i/n * n_templates templates), so regspace keeps finding new centers until late in the stream
        live = 1 + (np.arange(s, e) * (n_templates / float(n))).astype(np.int64)
        P = T[(rng.rand(e - s) * live).astype(np.int64)] + noise * rng.standard_normal((e - s, n_atoms, 3))
        P = np.einsum("nij,nkj->nki", R[s:e], P) + rng.uniform(-5, 5, (e - s, 1, 3))
        out[s:e] = P.reshape(e - s, -1)
    return out


def test_cfg5_regspace_minrmsd_vs_oracle(oracle):
    import warnings
    X = _conformations(5, 100_000, 300, 1200, 0.05)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rs = _timed("cfg5 cluster_regspace minRMSD 1e5 x 900, max 1000", coor.cluster_regspace, X, dmin=0.4,
                    max_centers=1000, metric="minRMSD")
    ref_c, ref_idx, full = _timed("cfg5 oracle.regspace", oracle.regspace, X, 0.4, 1000, "minRMSD", n_threads=THREADS)
    assert full and len(ref_c) == 1000                                          # max_centers reached
    np.testing.assert_array_equal(rs.clustercenters, ref_c)                     # same frames, same order, same bits
    got = _timed("cfg5 dtrajs (minRMSD assign, k=1000)", lambda: rs.dtrajs[0])
    ref = _timed("cfg5 oracle.assign minRMSD", oracle.assign, X, ref_c, "minRMSD", n_threads=THREADS)
    np.testing.assert_array_equal(got, ref)
    # the sweep's other regime: dmin right at the typical template-template RMSD (~2.7), where most decisions are
    # within a few percent of the threshold and the pass completes with a few dozen centers
    for dmin in (2.5, 2.7):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rs2 = coor.cluster_regspace(X[:20000], dmin=dmin, max_centers=1000, metric="minRMSD")
        c2, _, full2 = oracle.regspace(X[:20000], dmin, 1000, "minRMSD", n_threads=THREADS)
        assert not full2 and len(c2) > 10
        np.testing.assert_array_equal(rs2.clustercenters, c2)
