"""CPU: the oracle against (a) every known-answer test the reference holds for this path that can run
without deeptime/mdtraj, (b) independent fp64 numpy implementations, (c) the committed
golden fixtures, (d) the op-order claim of SURVEY Appendix B."""
import itertools
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- (d) summation order: explicit 4-lane model == the literal `#pragma omp simd` loop --------------
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 7, 10, 45, 64, 256, 900])
def test_omp4_model_equals_pragma_loop(oracle, d):
    rng = np.random.RandomState(d)
    for _ in range(300):
        x = (rng.randn(d) * rng.uniform(0.1, 100)).astype(np.float32)
        y = rng.randn(d).astype(np.float32)
        assert oracle.euclid_sq(x, y).tobytes() == oracle.euclid_sq(x, y, "pragma").tobytes()
    if d <= 3:  # for d<=3 all orders agree (SURVEY finding 0.3)
        assert oracle.euclid_sq(x, y).tobytes() == oracle.euclid_sq(x, y, "seq").tobytes()


# ---- (a) reference known-answer tests ------------------------------------------------------------------
def test_kat_k1_cube_corners(oracle):
    # pyemma/coordinates/clustering/tests/test_kmeans.py:154-166
    X = np.array([[1, 1, 1], [1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, 1], [-1, -1, 1], [-1, 1, -1],
                  [1, -1, 1]], np.float32)
    for seed in range(5):
        c0 = oracle.kmpp_init(X, 1, seed)
        cen, code, it, _ = oracle.cluster_loop(X, c0, 10, 1e-5)
        np.testing.assert_equal(cen.squeeze(), [0, 0, 0])


def _truncated_octahedron(scale=0.7071, seed=3):
    """24 vertices of a truncated octahedron (all permutations of (0, +-1, +-2)), scaled and rotated: the kind of point
    set test_kmeans.py:181-233 uses (24 vertices of a convex polytope, k=1)."""
    pts = set()
    for perm in itertools.permutations((0.0, 1.0, 2.0)):
        for s1 in (1.0, -1.0):
            for s2 in (1.0, -1.0):
                v = [c * (s1 if c == 1.0 else (s2 if c == 2.0 else 1.0)) for c in perm]
                pts.add(tuple(v))
    P = np.array(sorted(pts), np.float64) * scale
    q, _ = np.linalg.qr(np.random.RandomState(seed).randn(3, 3))
    return (P @ q.T).astype(np.float32)


def hull_inequalities(points):
    from scipy.spatial import ConvexHull
    return ConvexHull(points.astype(np.float64)).equations  # rows (a, b): inside <=> a.x + b <= 0


def test_kat_k1_center_inside_convex_hull(oracle):
    # test_kmeans.py:181-233: k=1 on the 24 vertices of a convex polytope -> the center satisfies every facet
    # inequality of the hull (and is the members' mean, here the origin up to fp32 rounding)
    X = _truncated_octahedron()
    assert len(X) == 24
    eq = hull_inequalities(X)
    for seed in range(3):
        c0 = oracle.kmpp_init(X, 1, seed)
        cen, code, it, _ = oracle.cluster_loop(X, c0, 10, 1e-5)
        assert cen.shape == (1, 3)
        assert np.all(eq[:, :3] @ cen[0].astype(np.float64) + eq[:, 3] <= 0.0)
        np.testing.assert_allclose(cen[0], X.astype(np.float64).mean(0), atol=1e-6)


def test_kat_outlier_equilibrium(oracle):
    # test_kmeans.py:168-179
    X = np.array([[1, 1.5, 1], [1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, 1], [-1, -1, 1], [-1, 1, -1],
                  [1, -1, 1]], np.float32)
    c0 = np.array([[2, 0, 0], [-2, 0, 0]], np.float32)
    cen, code, it, _ = oracle.cluster_loop(X, c0, 500, 1e-5)
    assert np.all(np.abs(cen) <= 1)


def test_kat_synthetic_trivial(oracle):
    # test_kmeans.py:411-426: 4 constant blobs -> centers exactly 30/60/90/120
    X = np.zeros((40000, 4), np.float32)
    X[0:10000] = 30.0
    X[10000:20000] = 60.0
    X[20000:30000] = 90.0
    X[30000:] = 120.0
    c0 = oracle.kmpp_init(X, 4, 42)
    cen, code, it, _ = oracle.cluster_loop(X, c0, 10, 1e-5)
    assert sorted(cen[:, 0].tolist()) == [30.0, 60.0, 90.0, 120.0]
    assert (cen == cen[:, :1]).all()


def test_kat_assign_tight_blobs(oracle):
    # clustering/tests/test_assign.py:102-109: 5 tight blobs, label == i // nsample
    rng = np.random.RandomState(0)
    centers = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10], [10, 10, 10]], np.float32)
    X = np.concatenate([c + 0.1 * rng.randn(1000, 3) for c in centers]).astype(np.float32)
    lab = oracle.assign(X, centers, n_threads=2)
    assert (lab == np.arange(5000) // 1000).all()
    np.testing.assert_array_equal(lab, oracle.assign(X, centers, n_threads=1))  # test_assign.py:232-239


def test_kat_regspace_center_order(oracle):
    # clustering/tests/test_cluster_samples.py:41-60: centers in first-appearance order
    trajs = [[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 1, 2], [3, 4, 5], [6, 7, 8]]
    X = np.concatenate([np.asarray(t, np.float32).reshape(-1, 1) for t in trajs])
    cen, idx, full = oracle.regspace(X, 0.5, 1000)
    np.testing.assert_array_equal(cen.ravel(), np.arange(9))
    lab = oracle.assign(X, cen)
    np.testing.assert_array_equal(lab, [0, 1, 2, 3, 4, 5, 6, 7, 8] * 2)


@pytest.mark.parametrize("metric", ["euclidean", "minRMSD"])
def test_kat_regspace_pairwise_dmin_and_threads(oracle, metric):
    # test_regspace.py:59-75 (pairs >= dmin), :77-89 (#states == #centers), :137-141 (thread invariance)
    rng = np.random.RandomState(1)
    X = rng.uniform(-2, 2, size=(2000, 3 if metric == "euclidean" else 9)).astype(np.float32)
    dmin = 0.3 if metric == "euclidean" else 0.9
    cen, idx, full = oracle.regspace(X, dmin, 2000, metric)
    assert len(cen) > 1 and not full
    if metric == "euclidean":
        for a, b in itertools.combinations(cen, 2):
            assert np.linalg.norm(a - b) >= dmin * (1 - 1e-6)
    assert len(np.unique(oracle.assign(X, cen, metric))) == len(cen)
    cen2, _, _ = oracle.regspace(X, dmin, 2000, metric, n_threads=2)
    np.testing.assert_array_equal(cen, cen2)


def test_kat_regspace_max_centers(oracle):
    # test_regspace.py:117-135
    X = np.random.RandomState(2).rand(1000, 3).astype(np.float32)
    cen, idx, full = oracle.regspace(X, 1e-8, 50)
    assert full and len(cen) == 50
    np.testing.assert_array_equal(cen, X[:50])


def test_kat_minrmsd_manual_argmin(oracle):
    # test_kmeans.py:235-252: dtraj == argmin over compute_metric(frame, center)
    X = np.random.RandomState(123).uniform(-50, 50, size=(500, 45)).astype(np.float32)
    C = oracle.kmpp_init(X, 15, 32, "minRMSD")
    lab = oracle.assign(X, C, "minRMSD")
    manual = [int(np.argmin([oracle.compute_metric(f, c, "minRMSD") for c in C])) for f in X[:100]]
    np.testing.assert_array_equal(manual, lab[:100])


def test_kat_minrmsd_invariance(oracle):
    # test_kmeans.py:266-316: rotated+translated noisy copies of templates land in one cluster each
    rng = np.random.RandomState(5)

    def rot_y(theta):
        c, s = np.cos(theta), np.sin(theta)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    n_clusters, n_particles, per = 5, 3, 25
    out = np.zeros((n_clusters * per, 3 * n_particles))
    for i in range(n_clusters):
        base = rng.choice(np.arange(3 * n_particles), size=3 * n_particles).astype(float)
        for n in range(per):
            pos = base + rng.normal(size=base.shape, scale=.1)
            r, t = rot_y(np.pi * rng.rand()), rng.normal(size=3)
            for m in range(n_particles):
                out[i * per + n, 3 * m:3 * m + 3] = r @ pos[3 * m:3 * m + 3] - t
    cc = oracle.kmpp_init(out, n_clusters, 1, "minRMSD")
    lab = oracle.assign(out, cc, "minRMSD")
    for i in range(n_clusters):
        assert len(np.unique(lab[i * per:(i + 1) * per])) == 1


def test_kat_minrmsd_dim_not_multiple_of_3(oracle):
    # clustering_module.cpp:12-14
    with pytest.raises(ValueError):
        oracle.compute_metric(np.zeros(10), np.zeros(10), "minRMSD")


def test_kat_seed_determinism(oracle):
    # test_kmeans.py:93: same seed -> same initial centers
    X = np.random.RandomState(3).randn(3000, 3).astype(np.float32)
    a = oracle.kmpp_init(X, 20, 463498)
    b = oracle.kmpp_init(X, 20, 463498)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a, oracle.kmpp_init(X, 20, 42))


# ---- (b) independent fp64 numpy implementations ------------------------------------------------------
def kabsch_rmsd(a, b):
    a = a.reshape(-1, 3).astype(np.float64)
    b = b.reshape(-1, 3).astype(np.float64)
    a, b = a - a.mean(0), b - b.mean(0)
    U, S, Vt = np.linalg.svd(a.T @ b)
    d = np.sign(np.linalg.det(U @ Vt))
    e = S[0] + S[1] + d * S[2]
    return np.sqrt(max(0.0, ((a ** 2).sum() + (b ** 2).sum() - 2 * e) / len(a)))


def test_qcp_matches_kabsch(oracle):
    rng = np.random.RandomState(0)
    for _ in range(300):
        na = rng.randint(3, 60)
        a = rng.uniform(-5, 5, na * 3).astype(np.float32)
        b = rng.uniform(-5, 5, na * 3).astype(np.float32)
        r, k = float(oracle.compute_metric(a, b, "minRMSD")), kabsch_rmsd(a, b)
        assert abs(r - k) <= 2e-6 * max(k, 1.0)
    # rotated copy -> ~0 (cancellation regime: absolute tolerance only)
    a = rng.uniform(-5, 5, 30 * 3).astype(np.float32)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    b = (a.reshape(-1, 3) @ R.T + 3.0).astype(np.float32).ravel()
    assert float(oracle.compute_metric(a, b, "minRMSD")) < 5e-3


def test_lloyd_matches_numpy_fp64(oracle):
    rng = np.random.RandomState(4)
    X = rng.randn(5000, 7).astype(np.float32)
    C = X[:30].copy()
    newc, lab = oracle.kmeans_cluster(X, C, acc="f64")
    d2 = ((X[:, None, :].astype(np.float64) - C[None].astype(np.float64)) ** 2).sum(-1)
    ref = d2.argmin(1)
    srt = np.sort(d2, 1)
    clear = (srt[:, 1] - srt[:, 0]) > 1e-5 * srt[:, 1]
    np.testing.assert_array_equal(lab[clear], ref[clear])
    for j in range(30):
        m = lab == j
        if m.any():
            np.testing.assert_allclose(newc[j], X[m].astype(np.float64).mean(0), rtol=1e-6, atol=1e-7)
        else:
            np.testing.assert_array_equal(newc[j], C[j])
    c32 = float(oracle.cost(X, newc, lab))
    c64 = float(((X.astype(np.float64) - newc[lab]) ** 2).sum())
    assert abs(c32 - c64) <= 1e-4 * c64


def test_cluster_loop_contract(oracle):
    rng = np.random.RandomState(6)
    X = rng.randn(2000, 2).astype(np.float32)
    C = X[:10].copy()
    calls = []
    cen, code, it, inert = oracle.cluster_loop(X, C, 3, 0.0, callback=lambda: calls.append(1))
    assert (code, it, len(inert), len(calls)) == (1, 3, 3, 3)
    cen, code, it, inert = oracle.cluster_loop(X, C, 0, 1e-5)  # max_iter=0 still runs one step (do-while)
    assert it == 1
    cen, code, it, inert = oracle.cluster_loop(X, C, 500, 1e-5)
    assert code == 0 and it < 500 and (np.diff(inert) <= 1e-3 * inert[0]).all()


def test_kmpp_rng_and_modes(oracle):
    first, u = oracle.rng_stream(42, 1000, 8)
    assert 0 <= first < 1000 and ((0 <= u) & (u <= 1)).all()
    X = np.random.RandomState(8).randn(3000, 2).astype(np.float32)
    calls = []
    c, idx = oracle.kmpp_init(X, 25, 7, scan="serial", return_indices=True, callback=lambda: calls.append(1))
    assert len(calls) == 25 and len(set(idx.tolist())) == 25 and idx[0] == oracle.rng_stream(7, 3000, 1)[0]
    np.testing.assert_array_equal(c, X[idx])
    cb, idxb = oracle.kmpp_init(X, 25, 7, scan="blocked", return_indices=True)
    assert len(set(idxb.tolist())) == 25 and idxb[0] == idx[0]
    # D^2 seeding spreads: potential far below uniform picks
    def pot(C):
        return ((X[:, None] - C[None]) ** 2).sum(-1).min(1).sum()
    uni = X[np.random.RandomState(0).randint(0, 3000, 25)]
    assert pot(c) < pot(uni) and pot(cb) < pot(uni)
    np.testing.assert_array_equal(oracle.kmpp_init(X, 25, 7, n_threads=4), c)  # thread-count invariant
    with pytest.raises(ValueError):
        oracle.kmpp_init(X[:5], 6, 1)


# ---- (c) golden fixtures ------------------------------------------------------------------------------
def test_golden_cfg1(oracle):
    g = np.load(os.path.join(GOLD, "cfg1_small.npz"))
    X = g["X"]
    c, i = oracle.kmpp_init(X, 100, 42, scan="serial", return_indices=True)
    np.testing.assert_array_equal(i, g["kmpp_serial_idx"])
    _, ib = oracle.kmpp_init(X, 100, 42, scan="blocked", return_indices=True)
    np.testing.assert_array_equal(ib, g["kmpp_blocked_idx"])
    cen, code, it, inert = oracle.cluster_loop(X, c, 10, 1e-5)
    np.testing.assert_array_equal(cen, g["centers_f32seq"])
    np.testing.assert_array_equal(inert, g["inertias_f32seq"])
    assert (code, it) == (int(g["code"]), int(g["iters"]))
    np.testing.assert_array_equal(oracle.assign(X, cen, n_threads=4), g["dtraj"])


def test_golden_others(oracle):
    g = np.load(os.path.join(GOLD, "cfg2_small.npz"))
    np.testing.assert_array_equal(oracle.assign(g["X"], g["C"]), g["dtraj"])
    np.testing.assert_array_equal(oracle.kmeans_cluster(g["X"], g["C"])[0], g["newC_f32seq"])
    g = np.load(os.path.join(GOLD, "minrmsd_small.npz"))
    np.testing.assert_array_equal(oracle.assign(g["X"], g["C"], "minRMSD"), g["dtraj"])
    np.testing.assert_array_equal(oracle.pairwise(g["X"][:20], g["C"], "minRMSD"), g["dist"])
    g = np.load(os.path.join(GOLD, "regspace_small.npz"))
    np.testing.assert_array_equal(oracle.regspace(g["X"], 3.0, 500)[1], g["idx_euclid"])
    np.testing.assert_array_equal(oracle.regspace(g["X"], 1.2, 500, "minRMSD")[1], g["idx_rmsd"])
