"""GPU: the estimator API (cluster_kmeans / cluster_regspace / assign_to_centers) -- the reference's own
clustering tests restated (pyemma/coordinates/clustering/tests/*.py), plus oracle parity end to end."""
import os
import warnings

import numpy as np
import pytest

import pyemma_b200 as coor
from pyemma_b200.data import DataInMemory

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_blobs(rng, n, centers, std):
    X = np.concatenate([c + std * rng.randn(n // len(centers), len(c)) for c in centers])
    return X.astype(np.float32)


# ---- test_kmeans.py ---------------------------------------------------------------------------------
def test_kmeans_api_and_dtraj_dtype():
    X = np.random.RandomState(0).randn(5000, 3)
    km = coor.cluster_kmeans(X, k=10)
    assert km.dtrajs[0].dtype == km.output_type()          # test_kmeans.py:48
    assert km.clustercenters.shape == (10, 3) and km.clustercenters.dtype == np.float32
    assert len(km.dtrajs) == 1 and len(km.dtrajs[0]) == 5000
    assert km.dimension() == 1 and "Kmeans" in km.describe()


@pytest.mark.parametrize("init_strategy", ["uniform", "kmeans++"])
@pytest.mark.parametrize("fixed_seed", [True, 463498])
def test_kmeans_seed_determinism(init_strategy, fixed_seed):
    # test_kmeans.py:75-102
    rng = np.random.RandomState(1)
    X = make_blobs(rng, 3000, [[-2, -2], [2, 2], [-2, 2]], 0.4)
    a = coor.cluster_kmeans(X, k=10, init_strategy=init_strategy, fixed_seed=fixed_seed, max_iter=20, n_jobs=1)
    b = coor.cluster_kmeans(X, k=10, init_strategy=init_strategy, fixed_seed=fixed_seed, max_iter=20, n_jobs=1)
    np.testing.assert_array_equal(a.initial_centers_, b.initial_centers_)
    np.testing.assert_array_equal(a.clustercenters, b.clustercenters)  # deterministic: exact fixed-point sums
    np.testing.assert_array_equal(a.dtrajs[0], b.dtrajs[0])


def test_kmeans_known_answers():
    cube = np.array([[1, 1, 1], [1, 1, -1], [1, -1, -1], [-1, -1, -1], [-1, 1, 1], [-1, -1, 1], [-1, 1, -1],
                     [1, -1, 1]], np.float32)
    km = coor.cluster_kmeans(cube, k=1)                      # test_kmeans.py:154-166
    np.testing.assert_equal(km.clustercenters.squeeze(), [0, 0, 0])
    X = cube.copy()
    X[0, 1] = 1.5
    km = coor.cluster_kmeans(X, k=2, clustercenters=np.array([[2, 0, 0], [-2, 0, 0]]), max_iter=500, n_jobs=1)
    assert np.all(np.abs(km.clustercenters) <= 1)           # test_kmeans.py:168-179
    T = np.zeros((40000, 4))
    for i, v in enumerate((30.0, 60.0, 90.0, 120.0)):
        T[i * 10000:(i + 1) * 10000] = v
    cl = coor.cluster_kmeans(T, k=4)                         # test_kmeans.py:411-426
    assert sorted(cl.clustercenters[:, 0].tolist()) == [30.0, 60.0, 90.0, 120.0]
    from test_oracle_kats import _truncated_octahedron, hull_inequalities
    P = _truncated_octahedron()                              # test_kmeans.py:181-233: k=1 center inside the hull
    eq = hull_inequalities(P)
    c = coor.cluster_kmeans(P, k=1).clustercenters[0].astype(np.float64)
    assert np.all(eq[:, :3] @ c + eq[:, 3] <= 0.0)


def test_kmeans_minrmsd_assignment_manual_argmin(b2k):
    # test_kmeans.py:235-252
    data = np.random.RandomState(123).uniform(-50, 50, size=(500, 3 * 15))
    km = coor.cluster_kmeans([data], 15, metric="minRMSD", max_iter=0, fixed_seed=32, init_strategy="kmeans++",
                             n_jobs=1)
    km2 = coor.cluster_kmeans([data], 15, metric="minRMSD", max_iter=0, fixed_seed=32, init_strategy="kmeans++",
                              n_jobs=1)
    np.testing.assert_array_equal(km.dtrajs[0], km2.dtrajs[0])
    np.testing.assert_array_equal(km.clustercenters, km2.clustercenters)
    assert km.metric == "minRMSD"
    manual = [int(np.argmin([b2k.compute_metric(f, c, "minRMSD") for c in km.clustercenters])) for f in data[:60]]
    np.testing.assert_array_equal(manual, km.dtrajs[0][:60])


def test_kmeans_skip_stride_resume_keep_data():
    X = np.random.RandomState(2).rand(100, 3)
    assert len(coor.cluster_kmeans(X, k=3, skip=42).dtrajs[0]) == 100 - 42   # test_kmeans.py:318-320
    rng = np.random.RandomState(3)
    Y = make_blobs(rng, 6000, [[0, 0], [5, 5], [-5, 5]], 0.5)
    init = np.array([[1, 1], [4, 4], [-4, 4]], np.float32)
    cl = coor.cluster_kmeans(Y, clustercenters=init, k=3, max_iter=1, keep_data=True, tolerance=0)
    assert not cl.converged and cl._dev_frames is not None   # test_kmeans.py:401-409
    d1 = np.abs(cl.clustercenters - [[0, 0], [5, 5], [-5, 5]]).max()
    cl.estimate(Y, clustercenters=cl.clustercenters, max_iter=50, tolerance=1e-7)
    assert cl.converged and cl._dev_frames is None           # freed on convergence (test_kmeans.py:390-399)
    assert np.abs(cl.clustercenters - [[0, 0], [5, 5], [-5, 5]]).max() <= d1   # resume improves (:357-374)
    st = coor.cluster_kmeans(Y, k=3, stride=7, fixed_seed=True)
    assert len(st.dtrajs[0]) == 6000  # dtrajs are assigned at stride 1
    assert coor.cluster_kmeans(Y, k=None, max_iter=1).n_clusters == min(int(np.sqrt(6000)), 5000)


def test_kmeans_rejects_nan():
    X = np.random.RandomState(0).rand(100, 2)
    X[17, 1] = np.nan
    with pytest.raises(Exception, match="invalid"):
        coor.cluster_kmeans(X, k=3)
    # assignment: the staged chunks are checked on the device (no host-side pass over the frames)
    good = np.random.RandomState(1).rand(5000, 2)
    cen = good[:7].copy()
    bad = good.copy()
    bad[4321, 0] = np.inf
    with pytest.raises(Exception, match="invalid"):
        coor.assign_to_centers(bad, cen)
    assert len(coor.assign_to_centers(good, cen)[0]) == 5000


# ---- oracle parity end to end (cfg1 shape at reduced N, golden fixture) ------------------------------------
def test_cfg1_golden_end_to_end(oracle):
    g = np.load(os.path.join(GOLD, "cfg1_small.npz"))
    X = g["X"]
    km = coor.cluster_kmeans(X, k=100, max_iter=10, fixed_seed=42, kmpp_scan="serial")
    c0 = oracle.kmpp_init(X, 100, 42)
    np.testing.assert_array_equal(km.initial_centers_, c0)                 # k-means++ bit-exact (serial scan)
    assert (int(not km.converged), len(km.inertias_)) == (int(g["code"]), int(g["iters"]))  # iteration count
    np.testing.assert_allclose(km.inertias_, g["inertias_f64"], rtol=2e-6)
    assert np.abs(km.clustercenters - g["centers_f64"]).max() <= 1e-5 * np.abs(g["centers_f64"]).max()
    np.testing.assert_allclose(km.inertias_, g["inertias_f32seq"], rtol=1e-4)
    # dtrajs bit-exact given the same centers
    np.testing.assert_array_equal(km.dtrajs[0], oracle.assign(X, km.clustercenters, n_threads=4))
    np.testing.assert_array_equal(coor.assign_to_centers(X, g["centers_f32seq"])[0], g["dtraj"])
    kb = coor.cluster_kmeans(X, k=100, max_iter=1, fixed_seed=42, kmpp_scan="blocked")
    np.testing.assert_array_equal(kb.initial_centers_, X[g["kmpp_blocked_idx"]])


def test_golden_minrmsd_and_cfg2(oracle, b2k):
    g = np.load(os.path.join(GOLD, "minrmsd_small.npz"))
    np.testing.assert_array_equal(coor.assign_to_centers(g["X"], g["C"], metric="minRMSD")[0], g["dtraj"])
    g = np.load(os.path.join(GOLD, "cfg2_small.npz"))
    np.testing.assert_array_equal(coor.assign_to_centers(g["X"], g["C"])[0], g["dtraj"])
    newc, lab = b2k.kmeans_cluster(g["X"], g["C"])
    np.testing.assert_array_equal(lab, g["dtraj"])
    assert np.abs(newc - g["newC_f32seq"]).max() <= 1e-5 * np.abs(g["newC_f32seq"]).max()
    np.testing.assert_allclose(newc, g["newC_f64"], rtol=3e-7, atol=1e-7)


# ---- test_assign.py -----------------------------------------------------------------------------------
def test_assign_to_centers_reference_cases():
    rng = np.random.RandomState(0)
    centers = np.array([[0, 0, 0], [10, 0, 0], [0, 10, 0], [0, 0, 10], [10, 10, 10]], np.float32)
    X = np.concatenate([c + 0.1 * rng.randn(1000, 3) for c in centers])
    ass = coor.assign_to_centers(X, centers, return_dtrajs=False)
    assert len(ass.dtrajs) == 1 and ass.dtrajs[0].dtype == ass.output_type()
    assert (ass.dtrajs[0] == np.arange(5000) // 1000).all()                # test_assign.py:102-109
    np.testing.assert_array_equal(ass.transform(X), ass.get_output()[0])    # :137-143
    dtr = coor.assign_to_centers(data=X, centers=centers)
    np.testing.assert_array_equal(dtr[0], ass.dtrajs[0])
    with pytest.raises(ValueError):                                          # :183-197
        coor.assign_to_centers(X, centers[:, :2])
    with pytest.raises(ValueError):
        ass.assign(X, stride=2)
    # chunked / list input / stride
    two = coor.assign_to_centers([X[:1234], X[1234:]], centers, chunksize=100)
    np.testing.assert_array_equal(np.concatenate(two), ass.dtrajs[0])
    np.testing.assert_array_equal(ass.assign(stride=3)[0], ass.dtrajs[0][::3] if False else (np.arange(5000) // 1000)[::3])


def test_assign_from_file(tmp_path):
    c = np.random.RandomState(1).randn(6, 4)
    X = np.random.RandomState(2).randn(300, 4)
    np.save(tmp_path / "c.npy", c)
    np.savetxt(tmp_path / "c.dat", c)
    a = coor.assign_to_centers(X, str(tmp_path / "c.npy"))[0]
    b = coor.assign_to_centers(X, str(tmp_path / "c.dat"))[0]
    np.testing.assert_array_equal(a, coor.assign_to_centers(X, c)[0])
    assert (a == b).mean() > 0.99


# ---- test_regspace.py / test_cluster_samples.py --------------------------------------------------------------
def test_regspace_reference_cases(oracle):
    trajs = [[0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 1, 2], [3, 4, 5], [6, 7, 8]]
    cl = coor.cluster_regspace(data=trajs, dmin=.5)           # test_cluster_samples.py:41-60
    np.testing.assert_array_equal(cl.clustercenters.ravel(), np.arange(9))
    assert cl.n_clusters == 9
    ref = [[[0, 0], [3, 0]], [[0, 1], [3, 1]], [[0, 2], [3, 2]], [[1, 0], [4, 0]], [[1, 1], [4, 1]],
           [[1, 2], [4, 2]], [[2, 0], [5, 0]], [[2, 1], [5, 1]], [[2, 2], [5, 2]]]
    for cc in range(cl.n_clusters):
        np.testing.assert_array_equal(cl.index_clusters[cc], ref[cc])
    for ii, s in enumerate(cl.sample_indexes_by_cluster(np.arange(cl.n_clusters), 10)):
        assert all(cl.dtrajs[a][b] == ii for a, b in s)

    rng = np.random.RandomState(0)
    src = DataInMemory([rng.rand(1000, 3), rng.rand(700, 3)], chunksize=250)
    rs = coor.RegularSpaceClustering(dmin=0.3)
    rs.estimate(src)
    X = np.concatenate([t for t in src.data]).astype(np.float32)
    ref_c, _, _ = oracle.regspace(X, 0.3, 1000)
    np.testing.assert_array_equal(rs.clustercenters, ref_c)
    assert len(np.unique(np.concatenate(rs.dtrajs))) == len(rs.clustercenters)   # test_regspace.py:77-89
    assert coor.cluster_regspace(rng.rand(500), dmin=0.1).clustercenters.shape[1] == 1   # 1-D input (:96-98)
    for metric in ("euclidean", "minRMSD"):                      # :107-115, :137-141
        a = coor.cluster_regspace(src, dmin=0.3, metric=metric, n_jobs=1)
        b = coor.cluster_regspace(src, dmin=0.3, metric=metric, n_jobs=2)
        np.testing.assert_equal(a.clustercenters, b.clustercenters)
        np.testing.assert_array_equal(a.clustercenters, oracle.regspace(X, 0.3, 1000, metric)[0])


def test_regspace_max_centers_warns_once():
    # test_regspace.py:117-135
    rng = np.random.RandomState(1)
    src = DataInMemory([rng.rand(1000, 3)], chunksize=300)
    rs = coor.RegularSpaceClustering(dmin=1e-8, max_centers=50)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        rs.estimate(src)
        assert len(w) == 1
    assert len(rs.clustercenters) == 50 and not rs.converged
    out = rs.get_output()
    assert len(out) == rs.number_of_trajectories() and len(out[0]) == rs.trajectory_lengths()[0]


# ---- test_cluster.py: generic estimator contract ---------------------------------------------------------------
@pytest.mark.parametrize("make", [lambda X: coor.cluster_kmeans(X, k=100, max_iter=3),
                                  lambda X: coor.cluster_regspace(X, dmin=0.5)])
def test_generic_contract(make):
    rng = np.random.RandomState(5)
    X = [rng.randn(1500, 3), rng.randn(800, 3)]
    cl = make(X)
    assert isinstance(cl.chunksize, int)
    assert cl.clustercenters.shape[1] == 3 and cl.clustercenters.shape[0] == cl.n_clusters
    assert cl.dimension() == 1
    assert [len(d) for d in cl.dtrajs] == [1500, 800] and all(d.dtype == np.int32 for d in cl.dtrajs)
    out = cl.get_output()
    assert all(o.shape == (n, 1) and o.dtype == np.int32 for o, n in zip(out, (1500, 800)))
    for itraj, chunk in cl.iterator(chunk=400):
        assert chunk.shape[1] == 1 and chunk.dtype == np.int32 and len(chunk) <= 400
    np.testing.assert_array_equal(cl.transform(X[0]), out[0])
    assert isinstance(cl.describe(), str)


def test_save_dtrajs(tmp_path):
    X = np.random.RandomState(0).randn(200, 2)
    cl = coor.cluster_kmeans([X, X[:50]], k=5, max_iter=2)
    cl.save_dtrajs(prefix="pre", output_dir=str(tmp_path))
    np.testing.assert_array_equal(np.loadtxt(tmp_path / "pre_0.dtraj", dtype=int), cl.dtrajs[0])
    cl.save_dtrajs(prefix="pre", output_dir=str(tmp_path), output_format="npy", extension=".npy")
    np.testing.assert_array_equal(np.load(tmp_path / "pre_1.npy"), cl.dtrajs[1])
    with pytest.raises(EnvironmentError):
        cl.save_dtrajs(prefix="pre", output_dir=str(tmp_path))


def test_mini_batch_kmeans_matches_oracle_emulation(oracle):
    """MiniBatchKmeansClustering (kmeans.py:341-447): with the numpy RNG seeded the batches are reproducible, so the
    whole pass loop -- one Lloyd step per batch, cost with the batch re-assigned to the new centers, relative-change
    stop -- is emulated with the oracle on the same batches."""
    import pyemma_b200 as coor
    rng = np.random.RandomState(9)
    cen = rng.uniform(-4, 4, size=(12, 3))
    trajs = [(cen[rng.randint(0, 12, L)] + 0.3 * rng.randn(L, 3)).astype(np.float32) for L in (9000, 4000, 500)]
    C0 = np.concatenate(trajs)[rng.choice(13500, 40, replace=False)].copy()
    np.random.seed(123)
    mb = coor.cluster_mini_batch_kmeans(trajs, k=40, max_iter=5, batch_size=0.3, clustercenters=C0)
    # emulation
    emu = coor.MiniBatchKmeansClustering(40, max_iter=5, batch_size=0.3)
    src = DataInMemory(trajs)
    emu.skip = 0
    emu._init_batches(src)
    np.random.seed(123)
    emu._draw_mini_batch_sample()
    c, prev, inert, conv = C0, 0.0, [], False
    for _ in range(5):
        batch = np.ascontiguousarray(src.ra_gather(emu._draw_mini_batch_sample()), dtype=np.float32)
        c, _lab = oracle.kmeans_cluster(batch, c, n_threads=4, acc="f64")
        lab2 = oracle.assign(batch, c, n_threads=4)
        cost = float(oracle.cost(batch, c, lab2, acc="f64"))
        inert.append(cost)
        rel = abs(cost - prev) / cost if cost != 0 else 0.0
        prev = cost
        if rel <= 1e-5:
            conv = True
            break
    assert mb.converged == conv and len(mb.inertias_) == len(inert)
    np.testing.assert_allclose(mb.inertias_, inert, rtol=5e-6)
    assert np.abs(mb.clustercenters - c).max() <= 1e-5 * np.abs(c).max()
    assert len(mb.dtrajs) == 3 and [len(x) for x in mb.dtrajs] == [9000, 4000, 500]
    np.testing.assert_array_equal(np.concatenate(mb.dtrajs), oracle.assign(np.concatenate(trajs), mb.clustercenters))
    # without given centers: k-means++ on the first batch, then the same loop; result is a usable clustering
    np.random.seed(5)
    mb2 = coor.cluster_mini_batch_kmeans(trajs, k=12, max_iter=8, batch_size=0.5)
    assert mb2.clustercenters.shape == (12, 3) and mb2.initial_centers_.shape == (12, 3)
    assert len(np.unique(np.concatenate(mb2.dtrajs))) == 12


def test_minrmsd_clustering_is_rotation_translation_invariant(b2k):
    """reference tests/test_kmeans.py:266-316: noisy, randomly rotated and translated copies of 5 template structures
    -- k-means++ (metric minRMSD) picks one center per template and every copy lands in its template's cluster."""
    rng = np.random.RandomState(0)
    n_atoms, n_templates, copies = 30, 5, 60
    T = rng.uniform(-3, 3, size=(n_templates, n_atoms, 3))
    frames, truth = [], []
    for t in range(n_templates):
        for _ in range(copies):
            q = rng.randn(4)
            q /= np.linalg.norm(q)
            a, b, c, d = q
            R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                          [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                          [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])
            frames.append(((T[t] + 0.01 * rng.randn(n_atoms, 3)) @ R.T + rng.uniform(-10, 10, 3)).reshape(-1))
            truth.append(t)
    X = np.array(frames, np.float32)
    truth = np.array(truth)
    centers = b2k.kmeans_init_centers_kmpp(X, n_templates, 3, metric="minRMSD", scan="blocked")
    dt = coor.assign_to_centers(X, centers, metric="minRMSD")[0]
    assert len(np.unique(dt)) == n_templates
    for t in range(n_templates):
        assert len(np.unique(dt[truth == t])) == 1
    km = coor.cluster_kmeans(X, k=n_templates, max_iter=3, metric="minRMSD", fixed_seed=3)
    for t in range(n_templates):
        assert len(np.unique(km.dtrajs[0][truth == t])) == 1


def test_uniform_time_clustering(oracle):
    """uniform_time.py:33-106: centers are frames picked uniformly in time over the concatenated trajectories"""
    rng = np.random.RandomState(2)
    trajs = [rng.randn(L, 3).astype(np.float32) for L in (1000, 300, 57)]
    ut = coor.cluster_uniform_time(trajs, k=20)
    T, k = 1357, 20
    next_t = (T // k) // 2
    idx = np.arange(next_t, T - next_t + 1, (T - 2 * next_t + 1) // k)[:k]
    allf = np.concatenate(trajs)
    np.testing.assert_array_equal(ut.clustercenters, allf[idx])
    assert ut.clustercenters.shape == (20, 3) and ut.n_clusters == 20
    np.testing.assert_array_equal(np.concatenate(ut.dtrajs), oracle.assign(allf, ut.clustercenters))
    big = coor.cluster_uniform_time(trajs[2], k=500)            # more clusters than frames -> clipped
    assert big.n_clusters == 57 and len(big.clustercenters) == 57
    auto = coor.cluster_uniform_time(trajs, k=None)
    assert auto.n_clusters == int(np.sqrt(1357))


# ---- the tier below HBM (reference: host memmap, kmeans.py:181-200) ---------------------------------------------------
@pytest.mark.parametrize("metric,d", [("euclidean", 10), ("euclidean", 40), ("minRMSD", 30)])
def test_out_of_core_kmeans_is_bit_identical(monkeypatch, oracle, metric, d):
    """with a tiny artificial HBM budget the frames stay in pinned host memory and pass through the device once per
    iteration (b2k_stage_lloyd_pass); centers, inertias, iteration count and dtrajs equal the resident fit bit for bit"""
    from pyemma_b200 import _lib
    rng = np.random.RandomState(31)
    n, k = 60_000, 150
    cen = rng.uniform(-4, 4, size=(9, d))
    X = (cen[rng.randint(0, 9, n)] + 0.6 * rng.randn(n, d)).astype(np.float32)
    trajs = [X[:25_000], X[25_000:25_100], X[25_100:]]
    C0 = X[rng.choice(n, k, replace=False)].copy()
    resident = coor.cluster_kmeans(trajs, k=k, max_iter=7, clustercenters=C0.copy(), metric=metric, tolerance=1e-7)
    ctx = _lib.context()
    ctx.set_option("stage_bytes", 1 << 19)      # a few thousand frames per chunk: many chunks per pass
    monkeypatch.setenv("B2K_HBM_BUDGET_BYTES", str(n * d * 4 // 3))
    try:
        ooc = coor.cluster_kmeans(trajs, k=k, max_iter=7, clustercenters=C0.copy(), metric=metric, tolerance=1e-7)
    finally:
        ctx.set_option("stage_bytes", 64 << 20)
    np.testing.assert_array_equal(ooc.clustercenters, resident.clustercenters)
    np.testing.assert_array_equal(ooc.inertias_, resident.inertias_)
    assert ooc.converged == resident.converged and len(ooc.inertias_) == len(resident.inertias_)
    for a, b in zip(ooc.dtrajs, resident.dtrajs):
        np.testing.assert_array_equal(a, b)
    ref = oracle.assign(X, ooc.clustercenters, metric, n_threads=8)
    np.testing.assert_array_equal(np.concatenate(ooc.dtrajs), ref)
    # k-means++ seeding out of core runs on a strided subset that fits the budget: a valid fit, centers are frames
    km = coor.cluster_kmeans(trajs, k=20, max_iter=3, fixed_seed=7, metric=metric)
    assert km.clustercenters.shape == (20, d) and np.isfinite(km.clustercenters).all()
    assert all((km.initial_centers_[i] == X).all(axis=1).any() for i in range(20))
    monkeypatch.setenv("B2K_HBM_BUDGET_BYTES", "1000")
    with pytest.raises(MemoryError):
        from pyemma_b200.clustering import KmeansClustering
        KmeansClustering(20, max_iter=2, clustercenters=C0[:20].copy(), oom_strategy="raise").estimate(trajs)
