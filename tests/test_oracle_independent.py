"""CPU: the oracle against THIRD-PARTY implementations of the same algorithms (scikit-learn, scipy).

deeptime / mdtraj -- the packages whose arithmetic the oracle restates -- are absent offline, so the bits of the
reference's fp32 summation order cannot be pinned here ("parity unpinned upstream", oracle/oracle.cpp header).  What
CAN be pinned independently of anything written for this repo is the semantics: Lloyd trajectories (labels, centers,
iteration where the labels stop changing, inertia), the k-means++ D^2-sampling statistics and the minimal RMSD.
  * sklearn.cluster.KMeans(algorithm="lloyd", init=C0, n_init=1)  -- exact Lloyd iterations in fp64
  * sklearn.cluster.kmeans_plusplus                                -- greedy k-means++ with 2+log(k) local trials,
                                                                      the same variant deeptime implements
  * scipy.spatial.transform.Rotation.align_vectors                 -- Kabsch / Horn optimal rotation, rssd
Tolerances are stated at each check.
"""
import warnings

import numpy as np
import pytest

sklearn_cluster = pytest.importorskip("sklearn.cluster")
Rotation = pytest.importorskip("scipy.spatial.transform").Rotation


def _blobs(rng, n, d, nb, spread=4.0, sigma=0.5):
    cen = rng.uniform(-spread, spread, size=(nb, d))
    return (cen[rng.randint(0, nb, n)] + sigma * rng.randn(n, d)).astype(np.float32)


def _sk_lloyd(X, C0, iters):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # ConvergenceWarning when max_iter is hit
        km = sklearn_cluster.KMeans(n_clusters=len(C0), init=C0.astype(np.float64), n_init=1, max_iter=iters, tol=0.0,
                                    algorithm="lloyd").fit(X.astype(np.float64))
    return km


@pytest.mark.parametrize("n,d,k", [(20000, 2, 30), (30000, 10, 100), (8000, 64, 40)])
def test_lloyd_trajectory_matches_sklearn(oracle, n, d, k):
    rng = np.random.RandomState(d)
    X = _blobs(rng, n, d, max(3, k // 5))
    C0 = X[rng.choice(n, k, replace=False)].copy()
    for iters in (1, 3, 8):
        cen, code, it, inert, hist, lab = oracle.cluster_loop(X, C0, iters, 0.0, acc="f64", n_threads=4, history=True)
        km = _sk_lloyd(X, C0, iters)
        if km.n_iter_ < iters:
            continue  # sklearn stopped early because its labels stopped changing; compared below at convergence
        # centers after exactly `iters` M-steps: fp32 distances + exact sums vs fp64 everything; a handful of frames
        # that are equidistant to ~1e-7 may be labelled differently, each moves a center by <= |x|/count
        np.testing.assert_allclose(cen, km.cluster_centers_, rtol=0, atol=2e-5 * np.abs(X).max())
    # run both to their fixed point: same labels, same inertia, and the oracle's converged flag agrees
    cen, code, it, inert, hist, lab = oracle.cluster_loop(X, C0, 300, 0.0, acc="f64", n_threads=4, history=True)
    km = _sk_lloyd(X, C0, 300)
    assert code == 0 and km.n_iter_ < 300
    np.testing.assert_allclose(cen, km.cluster_centers_, rtol=0, atol=2e-5 * np.abs(X).max())
    final = oracle.assign(X, cen, n_threads=4)
    assert (final != km.labels_).mean() <= 1e-4   # fp32-vs-fp64 near-ties only
    np.testing.assert_allclose(float(inert[-1]), km.inertia_, rtol=2e-5)
    # deeptime's cost is sum of (fp32 distance)^2 -- independent of sklearn's GEMM-based distances
    d2 = ((X.astype(np.float64) - cen.astype(np.float64)[final]) ** 2).sum()
    np.testing.assert_allclose(float(oracle.cost(X, cen, final, acc="f64")), d2, rtol=1e-6)


def test_lloyd_step_labels_match_sklearn_estep(oracle):
    # one E-step: every label equals the fp64 argmin unless the two best distances agree to 1e-6 relative
    rng = np.random.RandomState(7)
    X = _blobs(rng, 50000, 10, 20)
    C = X[rng.choice(len(X), 500, replace=False)].copy()
    lab = oracle.assign(X, C, n_threads=4)
    D = ((X.astype(np.float64)[:, None, :] - C.astype(np.float64)[None, :, :]) ** 2).sum(-1) if len(X) <= 2000 else None
    km = sklearn_cluster.KMeans(n_clusters=len(C), init=C.astype(np.float64), n_init=1, max_iter=1, tol=0.0)
    km.cluster_centers_ = C.astype(np.float64)
    km._n_threads = 1
    sk = km.predict(X.astype(np.float64))
    bad = np.nonzero(sk != lab)[0]
    assert len(bad) <= 5
    for i in bad:  # genuine near-ties only
        di = np.sqrt(((X[i].astype(np.float64) - C.astype(np.float64)[[lab[i], sk[i]]]) ** 2).sum(-1))
        assert abs(di[0] - di[1]) <= 1e-6 * di.max()


def test_kmpp_potential_statistics_match_sklearn(oracle):
    """Same algorithm (first center uniform, then 2+floor(ln k) D^2-weighted trials, keep the trial with the lowest
    potential) but different RNG streams: the distribution of the final potential must agree.  40 seeds each;
    the means are compared at 4 standard errors."""
    rng = np.random.RandomState(11)
    X = _blobs(rng, 4000, 5, 12, spread=6.0, sigma=0.7)
    k = 24

    def potential(C):
        D = ((X[:, None, :].astype(np.float64) - C[None].astype(np.float64)) ** 2).sum(-1)
        return D.min(1).sum()

    ours = np.array([potential(oracle.kmpp_init(X, k, seed, scan="serial")) for seed in range(40)])
    blocked = np.array([potential(oracle.kmpp_init(X, k, seed, scan="blocked")) for seed in range(40)])
    sk = np.array([potential(sklearn_cluster.kmeans_plusplus(X.astype(np.float64), k, random_state=seed)[0])
                   for seed in range(40)])
    se = np.sqrt(ours.var() / 40 + sk.var() / 40)
    assert abs(ours.mean() - sk.mean()) <= 4 * se, (ours.mean(), sk.mean(), se)
    se_b = np.sqrt(blocked.var() / 40 + sk.var() / 40)
    assert abs(blocked.mean() - sk.mean()) <= 4 * se_b
    # greedy k-means++ beats plain uniform seeding by a wide margin on clustered data: a sanity anchor for the scale
    uni = np.array([potential(X[np.random.RandomState(s).choice(len(X), k, replace=False)]) for s in range(40)])
    assert ours.mean() < 0.8 * uni.mean()
    # every pick is a data row and no row is picked twice
    C, idx = oracle.kmpp_init(X, k, 3, return_indices=True)
    assert len(set(idx.tolist())) == k and np.array_equal(C, X[idx])


def test_minrmsd_matches_scipy_align_vectors(oracle):
    """minRMSD = sqrt(min_R sum |a_i - R b_i|^2 / n_atoms) after centring both structures: scipy's align_vectors returns
    that root-sum-squared distance.  fp32 QCP (Newton on the quartic) vs fp64 SVD; tolerance stated at the assert."""
    rng = np.random.RandomState(0)
    for trial in range(200):
        na = int(rng.randint(3, 80))
        a = rng.uniform(-5, 5, (na, 3))
        if trial % 3 == 0:   # a noisy rotated copy: small RMSD, the regime regspace/minRMSD clustering lives in
            R = Rotation.random(random_state=trial).as_matrix()
            b = a @ R.T + 0.05 * rng.randn(na, 3) + rng.uniform(-3, 3, 3)
        else:
            b = rng.uniform(-5, 5, (na, 3))
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        ac = a32.astype(np.float64) - a32.astype(np.float64).mean(0)
        bc = b32.astype(np.float64) - b32.astype(np.float64).mean(0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _rot, rssd = Rotation.align_vectors(ac, bc)
        ref = rssd / np.sqrt(na)
        got = float(oracle.compute_metric(a32.ravel(), b32.ravel(), "minRMSD"))
        # the QCP msd is (G_a + G_b - 2 lambda_max) / n evaluated in fp32: its ABSOLUTE error scales with the mean
        # square radius of the (uncentred, clustering_module.cpp:28) structures, so small RMSDs carry a larger
        # relative error -- 2e-6 of that scale on the msd, i.e. ~30 ulp of the cancelling terms
        scale = float((a32.astype(np.float64) ** 2).sum() + (b32.astype(np.float64) ** 2).sum()) / na
        assert abs(got * got - ref * ref) <= 2e-6 * scale, (trial, got, ref, scale)


def test_minrmsd_assign_matches_scipy_argmin(oracle):
    rng = np.random.RandomState(4)
    na = 20
    T = rng.uniform(-3, 3, (6, na, 3))
    X = np.stack([T[rng.randint(6)] @ Rotation.random(random_state=i).as_matrix().T + 0.05 * rng.randn(na, 3)
                  for i in range(150)]).astype(np.float32)
    Cn = T.astype(np.float32)
    lab = oracle.assign(X.reshape(len(X), -1), Cn.reshape(6, -1), "minRMSD")
    for i in range(len(X)):
        a = X[i].astype(np.float64)
        a -= a.mean(0)
        r = []
        for c in Cn.astype(np.float64):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r.append(Rotation.align_vectors(a, c - c.mean(0))[1])
        assert int(np.argmin(r)) == lab[i]


def test_regspace_matches_bruteforce_python(oracle):
    """regular-space clustering is fully specified by its definition (a frame becomes a center iff it is >= dmin from
    every earlier center): a 15-line numpy loop in fp64 is an independent implementation."""
    rng = np.random.RandomState(9)
    X = rng.uniform(-2, 2, (3000, 3)).astype(np.float32)
    for dmin in (0.25, 0.5, 1.0):
        cen, idx, full = oracle.regspace(X, dmin, 5000)
        mine = [0]
        for i in range(1, len(X)):
            dd = np.sqrt(((X[mine].astype(np.float64) - X[i].astype(np.float64)) ** 2).sum(1))
            if dd.min() >= dmin:
                mine.append(i)
        assert not full
        # fp32 vs fp64 can only differ for a frame whose nearest center is within 1e-6 of dmin
        assert len(set(idx.tolist()) ^ set(mine)) <= 1
