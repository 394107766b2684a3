"""GPU: the linear projection fused into the chunk hand-off against the reference formula
`((X - mean) @ eigenvectors[:, :dim]).astype(float32)` (pyemma/coordinates/transform/_tica_base.py:130-133) in numpy
fp64, and against a plain torch fp32 matmul (tolerances stated below)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,din,dim", [(5000, 64, 10), (70001, 30, 3), (1234, 7, 7), (300, 513, 2), (2, 4, 1)])
def test_projection_matches_numpy_fp64(n, din, dim):
    import torch
    from pyemma_b200.transform import LinearProjection
    rng = np.random.RandomState(din)
    X = (rng.randn(n, din) * 3 + 10).astype(np.float32)
    mean = X.astype(np.float64).mean(0)
    W = np.linalg.qr(rng.randn(din, din))[0][:, :max(dim, min(din, dim + 2))]
    lp = LinearProjection(X, mean, W, dim=dim)
    got = lp.transform(X)
    ref = ((X.astype(np.float64) - mean) @ W[:, :dim]).astype(np.float32)
    assert got.dtype == np.float32 and got.shape == (n, dim)
    # one fp64 FMA chain per output, rounded once: within 1 ulp(fp32) of the fp64 reference (plus its own 1e-13 noise)
    np.testing.assert_allclose(got, ref, rtol=2.5e-7, atol=1e-6 * np.abs(ref).max())
    # torch fp32 reference of the same op: fp32 accumulation error only
    t = ((torch.from_numpy(X) - torch.from_numpy(mean.astype(np.float32))) @ torch.from_numpy(W[:, :dim].astype(np.float32))).numpy()
    np.testing.assert_allclose(got, t, rtol=0, atol=3e-5 * np.abs(X - mean).max() * np.sqrt(din))
    assert lp.dimension() == dim and [len(y) for y in lp.get_output()] == [n]


def test_kmeans_on_projection_source_is_fused_and_identical(b2k):
    """cluster_kmeans(LinearProjection(raw, ...)) == cluster_kmeans(projected array): the fused path stages raw chunks,
    projects them on the device and never materialises the projected array on the host"""
    import pyemma_b200 as coor
    from pyemma_b200.transform import LinearProjection
    rng = np.random.RandomState(1)
    raw = [(rng.randn(L, 40) + rng.randint(0, 5, (L, 1))).astype(np.float32) for L in (30000, 12000)]
    mean = np.concatenate(raw).astype(np.float64).mean(0)
    W = np.linalg.qr(rng.randn(40, 40))[0]
    lp = LinearProjection(raw, mean, W, dim=6)
    Y = lp.get_output()
    C0 = np.concatenate(Y)[rng.choice(42000, 50, replace=False)].copy()
    a = coor.cluster_kmeans(lp, k=50, max_iter=4, clustercenters=C0, chunksize=7000)
    b = coor.cluster_kmeans(Y, k=50, max_iter=4, clustercenters=C0)
    np.testing.assert_array_equal(a.clustercenters, b.clustercenters)
    np.testing.assert_array_equal(a.inertias_, b.inertias_)
    for da, db in zip(a.dtrajs, b.dtrajs):
        np.testing.assert_array_equal(da, db)
    with pytest.raises(ValueError):
        LinearProjection(raw, mean, W[:10], dim=2)
