"""GPU: randomized shape sweep -- labels of assign / one Lloyd step must equal the oracle's bit for bit for arbitrary
(n, d, k), both assignment engines, including degenerate sizes (n=1, k=1, d=1, k>n duplicates, ragged tails)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.RandomState(1000 + seed)
    n = int(rng.choice([1, 2, 31, 127, 128, 129, 1000, 4097, 6000, rng.randint(1, 9000)]))
    d = int(rng.choice([1, 2, 3, 4, 5, 9, 10, 15, 16, 17, 31, 32, 33, 64, 100, 130, rng.randint(1, 200)]))
    k = int(rng.choice([1, 2, 7, 8, 9, 31, 32, 33, 127, 128, 129, 255, 256, 257, 700, rng.randint(1, 900)]))
    nb = int(rng.randint(1, 9))
    cen = rng.uniform(-6, 6, size=(nb, d))
    X = (cen[rng.randint(0, nb, n)] + rng.randn(n, d) * rng.choice([0.01, 0.5, 3.0])).astype(np.float32)
    pick = rng.randint(0, n, k)
    C = X[pick].copy()
    jitter = rng.rand(k) < 0.5
    C[jitter] += (rng.randn(int(jitter.sum()), d) * 0.02).astype(np.float32)
    if rng.rand() < 0.3:                                    # large common offset: centring must cope
        off = np.float32(rng.choice([100.0, -3000.0]))
        X, C = X + off, C + off
    return X, C


@pytest.mark.parametrize("seed", range(40))
def test_random_assign_and_lloyd_step(b2k, oracle, seed):
    X, C = _case(seed)
    ref = oracle.assign(X, C, n_threads=4)
    ctx = b2k.context()
    try:
        for eng in (b2k.ENGINE_AUTO, b2k.ENGINE_SCREEN, b2k.ENGINE_DIRECT):
            ctx.set_option("assign_engine", eng)
            np.testing.assert_array_equal(b2k.assign(X, C), ref, err_msg="engine=%d shape=%s k=%d" % (eng, X.shape, len(C)))
        ctx.set_option("assign_engine", b2k.ENGINE_AUTO)
        newc, lab = b2k.kmeans_cluster(X, C)
        rnew, rlab = oracle.kmeans_cluster(X, C, n_threads=4, acc="f64")
        np.testing.assert_array_equal(lab, rlab)
        scale = max(np.abs(X).max(), 1e-30)
        assert np.abs(newc - rnew).max() <= 2e-7 * scale + 1e-7 * np.abs(rnew).max()
    finally:
        ctx.set_option("assign_engine", b2k.ENGINE_AUTO)
