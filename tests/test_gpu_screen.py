"""GPU: the tcgen05 screen + exact verify path must give the oracle's labels bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def blobs(rng, n, d, nb, spread=5.0, sigma=1.0):
    cen = rng.uniform(-spread, spread, size=(nb, d))
    lab = rng.randint(0, nb, size=n)
    return (cen[lab] + sigma * rng.randn(n, d)).astype(np.float32)


@pytest.fixture()
def screen_ctx(b2k):
    ctx = b2k.context()
    ctx.set_option("assign_engine", b2k.ENGINE_SCREEN)
    yield ctx
    ctx.set_option("assign_engine", b2k.ENGINE_AUTO)
    ctx.set_option("screen_terms", 0)
    ctx.set_option("screen_group", 0)
    ctx.set_option("verify_mode", 0)


SHAPES = [(5000, 2, 100, 0), (20000, 10, 1000, 1), (20000, 10, 1000, 3), (4096, 3, 257, 0), (3000, 16, 300, 0),
          (6000, 17, 513, 0), (5000, 64, 2000, 3), (5000, 64, 2000, 1), (5000, 64, 2000, 2), (2500, 256, 1000, 3),
          (2500, 256, 1000, 2), (2500, 256, 1000, 1), (20000, 10, 1000, 2), (1000, 300, 64, 0), (129, 5, 40, 0)]


@pytest.mark.parametrize("n,d,k,terms", SHAPES)
def test_screen_assign_bit_exact(b2k, oracle, screen_ctx, n, d, k, terms):
    rng = np.random.RandomState(n + 7 * d + k)
    X = blobs(rng, n, d, 12)
    C = X[rng.choice(n, k, replace=k > n)].copy()
    C[: k // 2] += (0.05 * rng.randn(k // 2, d)).astype(np.float32)  # near-duplicates: small gaps
    screen_ctx.set_option("screen_terms", terms)
    ref = oracle.assign(X, C, n_threads=8)
    for group in (0, 8, 4, 2):  # centers per candidate group handed to the exact verify (0: automatic)
        for vmode in ((0, 1) if d > 16 else (0,)):  # wide rows: direct / shared-memory staged verify kernels
            screen_ctx.set_option("screen_group", group)
            screen_ctx.set_option("verify_mode", vmode)
            got = b2k.assign(X, C)
            np.testing.assert_array_equal(got, ref, err_msg="screen_group=%d verify_mode=%d" % (group, vmode))


def test_screen_measured_term_count(b2k, oracle, screen_ctx):
    """option screen_terms=0: a Lloyd session on wide rows measures the candidate counts of 1 and 2 operand terms on a
    sample and picks one; whatever it picks, the labels -- and therefore the exact member sums -- do not change"""
    rng = np.random.RandomState(11)
    X = blobs(rng, 100_000, 256, 40, spread=3.0, sigma=1.0)
    C0 = X[rng.choice(len(X), 5000, replace=False)].copy()
    res = {}
    screen_ctx.set_option("probe_min_gflop", 100)   # the probe is meant for jobs of >= 1 TFLOP per pass: lower the bar
    try:
        for terms in (0, 3):
            screen_ctx.set_option("screen_terms", terms)
            res[terms] = b2k.kmeans_cluster_loop(X, C0, 2, 0.0)
            used = int(screen_ctx.get_stat("screen_terms_used"))
            assert used == 3 if terms == 3 else used in (1, 2, 3)
    finally:
        screen_ctx.set_option("probe_min_gflop", 1000)
    assert screen_ctx.get_stat("probe_centers_1") > 0          # the probe ran
    np.testing.assert_array_equal(res[0][0], res[3][0])        # centers bit-identical
    np.testing.assert_array_equal(res[0][3], res[3][3])        # inertias too
    rc, rcode, rit, rin = oracle.cluster_loop(X, C0, 2, 0.0, n_threads=16, acc="f64")
    assert np.abs(res[0][0] - rc).max() <= 1e-5 * np.abs(rc).max()


def test_screen_offset_data_and_ties(b2k, oracle, screen_ctx):
    # large common offset (centering must cope), exact duplicates of centers (lowest index wins)
    rng = np.random.RandomState(3)
    X = (blobs(rng, 8000, 12, 6, spread=2.0, sigma=0.3) + 1000.0).astype(np.float32)
    base = X[rng.choice(8000, 150, replace=False)]
    C = np.concatenate([base, base, base[:50]]).astype(np.float32)
    ref = oracle.assign(X, C, n_threads=8)
    got = b2k.assign(X, C)
    np.testing.assert_array_equal(got, ref)
    assert got.max() < 150


def test_screen_outliers_and_constant_data(b2k, oracle, screen_ctx):
    rng = np.random.RandomState(4)
    X = blobs(rng, 6000, 8, 5)
    X[100] *= 1e4  # far outlier changes the scale
    X[200] = 0
    C = X[rng.choice(6000, 300, replace=False)].copy()
    np.testing.assert_array_equal(b2k.assign(X, C), oracle.assign(X, C, n_threads=8))
    Z = np.full((5000, 4), 3.25, np.float32)
    Cz = np.full((130, 4), 3.25, np.float32)
    np.testing.assert_array_equal(b2k.assign(Z, Cz), oracle.assign(Z, Cz))


def test_screen_lloyd_matches_direct(b2k, oracle, screen_ctx):
    rng = np.random.RandomState(5)
    X = blobs(rng, 30000, 10, 20, spread=3.0, sigma=0.8)
    C0 = X[rng.choice(30000, 500, replace=False)].copy()
    cen_s, code_s, it_s, in_s = b2k.kmeans_cluster_loop(X, C0, 6, 0.0)
    screen_ctx.set_option("assign_engine", b2k.ENGINE_DIRECT)
    cen_d, code_d, it_d, in_d = b2k.kmeans_cluster_loop(X, C0, 6, 0.0)
    np.testing.assert_array_equal(cen_s, cen_d)      # same labels -> identical exact sums
    np.testing.assert_array_equal(in_s, in_d)
    rc, rcode, rit, rin = oracle.cluster_loop(X, C0, 6, 0.0, n_threads=8, acc="f64")
    assert np.abs(cen_s - rc).max() <= 1e-5 * np.abs(rc).max()


def test_screen_candidate_statistics(b2k, screen_ctx):
    import ctypes as C
    import torch
    rng = np.random.RandomState(6)
    X = blobs(rng, 100000, 10, 20, spread=3.0, sigma=0.8)
    Cn = X[rng.choice(100000, 1000, replace=False)].copy()
    dev = torch.device("cuda", screen_ctx.device)
    dX, dC = torch.from_numpy(X).to(dev), torch.from_numpy(Cn).to(dev)
    lab = torch.empty(len(X), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    b2k.check(screen_ctx.lib.b2k_dev_assign(screen_ctx.handle, C.c_void_p(dX.data_ptr()), len(X), 10,
                                            C.c_void_p(dC.data_ptr()), 1000, 0, C.c_void_p(lab.data_ptr()), None))
    frames = screen_ctx.get_stat("screen_frames")
    assert frames == len(X)
    chunks = screen_ctx.get_stat("screen_cand_chunks") / frames
    fb = screen_ctx.get_stat("screen_fallback_frames") / frames
    print("candidate chunks per frame %.3f, fallback fraction %.5f" % (chunks, fb))
    assert chunks < 3.0 and fb < 0.01


@pytest.mark.parametrize("n,d,k", [(5000, 64, 2000), (3000, 100, 900), (4000, 33, 3000)])
def test_screen_resident_frame_tile_mode(b2k, oracle, screen_ctx, n, d, k):
    """option screen_resident_a: the frame tile stays in shared memory while its center tiles stream (two polled TMA
    streams, per-k-block barriers) -- same labels as the streaming mode and the oracle"""
    rng = np.random.RandomState(d + k)
    X = blobs(rng, n, d, 9)
    Cn = X[rng.choice(n, k, replace=False)].copy()
    ref = oracle.assign(X, Cn, n_threads=8)
    for mode in (1, 0):
        screen_ctx.set_option("screen_resident_a", mode)
        try:
            np.testing.assert_array_equal(b2k.assign(X, Cn), ref, err_msg="screen_resident_a=%d" % mode)
        finally:
            screen_ctx.set_option("screen_resident_a", 0)


@pytest.mark.parametrize("d,k", [(8, 1500), (64, 2000)])
def test_screen_list_overflow_goes_to_exact_scan(b2k, oracle, screen_ctx, d, k):
    """centers on a sphere around the frames: hundreds of centers sit within the screen's margin of the best one, the
    candidate lists overflow and the frames take the exact fallback scan (both implementations), which must still
    give the oracle's labels; the statistic shows the path was really taken."""
    import ctypes as C
    import torch
    rng = np.random.RandomState(d)
    dirs = rng.randn(k, d)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    Cn = (3.0 * dirs).astype(np.float32)
    X = np.zeros((6000, d), np.float32)                      # frames AT the sphere's centre: every center ties
    X[::2] = (Cn[rng.randint(0, k, 3000)] + 0.05 * rng.randn(3000, d)).astype(np.float32)  # ordinary frames
    ref = oracle.assign(X, Cn, n_threads=8)
    for mode in (0, 1, 2):   # by queue length / CTA per frame / indexed tile kernel (3000 queued frames: mode 0 takes the tile kernel)
        screen_ctx.set_option("fallback_mode", mode)
        try:
            np.testing.assert_array_equal(b2k.assign(X, Cn), ref, err_msg="fallback_mode=%d" % mode)
        finally:
            screen_ctx.set_option("fallback_mode", 0)
    dev = torch.device("cuda", screen_ctx.device)
    dX, dC = torch.from_numpy(X).to(dev), torch.from_numpy(Cn).to(dev)
    lab = torch.empty(len(X), dtype=torch.int32, device=dev)
    b2k.check(screen_ctx.lib.b2k_dev_assign(screen_ctx.handle, C.c_void_p(dX.data_ptr()), len(X), d,
                                            C.c_void_p(dC.data_ptr()), k, 0, C.c_void_p(lab.data_ptr()), None))
    screen_ctx.sync()
    fb = screen_ctx.get_stat("screen_fallback_frames")
    print("fallback frames: %d of %d" % (fb, len(X)))
    assert fb >= 1000
    np.testing.assert_array_equal(lab.cpu().numpy(), ref)


@pytest.mark.parametrize("n,d,k", [(20000, 10, 1000), (5000, 64, 2000), (2500, 256, 1000), (3000, 3, 200), (1000, 300, 64)])
def test_screen_operand_builders_agree(b2k, oracle, screen_ctx, n, d, k):
    """both frame-operand builders (per input element / per output piece) feed the screen identical operands:
    same labels as the oracle, same candidate statistics."""
    rng = np.random.RandomState(n + d)
    X = blobs(rng, n, d, 9)
    X[5] *= 300.0                                            # an outlier sets the scale: small values get flushed
    Cn = X[rng.choice(n, k, replace=False)].copy()
    ref = oracle.assign(X, Cn, n_threads=8)
    stats = []
    for mode in (0, 1):
        screen_ctx.set_option("operand_kernel", mode)
        try:
            np.testing.assert_array_equal(b2k.assign(X, Cn), ref, err_msg="operand_kernel=%d" % mode)
            stats.append((screen_ctx.get_stat("screen_cand_chunks"), screen_ctx.get_stat("screen_fallback_frames")))
        finally:
            screen_ctx.set_option("operand_kernel", 0)
    assert stats[0] == stats[1]


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("n,d,k", [(20001, 256, 1000), (9000, 100, 2500), (4357, 300, 700)])
def test_screen_cluster_modes(b2k, oracle, screen_ctx, n, d, k, mode):
    """the streaming screen kernel as 2-CTA clusters (odd tile counts run a dummy tile in the second CTA).
    screen_cluster=2: the pair shares every center k-block through TMA multicast; screen_cluster=3: CTA-pair MMAs
    (tcgen05 cta_group::2, M=256: every SM holds its own frame rows and half of the center k-block, the leader CTA
    issues, commits are multicast, the peer's epilogue releases the accumulators on the leader's barrier) -- same
    labels as the oracle"""
    rng = np.random.RandomState(k)
    X = blobs(rng, n, d, 9)
    Cn = X[rng.choice(n, k, replace=False)].copy()
    ref = oracle.assign(X, Cn, n_threads=8)
    screen_ctx.set_option("screen_resident_a", 0)
    screen_ctx.set_option("screen_cluster", mode)
    try:
        np.testing.assert_array_equal(b2k.assign(X, Cn), ref)
    finally:
        screen_ctx.set_option("screen_cluster", 0)
