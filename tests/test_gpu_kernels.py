"""GPU parity of the C-ABI entry points against the CPU oracle (bit-exact for labels / indices).

Every call goes through libb2k.so's extern "C" functions via ctypes (pyemma_b200/_lib.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def blobs(rng, n, d, nb, spread=5.0, sigma=1.0):
    cen = rng.uniform(-spread, spread, size=(nb, d))
    lab = rng.randint(0, nb, size=n)
    return (cen[lab] + sigma * rng.randn(n, d)).astype(np.float32)


ASSIGN_SHAPES = [(1000, 1, 7), (5000, 2, 100), (3000, 3, 33), (2000, 4, 64), (2000, 5, 17), (4000, 10, 1000),
                 (1500, 16, 300), (1000, 17, 50), (1200, 45, 15), (900, 64, 200), (700, 130, 40), (300, 256, 500),
                 (257, 300, 9), (100, 1027, 5)]


@pytest.mark.parametrize("n,d,k", ASSIGN_SHAPES)
def test_assign_bit_exact(b2k, oracle, n, d, k):
    rng = np.random.RandomState(n + d + k)
    X = blobs(rng, n, d, 8)
    C = X[rng.choice(n, k, replace=k > n)].copy()
    C[: k // 2] += 0.01 * rng.randn(k // 2, d).astype(np.float32)
    ref = oracle.assign(X, C, n_threads=4)
    got = b2k.assign(X, C)
    np.testing.assert_array_equal(got, ref)


def test_assign_ties_lowest_index(b2k, oracle):
    rng = np.random.RandomState(3)
    X = rng.randn(500, 6).astype(np.float32)
    C = np.repeat(rng.randn(10, 6).astype(np.float32), 3, axis=0)  # every center three times
    ref = oracle.assign(X, C)
    got = b2k.assign(X, C)
    np.testing.assert_array_equal(got, ref)
    assert (got % 3 == 0).all()


def test_assign_sqrt_merge_ties(b2k, oracle):
    # centers that differ in the last bits: squared distances differ but sqrt may merge them
    rng = np.random.RandomState(4)
    base = rng.randn(1, 3).astype(np.float32)
    C = np.repeat(base, 64, axis=0)
    C[:, 0] = np.nextafter(C[:, 0], np.float32(10), dtype=np.float32) if False else C[:, 0]
    for j in range(64):
        C[j, 0] = np.float32(base[0, 0]) + np.float32(j % 5) * np.spacing(np.float32(base[0, 0]))
    X = (base + rng.randn(4000, 3).astype(np.float32) * 3).astype(np.float32)
    np.testing.assert_array_equal(b2k.assign(X, C), oracle.assign(X, C))


def test_assign_chunked_streaming(b2k, oracle):
    rng = np.random.RandomState(5)
    X = blobs(rng, 50000, 7, 10)
    C = X[:50].copy()
    ctx = b2k.context()
    ctx.set_option("stage_bytes", 1 << 16)  # many chunks through the pinned double buffer
    try:
        got = b2k.assign(X, C)
    finally:
        ctx.set_option("stage_bytes", 64 << 20)
    np.testing.assert_array_equal(got, oracle.assign(X, C, n_threads=4))


def test_assign_minrmsd(b2k, oracle):
    rng = np.random.RandomState(123)
    for n_atoms, n, k in [(15, 500, 15), (3, 300, 5), (30, 400, 40), (301, 64, 7)]:
        X = rng.uniform(-50, 50, size=(n, 3 * n_atoms)).astype(np.float32)
        C = X[rng.choice(n, k, replace=False)] + rng.randn(k, 3 * n_atoms).astype(np.float32)
        np.testing.assert_array_equal(b2k.assign(X, C, "minRMSD"), oracle.assign(X, C, "minRMSD", n_threads=4))


def test_compute_metric(b2k, oracle):
    rng = np.random.RandomState(9)
    for na in (3, 15, 16, 300):
        x = rng.uniform(size=3 * na).astype(np.float32)
        y = rng.uniform(size=3 * na).astype(np.float32)
        assert b2k.compute_metric(x, y, "minRMSD").tobytes() == oracle.compute_metric(x, y, "minRMSD").tobytes()
        assert b2k.compute_metric(x, y).tobytes() == oracle.compute_metric(x, y).tobytes()
    with pytest.raises(ValueError):
        b2k.compute_metric(np.zeros(10, np.float32), np.zeros(10, np.float32), "minRMSD")


@pytest.mark.parametrize("n,d,k,metric", [(20000, 2, 100, "euclidean"), (5000, 10, 50, "euclidean"),
                                          (3000, 64, 20, "euclidean"), (600, 45, 8, "minRMSD")])
def test_lloyd_step_and_cost(b2k, oracle, n, d, k, metric):
    rng = np.random.RandomState(d)
    X = blobs(rng, n, d, 6)
    C0 = X[rng.choice(n, k, replace=False)].copy()
    refC, refL = oracle.kmeans_cluster(X, C0, metric, n_threads=4, acc="f64")
    gotC, gotL = b2k.kmeans_cluster(X, C0, metric)
    np.testing.assert_array_equal(gotL, refL)
    # GPU sums are exact fixed point -> equal to the fp64 oracle up to 1 ulp of the final rounding
    np.testing.assert_allclose(gotC, refC, rtol=2e-7, atol=1e-7 * np.abs(X).max())
    # fp32-sequential reference sums (what the reference does): within the stated 1e-5
    refC32, _ = oracle.kmeans_cluster(X, C0, metric, acc="f32seq")
    assert np.abs(gotC - refC32).max() <= 1e-5 * np.abs(refC32).max()
    c_ref = oracle.cost(X, refC, refL, metric, acc="f64")
    c_got = b2k.kmeans_cost(X, refC, refL, metric)
    assert abs(float(c_got) - float(c_ref)) <= 2e-7 * float(c_ref)


def test_cluster_loop_matches_oracle(b2k, oracle):
    rng = np.random.RandomState(11)
    X = blobs(rng, 30000, 2, 3, spread=2.0, sigma=0.4)
    C0 = oracle.kmpp_init(X, 100, 42)
    calls = []
    cen, code, it, inert = b2k.kmeans_cluster_loop(X, C0, 10, 1e-5, callback=lambda: calls.append(1))
    assert len(calls) == it - (1 if code == 0 else 0)
    # free-running vs the oracle with fp64 sums (the GPU sums are exact fixed point): same trajectory
    rcen, rcode, rit, rinert = oracle.cluster_loop(X, C0, 10, 1e-5, acc="f64")
    assert (code, it) == (rcode, rit)
    np.testing.assert_allclose(inert, rinert, rtol=2e-6)
    assert np.abs(cen - rcen).max() <= 1e-5 * np.abs(rcen).max()
    # free-running vs the reference-faithful fp32-sequential sums: same iteration count and inertias; the
    # centers drift apart by label flips of boundary frames (chaotic), so they are compared per step
    # with identical inputs in test_lloyd_step_and_cost instead.
    scen, scode, sit, sinert = oracle.cluster_loop(X, C0, 10, 1e-5, acc="f32seq")
    assert (code, it) == (scode, sit)
    np.testing.assert_allclose(inert, sinert, rtol=1e-4)
    # labels under the final centers are bit-exact when both sides use the same centers
    np.testing.assert_array_equal(b2k.assign(X, rcen), oracle.assign(X, rcen, n_threads=4))


def test_cluster_loop_converges_trivial(b2k, oracle):
    # reference known-answer test (tests/test_kmeans.py:411-426): 4 constant blobs -> exact centers
    X = np.concatenate([np.full((100, 3), v, np.float32) for v in (30, 60, 90, 120)])
    C0 = np.array([[29] * 3, [61] * 3, [88] * 3, [125] * 3], np.float32)
    cen, code, it, inert = b2k.kmeans_cluster_loop(X, C0, 10, 1e-5)
    assert code == 0
    np.testing.assert_array_equal(np.sort(cen[:, 0]), [30, 60, 90, 120])


@pytest.mark.parametrize("scan", ["serial", "blocked"])
@pytest.mark.parametrize("n,d,k,metric", [(3000, 2, 20, "euclidean"), (1500, 10, 40, "euclidean"),
                                          (1100, 70, 12, "euclidean"), (300, 45, 6, "minRMSD"),
                                          (5000, 5, 3000, "euclidean"), (2500, 33, 1200, "euclidean"),
                                          (2000, 3, 150, "euclidean"), (4000, 64, 3000, "euclidean")])
def test_kmpp_bit_exact(b2k, oracle, scan, n, d, k, metric):
    rng = np.random.RandomState(n)
    X = blobs(rng, n, d, 5)
    calls = []
    ref, ridx = oracle.kmpp_init(X, k, 42, metric, scan=scan, return_indices=True)
    got, gidx = b2k.kmeans_init_centers_kmpp(X, k, 42, metric, scan=scan, return_indices=True,
                                             callback=lambda: calls.append(1))
    np.testing.assert_array_equal(gidx, ridx)
    np.testing.assert_array_equal(got, ref)
    assert len(calls) == k


def test_kmpp_k_larger_than_n(b2k):
    with pytest.raises(ValueError):
        b2k.kmeans_init_centers_kmpp(np.zeros((5, 2), np.float32), 6, 1)


@pytest.mark.parametrize("metric,d", [("euclidean", 2), ("euclidean", 20), ("minRMSD", 30)])
def test_regspace_bit_exact(b2k, oracle, metric, d):
    rng = np.random.RandomState(d)
    X = (rng.randn(6000, d) * 2).astype(np.float32)
    dmin = {2: 0.7, 20: 9.0, 30: 3.0}[d]
    ref, ridx, rfull = oracle.regspace(X, dmin, 1000, metric, n_threads=4)
    h = b2k.RegspaceHandle(d, dmin, 1000, metric)
    for a in range(0, len(X), 1700):  # several chunks
        h.partial_fit(X[a:a + 1700])
    got = h.centers()
    h.close()
    assert not rfull
    assert len(got) == len(ref) > 3
    np.testing.assert_array_equal(got, ref)


def test_regspace_max_centers(b2k, oracle):
    rng = np.random.RandomState(0)
    X = (rng.randn(5000, 3) * 3).astype(np.float32)
    ref, ridx, rfull = oracle.regspace(X, 0.5, 37, "euclidean")
    assert rfull and len(ref) == 37
    h = b2k.RegspaceHandle(3, 0.5, 37)
    with pytest.raises(b2k.MaxCentersReachedException):
        h.partial_fit(X)
    np.testing.assert_array_equal(h.centers(), ref)
    h.close()


# ---- member sums: segmented (counting sort by label + warp run sums) == one RED per element == numpy integers -----
SEG_SHAPES = [(70000, 1, 5), (40000, 2, 100), (30000, 3, 1), (20000, 10, 1000), (9000, 17, 64), (5000, 64, 2000),
              (3000, 100, 13000), (1500, 256, 50), (800, 300, 40), (300, 1030, 7)]


@pytest.mark.parametrize("n,d,k", SEG_SHAPES)
def test_member_sums_segmented_exact(b2k, n, d, k):
    import ctypes as C
    import torch
    rng = np.random.RandomState(n + d)
    X = (rng.randn(n, d) * 3).astype(np.float32)
    lab = rng.randint(0, k, n).astype(np.int32)
    lab[::97] = -1                                   # frames without a label are skipped
    run = rng.randint(0, n - 40)
    lab[run:run + 40] = lab[run]                     # a run of equal neighbours (warp-aggregated scatter)
    ctx = b2k.context()
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    Xd, ld = torch.from_numpy(X).to(dev), torch.from_numpy(lab).to(dev)
    am = float(np.abs(X).max())
    sess = C.c_void_p()
    b2k.check(ctx.lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(Xd.data_ptr()), n, d, k, 0, n, C.c_float(am),
                                           C.byref(sess)))
    got = {}
    try:
        acc_len = int(ctx.lib.b2k_dev_lloyd_acc_len(sess))
        for mode in (0, 1, 2, 3, 4):
            ctx.set_option("accumulate_mode", mode)
            acc = torch.full((acc_len,), 7, dtype=torch.int64, device=dev)
            b2k.check(ctx.lib.b2k_dev_lloyd_accumulate(sess, C.c_void_p(ld.data_ptr()), C.c_void_p(acc.data_ptr())))
            torch.cuda.synchronize()
            got[mode] = acc.cpu().numpy()
    finally:
        ctx.set_option("accumulate_mode", 0)
        ctx.lib.b2k_dev_lloyd_destroy(sess)
    for mode in (1, 2, 3, 4):
        np.testing.assert_array_equal(got[0], got[mode], err_msg="accumulate_mode=%d" % mode)
    # independent integer reference: q such that |sum| * 2^q < 2^62 (api.cu lloyd_scales)
    def clog2(v):
        m, e = np.frexp(v)
        return int(e - 1 if m == 0.5 else e)
    q = 62 - (clog2(float(n)) + 1) - clog2(am)
    fx = np.rint(X.astype(np.float64) * 2.0 ** q).astype(np.int64)
    ok = lab >= 0
    sums = np.zeros((k, d), np.int64)
    np.add.at(sums, lab[ok], fx[ok])
    np.testing.assert_array_equal(got[0][:k * d].reshape(k, d), sums)
    np.testing.assert_array_equal(got[0][k * d:k * d + k], np.bincount(lab[ok], minlength=k))
    assert got[0][-1] == 0


@pytest.mark.parametrize("n,d,k", [(5000, 17, 30), (4001, 64, 200), (3000, 70, 9), (1000, 256, 50), (700, 301, 12)])
def test_cost_wide_rows_kernels_agree(b2k, oracle, n, d, k):
    """the 4-lanes-per-frame cost kernel and the shared-memory staged one evaluate every l_i in the reference
    order: identical fixed-point sums, and both match the oracle's cost."""
    rng = np.random.RandomState(d)
    X = blobs(rng, n, d, 6)
    Cn = X[rng.choice(n, k, replace=False)].copy() + np.float32(0.01)
    lab = oracle.assign(X, Cn, n_threads=4)
    ctx = b2k.context()
    vals = []
    try:
        for mode in (0, 1):
            ctx.set_option("cost_kernel", mode)
            vals.append(b2k.kmeans_cost(X, Cn, lab))
    finally:
        ctx.set_option("cost_kernel", 0)
    assert vals[0].tobytes() == vals[1].tobytes()
    ref = oracle.cost(X, Cn, lab, acc="f64")
    assert abs(float(vals[0]) - float(ref)) <= 2e-7 * float(ref)


@pytest.mark.parametrize("n,d,k", [(5000, 1, 7), (7001, 3, 100), (4000, 4, 33), (9000, 7, 500), (20000, 10, 1000),
                                   (6000, 12, 64), (3000, 13, 20), (5000, 16, 900)])
def test_cost_narrow_rows_fused_kernel(b2k, oracle, n, d, k):
    """d <= 16: the one-pass cost kernel (center table in shared memory, zero-padded columns, integer sum in the same
    kernel) gives the same fixed-point sum as the two-pass path (per-frame distances, then the sum) and matches the
    oracle's cost."""
    rng = np.random.RandomState(100 + d)
    X = blobs(rng, n, d, 6)
    Cn = X[rng.choice(n, k, replace=False)].copy() + np.float32(0.01)
    lab = oracle.assign(X, Cn, n_threads=4)
    ctx = b2k.context()
    vals = []
    try:
        for mode in (0, 2):
            ctx.set_option("cost_kernel", mode)
            vals.append(b2k.kmeans_cost(X, Cn, lab))
    finally:
        ctx.set_option("cost_kernel", 0)
    assert vals[0].tobytes() == vals[1].tobytes()
    ref = oracle.cost(X, Cn, lab, acc="f64")
    assert abs(float(vals[0]) - float(ref)) <= 2e-7 * float(ref)


def test_regspace_multi_piece_chunk(b2k, oracle):
    """a chunk larger than the ~64 MB piece the device pass works on (d=1100 -> 16384-frame pieces): same centers,
    same order as the oracle's single pass; and the max_centers stop still keeps exactly max_centers centers."""
    rng = np.random.RandomState(77)
    d, n = 1100, 35000
    cen = rng.uniform(-1, 1, size=(40, d))
    X = (cen[rng.randint(0, 40, n)] + 0.01 * rng.randn(n, d)).astype(np.float32)
    dmin = 5.0                                         # blob spacing ~ sqrt(2/3*1100) = 27, blob radius ~ 0.47
    h = b2k.RegspaceHandle(d, dmin, 1000)
    h.partial_fit(X)
    ref_c, ref_idx, full = oracle.regspace(X, dmin, 1000, "euclidean", n_threads=8)
    assert not full and len(ref_c) == 40
    np.testing.assert_array_equal(h.centers(), ref_c)
    h.close()
    h = b2k.RegspaceHandle(d, 0.1, 25)                 # every frame is a new center -> stops in the first piece
    with pytest.raises(b2k.MaxCentersReachedException):
        h.partial_fit(X)
    assert h.n_centers == 25
    np.testing.assert_array_equal(h.centers(), X[:25])
    h.close()


@pytest.mark.parametrize("n,d,k,shards", [(5000, 3, 40, 2), (40000, 10, 300, 3), (2100, 64, 25, 4)])
def test_kmpp_sharded_matches_single(b2k, oracle, n, d, k, shards):
    """b2k_dev_kmeans_init_centers_kmpp_sharded with the shards driven by threads of one process (each its own
    context/stream on the same GPU; the all-reduce callback is a host-side reduction behind a barrier): picks are
    bit-identical to the oracle's blocked scan, i.e. to the single-GPU call."""
    import ctypes as C
    import threading
    import torch
    from pyemma_b200.staging import shard_bounds
    rng = np.random.RandomState(n)
    X = blobs(rng, n, d, 7)
    ref, ridx = oracle.kmpp_init(X, k, 42, scan="blocked", return_indices=True)
    dev = torch.device("cuda", 0)
    lib = b2k.load()
    nf = int(lib.b2k_kmpp_exchange_floats(n, d, k))
    bar = threading.Barrier(shards)
    host = {"f": [None] * shards, "i": [None] * shards}
    results, errors = [None] * shards, []

    def run(rank):
        try:
            ctx = b2k.Context(0)
            lo, hi = shard_bounds(n, rank, shards)
            Xd = torch.from_numpy(X[lo:hi].copy()).to(dev)
            xf = torch.zeros(max(nf, 1), dtype=torch.float32, device=dev)
            xi = torch.zeros(32, dtype=torch.int64, device=dev)
            cen = torch.empty((k, d), dtype=torch.float32, device=dev)
            chosen = np.full(k, -1, np.int64)

            def exchange(_u, which, count, op):
                t, key = (xf, "f") if which == 0 else (xi, "i")
                ctx.sync()
                host[key][rank] = t[:count].cpu().numpy()
                bar.wait()
                parts = np.stack(host[key])
                red = parts.sum(0, dtype=parts.dtype) if op == 0 else (parts.max(0) if op == 1 else parts.min(0))
                bar.wait()
                t[:count].copy_(torch.from_numpy(red))
                torch.cuda.synchronize()
                return 0

            fn = b2k.EXCHANGE(exchange)
            b2k.check(lib.b2k_dev_kmeans_init_centers_kmpp_sharded(
                ctx.handle, C.c_void_p(Xd.data_ptr()), hi - lo, d, k, 0, 42, lo, n, C.c_void_p(xf.data_ptr()),
                xf.numel(), C.c_void_p(xi.data_ptr()), fn, None, b2k.CALLBACK(0), None, C.c_void_p(cen.data_ptr()),
                C.c_void_p(chosen.ctypes.data)))
            results[rank] = (cen.cpu().numpy(), chosen)
            ctx.close()
        except Exception as e:  # pragma: no cover
            errors.append(e)
            bar.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(shards)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errors, errors
    for cen, chosen in results:
        np.testing.assert_array_equal(chosen, ridx)
        np.testing.assert_array_equal(cen, ref)


@pytest.mark.parametrize("n,d,k", [(30000, 16, 400), (20000, 64, 300), (8000, 256, 150), (9000, 7, 500)])
def test_kmpp_pruning_is_exact(b2k, oracle, n, d, k):
    """triangle-inequality pruning of the candidate distances (well separated blobs: most pairs are skipped) must not
    change a single pick: pruned == unpruned == oracle"""
    rng = np.random.RandomState(d)
    cen = rng.uniform(-40, 40, size=(25, d))
    X = (cen[rng.randint(0, 25, n)] + rng.randn(n, d)).astype(np.float32)
    X[::50] = X[1::50]                                      # exact duplicates: D2 == 0 frames
    ctx = b2k.context()
    picks = []
    for prune in (2, 0):                                    # 2: pruned whatever the size (1 = automatic skips small jobs)
        ctx.set_option("kmpp_prune", prune)
        try:
            picks.append(b2k.kmeans_init_centers_kmpp(X, k, 7, scan="blocked", return_indices=True)[1])
        finally:
            ctx.set_option("kmpp_prune", 1)
    np.testing.assert_array_equal(picks[0], picks[1])
    ref = oracle.kmpp_init(X, k, 7, scan="blocked", n_threads=8, return_indices=True)[1]
    np.testing.assert_array_equal(picks[0], ref)


def test_staged_lloyd_step_equals_resident_step(b2k):
    """b2k_stage_lloyd_assign_accumulate (host frames, chunk by chunk, member sums per chunk) gives the same labels and
    the same exchange buffer, bit for bit, as the device-resident b2k_dev_lloyd_assign_accumulate"""
    import ctypes as C
    import torch
    rng = np.random.RandomState(21)
    n, d, k = 70_001, 10, 300
    X = blobs(rng, n, d, 9)
    Cn = X[rng.choice(n, k, replace=False)].copy()
    ctx = b2k.context()
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    hx = torch.from_numpy(X).pin_memory()
    dXa = torch.from_numpy(X).to(dev)
    dXb = torch.zeros_like(dXa)
    dC = torch.from_numpy(Cn).to(dev)
    out = {}
    ctx.set_option("stage_bytes", 1 << 19)  # ~13000 frames per chunk
    try:
        for name, dX in (("resident", dXa), ("staged", dXb)):
            sess = C.c_void_p()
            b2k.check(ctx.lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(dX.data_ptr()), n, d, k, 0, n,
                                                   C.c_float(float(np.abs(X).max())), C.byref(sess)))
            acc = torch.full((int(ctx.lib.b2k_dev_lloyd_acc_len(sess)),), -5, dtype=torch.int64, device=dev)
            lab = torch.empty(n, dtype=torch.int32, device=dev)
            hl = torch.empty(n, dtype=torch.int32).pin_memory()
            if name == "resident":
                b2k.check(ctx.lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(dC.data_ptr()), C.c_void_p(lab.data_ptr()),
                                                                  C.c_void_p(acc.data_ptr())))
            else:
                b2k.check(ctx.lib.b2k_stage_lloyd_assign_accumulate(sess, C.c_void_p(hx.data_ptr()), C.c_void_p(dC.data_ptr()),
                                                                    C.c_void_p(dX.data_ptr()), C.c_void_p(lab.data_ptr()),
                                                                    C.c_void_p(hl.data_ptr()), C.c_void_p(acc.data_ptr())))
            torch.cuda.synchronize()
            out[name] = (lab.cpu().numpy(), acc.cpu().numpy(), hl.numpy().copy())
            ctx.lib.b2k_dev_lloyd_destroy(sess)
    finally:
        ctx.set_option("stage_bytes", 64 << 20)
    np.testing.assert_array_equal(out["staged"][0], out["resident"][0])
    np.testing.assert_array_equal(out["staged"][1], out["resident"][1])
    np.testing.assert_array_equal(out["staged"][2], out["resident"][0])      # labels on the host too
    assert torch.equal(dXb, dXa)                                             # frames landed in the session's array
