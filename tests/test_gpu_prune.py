"""GPU: exact center pruning of the Lloyd session (csrc/prune.cu + the listed tcgen05 screen).

After the first iteration the session sorts its frames by label and lets every 128-frame tile meet only the centers the
triangle inequality cannot exclude.  Nothing observable may change: labels, the int64 exchange buffer (member sums,
counts), costs, centers and iteration counts are compared bit for bit with the unpruned session, and with the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def blobs(rng, n, d, nb, spread=4.0, sigma=0.6):
    cen = rng.uniform(-spread, spread, size=(nb, d))
    return (cen[rng.randint(0, nb, n)] + sigma * rng.randn(n, d)).astype(np.float32)


def run_session(b2k, X, C0, steps, options):
    """`steps` Lloyd iterations through the device session API; returns per-step (labels, acc, cost) and the centers"""
    import torch
    ctx = b2k.context()
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    for name, val in options.items():
        ctx.set_option(name, val)
    n, d = X.shape
    k = len(C0)
    lib = ctx.lib
    dX = torch.from_numpy(X).to(dev)
    cur = torch.from_numpy(C0).to(dev)
    nxt = torch.empty_like(cur)
    sess = C.c_void_p()
    b2k.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(dX.data_ptr()), n, d, k, 0, n,
                                       C.c_float(float(np.abs(X).max())), C.byref(sess)))
    acc = torch.zeros(int(lib.b2k_dev_lloyd_acc_len(sess)), dtype=torch.int64, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    out = []
    try:
        for _ in range(steps):
            used = cur.cpu().numpy().copy()
            b2k.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), C.c_void_p(lab.data_ptr()),
                                                          C.c_void_p(acc.data_ptr())))
            sums = acc[:-1].cpu().numpy().copy()
            b2k.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                                 C.c_void_p(nxt.data_ptr())))
            b2k.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), C.c_void_p(lab.data_ptr()),
                                             C.c_void_p(acc.data_ptr())))
            torch.cuda.synchronize()
            out.append((lab.cpu().numpy().copy(), sums, int(acc[-1].item()), used))
            cur, nxt = nxt, cur
        stats = {s: ctx.get_stat(s) for s in ("prune_steps", "prune_sorts", "prune_mean_list", "delta_steps", "labels_changed",
                                                 "list_reuse_steps")}
    finally:
        lib.b2k_dev_lloyd_destroy(sess)
        for name in options:
            ctx.set_option(name, {"prune_mode": 1, "assign_engine": b2k.ENGINE_AUTO, "delta_sums": 1,
                                  "prune_list_margin": 50}.get(name, 0))
    return out, cur.cpu().numpy(), stats


# (n, d, k, blobs, prune_mode): mode 2 = lists used when they exclude enough (few labels per tile: n/k large),
# mode 3 = listed screen with whatever the lists hold (several passes of 256 entries, lists padded with the dummy row)
SHAPES = [(60000, 10, 100, 12, 2), (30000, 10, 500, 12, 3), (40000, 2, 100, 5, 2), (9000, 64, 300, 10, 3),
          (50000, 64, 96, 12, 2), (20000, 16, 1200, 30, 3), (5000, 3, 64, 4, 3), (12000, 33, 700, 25, 3),
          (6000, 256, 520, 40, 3), (40000, 256, 64, 16, 2)]


@pytest.mark.parametrize("n,d,k,nb,mode", SHAPES)
@pytest.mark.parametrize("resort,delta,margin", [(0, 1, 50), (1, 1, 50), (0, 2, 2000), (1, 0, 0)])
def test_pruned_session_is_bit_identical(b2k, oracle, n, d, k, nb, mode, resort, delta, margin):
    # margin: per mille of the mean tile radius the center lists tolerate as center movement before they are rebuilt
    # (50 default; 2000: lists twice a tile radius wider, reused from one iteration to the next even this early; 0: rebuilt
    # every iteration)
    # delta: member sums of the pruned steps -- 1 incremental once few labels change (default), 2 always incremental
    # (from the second pruned step on, every re-sort in between included), 0 always a full pass
    rng = np.random.RandomState(n + d + k)
    X = blobs(rng, n, d, nb)
    C0 = X[rng.choice(n, k, replace=False)].copy()
    steps = 6
    base, cen0, _ = run_session(b2k, X, C0, steps, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 0})
    before = b2k.context().get_stat("prune_steps")
    before_delta = b2k.context().get_stat("delta_steps")
    before_reuse = b2k.context().get_stat("list_reuse_steps")
    got, cen1, stats = run_session(b2k, X, C0, steps, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": mode,
                                                       "prune_resort": resort, "delta_sums": delta,
                                                       "prune_list_margin": margin})
    if delta == 2:
        assert stats["delta_steps"] - before_delta >= steps - 3, stats   # every pruned step after the first one
    if delta == 0:
        assert stats["delta_steps"] == before_delta
    if margin == 2000 and k <= 8192:
        assert stats["list_reuse_steps"] - before_reuse >= 1, stats      # an iteration without a re-sort kept its lists
    if margin == 0:
        assert stats["list_reuse_steps"] == before_reuse
    if k <= 8192:   # (lists longer than the 8192-entry capacity: the session uses the full screen on the sorted frames)
        assert stats["prune_steps"] - before >= steps - 2, stats     # iterations 2.. ran on per-tile center lists
    if mode == 2:
        assert stats["prune_mean_list"] <= 0.6 * ((k + 255) // 256 * 256)   # the lists (padded to 64) did exclude centers
    for it in range(steps):
        np.testing.assert_array_equal(got[it][0], base[it][0], err_msg="labels, iteration %d" % it)
        np.testing.assert_array_equal(got[it][1], base[it][1], err_msg="member sums, iteration %d" % it)
        assert got[it][2] == base[it][2], "cost, iteration %d" % it
    np.testing.assert_array_equal(cen1, cen0)
    # and against the oracle: the labels of the last iteration for the centers that went into it, the final centers
    np.testing.assert_array_equal(got[-1][0], oracle.assign(X, got[-1][3], n_threads=8))
    rc, rcode, rit, rin = oracle.cluster_loop(X, C0, steps, 0.0, n_threads=8, acc="f64")
    if rit == steps:
        assert np.abs(cen1 - rc).max() <= 1e-5 * np.abs(rc).max()


def test_session_keeps_labels_when_not_asked(b2k):
    """dlabels = NULL: the step does not write labels in frame order; cost uses the session's copy and
    b2k_dev_lloyd_get_labels returns exactly what a step with a label array writes (pruned and unpruned sessions)"""
    import torch
    rng = np.random.RandomState(9)
    X = blobs(rng, 40000, 10, 8)
    C0 = X[rng.choice(len(X), 90, replace=False)].copy()
    ref, cen_ref, _ = run_session(b2k, X, C0, 4, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 2})
    for mode in (2, 0):
        ctx = b2k.context()
        dev = torch.device("cuda", ctx.device)
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        ctx.set_option("assign_engine", b2k.ENGINE_SCREEN)
        ctx.set_option("prune_mode", mode)
        lib = ctx.lib
        try:
            dX = torch.from_numpy(X).to(dev)
            cur = torch.from_numpy(C0).to(dev)
            nxt = torch.empty_like(cur)
            sess = C.c_void_p()
            b2k.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(dX.data_ptr()), len(X), 10, 90, 0, len(X),
                                               C.c_float(float(np.abs(X).max())), C.byref(sess)))
            acc = torch.zeros(int(lib.b2k_dev_lloyd_acc_len(sess)), dtype=torch.int64, device=dev)
            lab = torch.empty(len(X), dtype=torch.int32, device=dev)
            for it in range(4):
                b2k.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), None, C.c_void_p(acc.data_ptr())))
                sums = acc[:-1].cpu().numpy().copy()
                b2k.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                                     C.c_void_p(nxt.data_ptr())))
                b2k.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), None, C.c_void_p(acc.data_ptr())))
                b2k.check(lib.b2k_dev_lloyd_get_labels(sess, C.c_void_p(lab.data_ptr())))
                torch.cuda.synchronize()
                np.testing.assert_array_equal(lab.cpu().numpy(), ref[it][0], err_msg="labels, iteration %d mode %d" % (it, mode))
                np.testing.assert_array_equal(sums, ref[it][1])
                assert int(acc[-1].item()) == ref[it][2]
                cur, nxt = nxt, cur
            lib.b2k_dev_lloyd_destroy(sess)
        finally:
            ctx.set_option("assign_engine", b2k.ENGINE_AUTO)
            ctx.set_option("prune_mode", 1)


def test_pruned_session_unclustered_data(b2k):
    """structureless data: the lists exclude little or nothing; the session must still be exact (it falls back to the
    full screen on the sorted frames whenever the lists would not pay)"""
    rng = np.random.RandomState(1)
    X = rng.standard_normal((20000, 12)).astype(np.float32)
    C0 = X[rng.choice(len(X), 300, replace=False)].copy()
    base, cen0, _ = run_session(b2k, X, C0, 4, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 0})
    got, cen1, _ = run_session(b2k, X, C0, 4, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 2})
    for it in range(4):
        np.testing.assert_array_equal(got[it][0], base[it][0])
        np.testing.assert_array_equal(got[it][1], base[it][1])
        assert got[it][2] == base[it][2]


def test_pruned_session_duplicates_and_constant_data(b2k):
    rng = np.random.RandomState(2)
    X = blobs(rng, 15000, 6, 5)
    X[1000:3000] = X[0]                      # 2000 identical frames
    C0 = np.concatenate([X[:100], X[:100], X[5000:5100]]).astype(np.float32)   # duplicate centers: lowest index wins
    base, cen0, _ = run_session(b2k, X, C0, 4, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 0})
    got, cen1, _ = run_session(b2k, X, C0, 4, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 2})
    for it in range(4):
        np.testing.assert_array_equal(got[it][0], base[it][0])
        np.testing.assert_array_equal(got[it][1], base[it][1])
    Z = np.full((9000, 5), 1.5, np.float32)
    Cz = np.full((128, 5), 1.5, np.float32)
    base, _, _ = run_session(b2k, Z, Cz, 3, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 0})
    got, _, _ = run_session(b2k, Z, Cz, 3, {"assign_engine": b2k.ENGINE_SCREEN, "prune_mode": 2})
    for it in range(3):
        np.testing.assert_array_equal(got[it][0], base[it][0])
