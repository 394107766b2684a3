"""GPU: dtraj consumers (state histogram, lagged transition count matrix) against a numpy restatement and the
known answers of the count-matrix docstrings the reference relies on (msmtools / deeptime count_matrix:
dtraj [0,0,1,0,1,1,0] -> lag 1 [[1,2],[2,1]], lag 2 [[1,2],[1,1]])."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def np_count_matrix(dtrajs, lag, sliding, ns):
    Cm = np.zeros((ns, ns), np.int64)
    for dt in dtrajs:
        dt = np.asarray(dt)
        if len(dt) <= lag:
            continue
        t = np.arange(0, len(dt) - lag, 1 if sliding else lag)
        a, b = dt[t], dt[t + lag]
        ok = (a >= 0) & (b >= 0)
        np.add.at(Cm, (a[ok], b[ok]), 1)
    return Cm


def test_count_matrix_known_answers():
    from pyemma_b200 import dtraj
    dt = np.array([0, 0, 1, 0, 1, 1, 0])
    np.testing.assert_array_equal(dtraj.count_matrix(dt, 1, sparse_return=False), [[1, 2], [2, 1]])
    np.testing.assert_array_equal(dtraj.count_matrix(dt, 2, sparse_return=False), [[1, 2], [1, 1]])
    np.testing.assert_array_equal(dtraj.count_matrix(dt, 2, sliding=False, sparse_return=False), [[0, 1], [1, 1]])
    C3 = dtraj.count_matrix([dt, dt[:1]], 1, nstates=3)      # sparse by default; a 1-frame trajectory adds nothing
    assert C3.shape == (3, 3) and C3.sum() == 6
    np.testing.assert_array_equal(dtraj.count_states(dt), [4, 3])
    np.testing.assert_array_equal(dtraj.visited_set([np.array([0, 2, 2]), np.array([5])]), [0, 2, 5])
    assert dtraj.number_of_states([np.array([0, 2, 2]), np.array([5])]) == 6
    assert dtraj.number_of_states([np.array([0, 2, 2]), np.array([5])], only_used=True) == 3
    with pytest.raises(ValueError):
        dtraj.count_states(np.array([0, -1, 2]))
    np.testing.assert_array_equal(dtraj.count_states(np.array([0, -1, 2]), ignore_negative=True), [1, 0, 1])
    with pytest.raises(ValueError):
        dtraj.count_matrix(dt, 0)
    with pytest.raises(ValueError):
        dtraj.count_matrix(dt, 1, nstates=1)


@pytest.mark.parametrize("ns,lag,sliding", [(7, 1, True), (1000, 10, True), (1000, 10, False), (50, 333, True),
                                            (3000, 2, True)])
def test_count_matrix_matches_numpy(ns, lag, sliding):
    from pyemma_b200 import dtraj
    rng = np.random.RandomState(ns + lag)
    dtrajs = []
    for L in (200_003, 5, lag, lag + 1, 77_777):
        # metastable: long dwell times (warp-aggregated adds) + some unassigned (-1) frames
        s = np.repeat(rng.randint(0, ns, L // 7 + 1), 7)[:L].astype(np.int32)
        s[rng.rand(L) < 0.01] = -1
        dtrajs.append(s)
    ref = np_count_matrix(dtrajs, lag, sliding, ns)
    got = dtraj.count_matrix(dtrajs, lag, sliding=sliding, sparse_return=False, nstates=ns)
    np.testing.assert_array_equal(got, ref)
    hist = dtraj.count_states(dtrajs, ignore_negative=True)
    allv = np.concatenate(dtrajs)
    np.testing.assert_array_equal(hist, np.bincount(allv[allv >= 0]))


def test_count_matrix_from_fit_labels_on_device(b2k, oracle):
    """dtrajs -> count matrix without leaving the GPU: labels of an assign call as a CUDA tensor"""
    import ctypes as C
    import torch
    from pyemma_b200 import dtraj
    rng = np.random.RandomState(1)
    X = np.cumsum(rng.randn(50_000, 2), axis=0).astype(np.float32) * 0.05
    Cn = X[rng.choice(len(X), 40, replace=False)].copy()
    ctx = b2k.context()
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    dX, dC = torch.from_numpy(X).to(dev), torch.from_numpy(Cn).to(dev)
    lab = torch.empty(len(X), dtype=torch.int32, device=dev)
    b2k.check(ctx.lib.b2k_dev_assign(ctx.handle, C.c_void_p(dX.data_ptr()), len(X), 2, C.c_void_p(dC.data_ptr()), 40, 0,
                                     C.c_void_p(lab.data_ptr()), None))
    Cdev = dtraj.count_matrix(lab, 5, return_device=True)
    ref_lab = oracle.assign(X, Cn, n_threads=4)
    np.testing.assert_array_equal(Cdev.cpu().numpy()[:40, :40], np_count_matrix([ref_lab], 5, True, 40)[:Cdev.shape[0], :Cdev.shape[1]])
