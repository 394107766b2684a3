"""CPU: host-side logic of the estimator layer (no kernel launches) + 2-rank gloo exchange logic."""
import os
import warnings

import numpy as np
import pytest

import pyemma_b200 as p
from pyemma_b200.clustering.interface import index_states, sample_indexes_by_state
from pyemma_b200.data import DataInMemory, DataIterator
from pyemma_b200.staging import shard_bounds


def test_fixed_seed_semantics():
    # kmeans.py:146-164 / tests/test_kmeans.py:130-137
    assert p.KmeansClustering(5, fixed_seed=True).fixed_seed == 42
    a, b = p.KmeansClustering(5, fixed_seed=False).fixed_seed, p.KmeansClustering(5, fixed_seed=None).fixed_seed
    assert 0 <= a < 2 ** 32 and 0 <= b < 2 ** 32
    assert p.KmeansClustering(5, fixed_seed=463498).fixed_seed == 463498
    assert 0 <= p.KmeansClustering(5, fixed_seed=-1).fixed_seed < 2 ** 32
    assert 0 <= p.KmeansClustering(5, fixed_seed=2 ** 32 + 5).fixed_seed < 2 ** 32
    with pytest.raises(ValueError):
        p.KmeansClustering(5, fixed_seed="x")


def test_param_validation():
    with pytest.raises(ValueError):
        p.KmeansClustering(5, init_strategy="random")
    with pytest.raises(ValueError):
        p.KmeansClustering(5, metric="manhattan")  # tests/test_regspace.py:100-105
    with pytest.raises(ValueError):
        p.RegularSpaceClustering(dmin=-1.0)
    with pytest.raises(ValueError):
        p.RegularSpaceClustering(dmin=1.0, max_centers=-3)
    with pytest.raises(ValueError):
        p.cluster_regspace(np.zeros((3, 2)), dmin=-1)  # api.py:2044-2045
    with pytest.raises(ValueError):
        p.AssignCenters(np.zeros(3))
    with pytest.raises(ValueError):
        p.assign_to_centers(np.zeros((3, 2)))
    r = p.RegularSpaceClustering(dmin=0.5, max_centers=7)
    assert r.n_clusters == 7
    r.n_clusters = 9
    assert r.max_centers == 9
    km = p.KmeansClustering(10)
    assert set(km.get_params()) >= {"n_clusters", "max_iter", "metric", "tolerance", "init_strategy", "fixed_seed",
                                    "oom_strategy", "stride", "n_jobs", "skip", "clustercenters", "keep_data"}
    km.set_params(max_iter=3, tolerance=1e-3)
    assert (km.max_iter, km.tolerance) == (3, 1e-3)
    assert km.dimension() == 1 and km.output_type() == np.int32()
    c = p.AssignCenters(np.zeros((4, 3)))
    assert c.clustercenters.dtype == np.float32 and c.clustercenters.flags.c_contiguous
    with pytest.raises(ValueError):  # assign.py:90-98
        c.data_producer = DataInMemory(np.zeros((5, 2)))


def test_n_jobs_resolution(monkeypatch):
    # _base/parallel.py:2-73, tests/test_assign.py:199-230
    monkeypatch.setenv("PYEMMA_NJOBS", "3")
    monkeypatch.delenv("SLURM_CPUS_ON_NODE", raising=False)
    assert p.KmeansClustering(2).n_jobs == 3
    monkeypatch.setenv("SLURM_CPUS_ON_NODE", "5")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert p.KmeansClustering(2).n_jobs == 5
    assert p.KmeansClustering(2, n_jobs=2).n_jobs == 2
    with pytest.raises(ValueError):
        p.KmeansClustering(2, n_jobs=0)
    with pytest.warns(DeprecationWarning):
        p.KmeansClustering(2, n_jobs=-1)


def test_data_in_memory_semantics():
    a = np.arange(20.0)
    src = DataInMemory([a, np.arange(7.0), np.arange(3.0)])
    assert src.dimension() == 1 and src.ntraj == 3
    np.testing.assert_array_equal(src.trajectory_lengths(), [20, 7, 3])
    np.testing.assert_array_equal(src.trajectory_lengths(stride=3, skip=2), [6, 2, 1])  # max((len-skip-1)//stride+1,0)
    np.testing.assert_array_equal(src.trajectory_lengths(stride=1, skip=5), [15, 2, 0])
    assert src.n_chunks(4) == 5 + 2 + 1 and src.n_chunks(0) == 3
    with pytest.raises(ValueError):
        DataInMemory([np.zeros((3, 2)), np.zeros((3, 4))])
    assert DataInMemory(np.zeros((5, 2, 3))).dimension() == 6
    # default chunk size: 256 MB / (dim*itemsize)  (iterable.py:43-61)
    assert DataInMemory(np.zeros((5, 10), np.float32)).chunksize == 256 * 1024 * 1024 // 40
    assert DataInMemory(np.zeros((5, 10), np.float64)).chunksize == 256 * 1024 * 1024 // 80


def test_iterator_chunks_and_positions():
    src = DataInMemory([np.arange(20.0), np.arange(100.0, 107.0)])
    seen = []
    with src.iterator(stride=3, skip=2, chunk=4) as it:
        for itraj, X in it:
            seen.append((itraj, it.pos, X.ravel().tolist(), it.last_chunk_in_traj, it.last_chunk))
    assert seen == [(0, 0, [2, 5, 8, 11], False, False), (0, 4, [14, 17], True, False),
                    (1, 0, [102, 105], True, True)]
    whole = list(src.iterator(chunk=0, return_trajindex=False))
    assert [len(x) for x in whole] == [20, 7]
    src2 = DataInMemory(np.array([1.0, np.nan, 2.0]))
    assert src2.check_output is False                       # pyemma.cfg default: coordinates_check_output = False
    assert len(list(src2.iterator(chunk=2))) == 2
    src2.check_output = True                                # what the reference test-suite switches on (conftest.py:14)
    with pytest.raises(Exception, match="invalid"):
        list(src2.iterator(chunk=2))


def test_index_states_reference_table():
    # clustering/tests/test_cluster_samples.py:41-60
    dtrajs = [np.array(t) for t in ([0, 1, 2], [3, 4, 5], [6, 7, 8], [0, 1, 2], [3, 4, 5], [6, 7, 8])]
    ref = [[[0, 0], [3, 0]], [[0, 1], [3, 1]], [[0, 2], [3, 2]], [[1, 0], [4, 0]], [[1, 1], [4, 1]],
           [[1, 2], [4, 2]], [[2, 0], [5, 0]], [[2, 1], [5, 1]], [[2, 2], [5, 2]]]
    idx = index_states(dtrajs)
    for cc in range(9):
        np.testing.assert_array_equal(idx[cc], ref[cc])
    for ii, s in enumerate(sample_indexes_by_state(idx, 10)):
        assert all(dtrajs[a][b] == ii for a, b in s)


def test_shard_bounds_cover_everything():
    for n, w in [(10, 3), (100_000_000, 8), (7, 8), (0, 2)]:
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(l % 1024 == 0 or l == n for l, h in b)          # k-means++ tree alignment
        if n >= 1024 * w * 8:
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1024 * w


def _gloo_worker(rank, ws, port, q):
    """Each rank owns a frame shard; the exchange is ONE int64 all-reduce of [sums|counts] + one cost word.
    Fixed-point integer sums make the result identical to the single-shard result, bit for bit."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    rng = np.random.RandomState(0)
    n, d, k = 5003, 4, 7
    X = rng.randn(n, d).astype(np.float32)
    labels = rng.randint(0, k - 1, n)  # cluster k-1 stays empty
    lo, hi = shard_bounds(n, rank, ws)
    q_bits = 62 - 14 - 3
    scale = float(2 ** q_bits)

    def local_acc(a, b):
        acc = np.zeros(k * d + k + 1, np.int64)
        fx = np.rint(X[a:b].astype(np.float64) * scale).astype(np.int64)
        np.add.at(acc[:k * d].reshape(k, d), labels[a:b], fx)
        np.add.at(acc[k * d:k * d + k], labels[a:b], 1)
        return acc

    acc = torch.from_numpy(local_acc(lo, hi))
    dist.all_reduce(acc[:-1])
    full = local_acc(0, n)
    ok = bool((acc.numpy()[:-1] == full[:-1]).all())
    old = np.full((k, d), 9.0, np.float32)
    cnt = acc.numpy()[k * d:k * d + k]
    newc = np.where(cnt[:, None] > 0, (acc.numpy()[:k * d].reshape(k, d) / scale) / np.maximum(cnt, 1)[:, None], old)
    ok = ok and bool((newc[k - 1] == 9.0).all())  # empty cluster keeps its old center
    ok = ok and np.allclose(newc[0], X[labels == 0].astype(np.float64).mean(0), rtol=1e-12)
    # ---- sharded k-means++ exchange (kmpp.cu kmpp_run_blocked): height-10 sums of the balanced fp32 sum tree are
    # owned by exactly one rank (shards start at multiples of 1024), so all-reduce(sum) over zero-filled buffers
    # reproduces the single-array tree level bit for bit, and all-reduce(max) publishes the owner's candidate
    def tree10(v):  # balanced pairwise fp32 sums of aligned 1024-blocks (zero padded)
        v = np.concatenate([v, np.zeros((-len(v)) % 1024, np.float32)]).reshape(-1, 1024)
        while v.shape[1] > 1:
            v = (v[:, 0::2] + v[:, 1::2]).astype(np.float32)
        return v[:, 0]
    D2 = (rng.rand(n).astype(np.float32)) ** 2
    n10 = -(-n // 1024)
    buf = torch.zeros(n10, dtype=torch.float32)
    if hi > lo:
        assert lo % 1024 == 0
        part = tree10(D2[lo:hi])
        buf[lo // 1024: lo // 1024 + len(part)] = torch.from_numpy(part)
    dist.all_reduce(buf)
    ok = ok and bool((buf.numpy() == tree10(D2)).all())
    cand = torch.tensor([4000 if lo <= 4000 < hi else -1, 17 if lo <= 17 < hi else -1], dtype=torch.int64)
    dist.all_reduce(cand, op=dist.ReduceOp.MAX)
    ok = ok and cand.tolist() == [4000, 17]
    dist.destroy_process_group()
    q.put((rank, ok))


def test_two_rank_exchange_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


def test_mini_batch_parameters_and_sampling():
    """kmeans.py:346-399: dummy-parameter errors, per-trajectory sample counts, sorted draws without replacement"""
    with pytest.raises(ValueError):
        p.MiniBatchKmeansClustering(5, stride=2)
    with pytest.raises(ValueError):
        p.MiniBatchKmeansClustering(5, batch_size=1.5)
    with pytest.raises(ValueError):
        p.MiniBatchKmeansClustering(5, keep_data=True)
    mb = p.MiniBatchKmeansClustering(5, batch_size=0.25)
    src = DataInMemory([np.zeros((100, 2)), np.zeros((60, 2)), np.zeros((7, 2))])
    mb.skip = 0
    mb._init_batches(src)
    total, samples = 167, int(np.ceil(167 * 0.25))
    assert [mb._n_samples_traj[i] for i in range(3)] == [int(np.floor(l / total * samples)) for l in (100, 60, 7)]
    np.random.seed(3)
    ra = mb._draw_mini_batch_sample()
    assert ra.shape == (mb._n_samples, 2)
    for i, L in enumerate((100, 60, 7)):
        fr = ra[ra[:, 0] == i, 1]
        assert len(fr) == mb._n_samples_traj[i] and len(set(fr)) == len(fr) and (np.diff(fr) > 0).all() and fr.max() < L
    X = [np.arange(200.).reshape(100, 2), np.arange(120.).reshape(60, 2) + 1000, np.arange(14.).reshape(7, 2) + 5000]
    got = DataInMemory(X).ra_gather(ra)
    want = np.array([X[i][f] for i, f in ra])
    np.testing.assert_array_equal(got, want)
    with pytest.raises(IndexError):
        DataInMemory(X).ra_gather(np.array([[2, 7]]))
