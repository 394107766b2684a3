"""KmeansClustering on the B200 backend.

Mirrors pyemma/coordinates/clustering/kmeans.py:48-339 (reference @ 3327f28): same constructor
parameters, fixed_seed semantics (:146-164), n_clusters default min(sqrt(N), 5000) (:294-297),
resume via clustercenters= / keep_data (:103-109, 205-207, 269-284), 'uniform' picks (:304-324).

What changes:
  * the gathered "in-memory" frame array lives in HBM (staging.gather_frames) instead of host RAM;
  * deeptime KMeans.fit (kmeans.py:254-255) -> libb2k: k-means++ (b2k_dev_kmeans_init_centers_kmpp)
    and Lloyd iterations driven through the device session API (assign+accumulate / finalize / cost);
  * frames shard over the ranks of an initialised torch.distributed job; per iteration ONE int64
    buffer [k*d sums | k counts] and one cost word are all-reduced (NCCL).  Sums are exact fixed
    point, so centers and inertias are bit-identical for any number of GPUs.
"""
import ctypes as C
import math
import os
import random

import numpy as np
import torch

from .. import _lib, staging
from .interface import AbstractClustering

__all__ = ["KmeansClustering", "MiniBatchKmeansClustering"]


class KmeansClustering(AbstractClustering):
    def __init__(self, n_clusters, max_iter=5, metric="euclidean", tolerance=1e-5, init_strategy="kmeans++",
                 fixed_seed=False, oom_strategy="memmap", stride=1, n_jobs=None, skip=0, clustercenters=None,
                 keep_data=False, kmpp_scan="auto"):
        super().__init__(metric=metric, n_jobs=n_jobs)
        if clustercenters is None:
            clustercenters = []
        self._in_memory_chunks_set = False
        self._converged = False
        self.initial_centers_ = None
        self.inertias_ = np.zeros(0, np.float32)
        self.kmpp_scan = kmpp_scan  # not a reference parameter: set directly so that subclasses need not list it
        self.set_params(n_clusters=n_clusters, max_iter=max_iter, tolerance=tolerance, init_strategy=init_strategy,
                        oom_strategy=oom_strategy, fixed_seed=fixed_seed, stride=stride, skip=skip,
                        clustercenters=clustercenters, keep_data=keep_data)

    # ---- parameters ---------------------------------------------------------------------------
    @property
    def init_strategy(self):
        return self._init_strategy

    @init_strategy.setter
    def init_strategy(self, value):
        valid = ("kmeans++", "uniform")
        if value not in valid:
            raise ValueError("invalid parameter '{}' for init_strategy. Should be one of {}".format(value, valid))
        self._init_strategy = value

    @property
    def fixed_seed(self):
        """seed for the random choice of initial centers (kmeans.py:141-164)"""
        return self._fixed_seed

    @fixed_seed.setter
    def fixed_seed(self, val):
        if isinstance(val, (bool, np.bool_)) or val is None:
            self._fixed_seed = 42 if val else random.randint(0, 2 ** 32 - 1)
        elif isinstance(val, (int, np.integer)):
            if val < 0 or val > 2 ** 32 - 1:
                self.logger.warning("seed has to be positive (or smaller than 2**32-1). Seed will be chosen randomly.")
                self.fixed_seed = False
            else:
                self._fixed_seed = int(val)
        else:
            raise ValueError("fixed seed has to be bool or integer")

    @property
    def converged(self):
        return self._converged

    def describe(self):
        return "[Kmeans, k=%i, inp_dim=%i]" % (self.n_clusters, self.data_producer.dimension())

    def _check_resume_iteration(self):
        return self.clustercenters is not None and self.clustercenters.size != 0

    # ---- estimation ---------------------------------------------------------------------------
    def _gather(self, iterable):
        """kmeans.py:286-312 + 326-338, with the array in HBM."""
        stride = self.stride if self.stride else 1
        rank, ws = staging.world()
        if (self._in_memory_chunks_set and getattr(self, "_dev_frames", None) is not None
                and self._dev_n_total == int(np.sum(iterable.trajectory_lengths(stride=stride, skip=self.skip)))):
            self.logger.debug("re-use in memory data.")
            return
        if self._check_resume_iteration() and not self._in_memory_chunks_set and not self.keep_data:
            self.logger.warning('Resuming kmeans iteration without the setting "keep_data=True", will re-create'
                                " the linear in-memory data. This is inefficient! Consider setting keep_data=True,"
                                " when you intend to resume the kmeans iteration.")
        lengths = iterable.trajectory_lengths(stride=stride, skip=self.skip)
        total = int(np.sum(lengths))
        need = total * iterable.dimension() * 4 // ws
        if self.init_strategy == "kmeans++" and not self._check_resume_iteration():
            # k-means++ scratch next to the frames: m = 2+ln k distance rows, D^2, labels, candidate masks
            need += (total // ws) * 4 * (6 + int(math.log(max(int(self.n_clusters or 1), 1))))
        free = staging.free_hbm(need)
        budget = int(os.environ.get("B2K_HBM_BUDGET_BYTES", "0")) or int(free * 0.9)
        self._host_frames = None
        if need > budget:
            # kmeans.py:181-200 spills to a host memmap (oom_strategy='memmap'); here the tier below HBM is pinned host
            # memory: the frames stay there and pass through the device once per Lloyd iteration (_lloyd_out_of_core)
            if self.oom_strategy != "memmap":
                self.logger.warning("K-means failed to load all the data (%d bytes required, %d available) into HBM. "
                                    "Consider using a larger stride or more GPUs.", need, budget)
                raise MemoryError()
            self.logger.warning("K-means: %d bytes do not fit the HBM budget of %d bytes; the frames stay in pinned host "
                                "memory and are streamed through the device every iteration.", need, budget)
            X, n_total, lo = staging.gather_frames(iterable, stride=stride, skip=self.skip, chunksize=self.chunksize,
                                                   rank=rank, world_size=ws, to_host=True)
            self._host_frames, self._dev_frames = X, None
            self._dev_n_total, self._dev_lo = n_total, lo
            self._in_memory_chunks_set = False
            return
        X, n_total, lo = staging.gather_frames(iterable, stride=stride, skip=self.skip, chunksize=self.chunksize,
                                               rank=rank, world_size=ws)
        self._dev_frames, self._dev_n_total, self._dev_lo = X, n_total, lo
        self._in_memory_chunks_set = True

    def _uniform_picks(self, iterable):
        """PyEMMA's own 'uniform' picks (kmeans.py:304-324): per trajectory ceil(len/total*k) random frame
        indices drawn with python `random` under random_seed(fixed_seed); the first n_clusters picked frames
        in iteration order become initial_centers_."""
        stride = self.stride if self.stride else 1
        lengths = [int(l) for l in iterable.trajectory_lengths(stride=stride, skip=self.skip)]
        total = sum(lengths)
        state = random.getstate()
        random.seed(self.fixed_seed)
        try:
            picks = {i: set(random.sample(list(range(0, L)), int(math.ceil((L / float(total)) * self.n_clusters))))
                     for i, L in enumerate(lengths)}
        finally:
            random.setstate(state)
        rows = []
        with iterable.iterator(stride=stride, skip=self.skip, chunk=self.chunksize, return_trajindex=True) as it:
            for itraj, X in it:
                for l in range(len(X)):
                    if len(rows) < self.n_clusters and it.pos + l in picks[itraj]:
                        rows.append(np.asarray(X[l], dtype=np.float32))
        return np.array(rows, dtype=np.float32).reshape(-1, iterable.dimension())

    def _estimate(self, iterable, **kw):
        stride = self.stride if self.stride else 1
        lengths = iterable.trajectory_lengths(stride=stride, skip=self.skip)
        total_length = int(np.sum(lengths))
        if not self.n_clusters:
            self.n_clusters = min(int(math.sqrt(total_length)), 5000)
            self.logger.info("The number of cluster centers was not specified, "
                             "using min(sqrt(N), 5000)=%s as n_clusters." % self.n_clusters)
        resume = self._check_resume_iteration()
        if resume and len(self.clustercenters) != self.n_clusters:
            raise RuntimeError("Passed clustercenters do not match n_clusters: {} vs. {}".format(
                len(self.clustercenters), self.n_clusters))
        if not resume and self.init_strategy == "uniform":
            self.initial_centers_ = self._uniform_picks(iterable)
        self._gather(iterable)
        if getattr(self, "_host_frames", None) is not None:
            return self._estimate_out_of_core(iterable, lengths, total_length, resume, stride)
        X = self._dev_frames
        n_local, d = X.shape
        k = int(self.n_clusters)
        if k > total_length:
            raise ValueError("n_clusters=%d larger than the number of frames %d" % (k, total_length))
        rank, ws = staging.world()
        ctx = _lib.context()
        dev = X.device
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        metric = _lib.metric_id(self.metric)
        lib = ctx.lib
        # NaN / inf frames are rejected here, on the device, before any seeding (the reference's guard is the
        # iterator's optional per-chunk host check, datasource.py:1067-1075)
        finite = C.c_int(1)
        _lib.check(lib.b2k_dev_all_finite(ctx.handle, C.c_void_p(X.data_ptr()), n_local * d, C.byref(finite)))
        am = torch.tensor([finite.value], dtype=torch.int32, device=dev)
        if ws > 1:
            import torch.distributed as dist
            dist.all_reduce(am, op=dist.ReduceOp.MIN)
        if int(am.item()) == 0:
            self._dev_frames = None
            self._in_memory_chunks_set = False
            raise _lib.InvalidDataInStreamException("Found invalid values (NaN/inf) in the input frames")

        # ---- initial centers ----
        if resume:
            centers = torch.from_numpy(np.array(self.clustercenters, dtype=np.float32)).to(dev)
        elif self.init_strategy == "uniform":
            # deeptime draws its own uniform picks: data[RandomState(seed).randint(0,N,k)]
            idx = np.random.RandomState(self.fixed_seed).randint(0, total_length, size=k)
            centers = self._rows_by_global_index(idx, d, dev)
        else:
            scan = self.kmpp_scan
            if scan == "auto":
                scan = "serial" if (total_length <= 20000 and ws == 1) else "blocked"
            centers = torch.empty((k, d), dtype=torch.float32, device=dev)
            # kmeans.py:241-249: stage 0 counts the k-means++ picks, stage 1 the Lloyd iterations
            if self.show_progress:
                self._progress_register(k, "initialize kmeans++ centers", stage=0)
            cb = self._callback(lambda: self._progress_update(1, stage=0)) if self.show_progress else _lib.CALLBACK(0)
            if ws > 1:
                if scan == "serial":
                    raise ValueError("kmpp_scan='serial' follows the frame order of ONE array; sharded (multi-GPU) "
                                     "k-means++ uses the blocked scan")
                self._kmpp_sharded(ctx, X, d, k, metric, total_length, centers, cb)
            else:
                _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(
                    ctx.handle, C.c_void_p(X.data_ptr()), n_local, d, k, metric, int(self.fixed_seed),
                    _lib.KMPP_SERIAL if scan == "serial" else _lib.KMPP_BLOCKED, cb, None,
                    C.c_void_p(centers.data_ptr()), None))
            self.initial_centers_ = centers.cpu().numpy()
        if resume:
            self.initial_centers_ = np.array(self.clustercenters, dtype=np.float32)

        # ---- Lloyd iterations (deeptime cluster_loop semantics) ----
        if self.show_progress:
            self._progress_register(self.max_iter, "kmeans iterations", stage=1)
        try:
            centers, converged, inertias = self._lloyd(ctx, X, centers, k, metric, total_length, rank, ws)
            self.clustercenters = centers.cpu().numpy()
            self._converged = converged
            self.inertias_ = np.asarray(inertias, dtype=np.float32)
            if stride == 1:
                # the frames the dtrajs are made of are resident right now: one more device pass against the final
                # centers (3.5 ms per 1e7 x 10 frames) instead of a second trip of every frame over PCIe when
                # `.dtrajs` is first read (interface.py:101-106 computes them lazily from the host data)
                self._dtrajs = self._resident_dtrajs(ctx, X, centers, k, metric, lengths, total_length, rank, ws,
                                                     final_labels=getattr(self, "_dev_final_labels", None))
                self._dev_final_labels = None
                self._previous_stride = 1
        finally:
            # kmeans.py:269-284: drop the big array unless the user keeps it for a resume
            if not self.keep_data or self._converged:
                self._dev_frames = None
                self._in_memory_chunks_set = False
        if self._converged:
            self.logger.debug("Cluster centers converged after %i steps.", len(self.inertias_))
        else:
            self.logger.warning("Algorithm did not reach convergence criterion"
                                " of %g in %i iterations. Consider increasing max_iter.",
                                self.tolerance, self.max_iter)
        return self

    # ---- the tier below HBM -------------------------------------------------------------------------------------------
    def _estimate_out_of_core(self, iterable, lengths, total_length, resume, stride):
        """Lloyd iterations over frames that stay in pinned host memory (kmeans.py:181-200 runs the same loop over a host
        memmap).  One pass over PCIe per iteration (b2k_stage_lloyd_pass: the cost of iteration i is taken during the
        pass of iteration i+1, when its labels and centers meet the frames again), every sum an exact integer, so
        centers, inertias, iteration count and dtrajs are bit-identical to the resident path.  k-means++ seeding needs
        k passes over all frames; out of core it runs on the largest strided subset of the frames that fits HBM
        (logged: the reference would seed on all frames)."""
        import torch.distributed as dist
        hx = self._host_frames
        n_local, d = hx.shape
        k = int(self.n_clusters)
        if k > total_length:
            raise ValueError("n_clusters=%d larger than the number of frames %d" % (k, total_length))
        rank, ws = staging.world()
        ctx = _lib.context()
        dev = staging.device(ctx)
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        metric = _lib.metric_id(self.metric)
        lib = ctx.lib
        hnp = hx.numpy()
        absmax = 0.0
        for a in range(0, n_local, 1 << 20):
            blk = hnp[a:a + (1 << 20)]
            if not np.isfinite(blk).all():
                raise _lib.InvalidDataInStreamException("Found invalid values (NaN/inf) in the input frames")
            absmax = max(absmax, float(np.abs(blk).max()))
        if resume:
            centers = torch.from_numpy(np.array(self.clustercenters, dtype=np.float32)).to(dev)
        elif self.init_strategy == "uniform":
            centers = torch.from_numpy(self.initial_centers_).to(dev)
        else:
            if ws > 1:
                raise NotImplementedError("out-of-core k-means++ seeding is single-GPU; pass clustercenters or use more GPUs")
            free = staging.free_hbm(n_local * d * 4, ctx)
            budget = int(os.environ.get("B2K_HBM_BUDGET_BYTES", "0")) or int(free * 0.9)
            per_frame = 4 * d + 4 * (6 + int(math.log(max(k, 1))))
            sub = max(1, -(-n_local * per_frame // max(budget, per_frame * k)))
            self.logger.warning("out-of-core k-means++: seeding on every %d-th frame (%d frames)", sub, len(hnp[::sub]))
            Xs = torch.from_numpy(np.ascontiguousarray(hnp[::sub])).to(dev)
            centers = torch.empty((k, d), dtype=torch.float32, device=dev)
            _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(
                ctx.handle, C.c_void_p(Xs.data_ptr()), Xs.shape[0], d, k, metric, int(self.fixed_seed), _lib.KMPP_BLOCKED,
                _lib.CALLBACK(0), None, C.c_void_p(centers.data_ptr()), None))
            del Xs
            self.initial_centers_ = centers.cpu().numpy()
        if resume:
            self.initial_centers_ = np.array(self.clustercenters, dtype=np.float32)
        am = torch.tensor([absmax, float(centers.abs().max())], dtype=torch.float32, device=dev)
        if ws > 1:
            dist.all_reduce(am, op=dist.ReduceOp.MAX)
        sess = C.c_void_p()
        _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, None, n_local, d, k, metric, total_length, C.c_float(float(am.max())),
                                            C.byref(sess)))
        try:
            acc_len = int(lib.b2k_dev_lloyd_acc_len(sess))
            acc = torch.zeros(acc_len, dtype=torch.int64, device=dev)
            labels = torch.empty(max(n_local, 1), dtype=torch.int32, device=dev)
            cur = centers.contiguous().clone()
            nxt = torch.empty_like(cur)
            it, converged, prev, have_prev = 0, False, np.float32(0), 0
            inertias = []
            tol = np.float32(self.tolerance)
            while True:
                _lib.check(lib.b2k_stage_lloyd_pass(sess, C.c_void_p(hx.data_ptr()), C.c_void_p(cur.data_ptr()),
                                                    C.c_void_p(labels.data_ptr()), have_prev, None, C.c_void_p(acc.data_ptr())))
                if ws > 1:
                    dist.all_reduce(acc)
                if have_prev:  # this pass measured the cost of the iteration that produced `cur`
                    cost = np.float32(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[acc_len - 1].item())))
                    inertias.append(cost)
                    rel = np.float32(abs(cost - prev) / cost) if cost != 0 else np.float32(0)
                    prev = cost
                    if rel <= tol:
                        converged = True
                    elif self.show_progress:
                        self._progress_update(1, stage=1)
                    it += 1
                    if not (it < self.max_iter and not converged):
                        break  # `labels` = the assignment to the final centers `cur`: the dtrajs
                _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                                      C.c_void_p(nxt.data_ptr())))
                cur, nxt = nxt, cur
                have_prev = 1
        finally:
            lib.b2k_dev_lloyd_destroy(sess)
        self.clustercenters = cur.cpu().numpy()
        self._converged = converged
        self.inertias_ = np.asarray(inertias, dtype=np.float32)
        if stride == 1:
            lab = labels[:n_local]
            if ws > 1:
                lab = staging.all_gather_shards(lab, total_length, rank, ws)
            host = lab.cpu().numpy()
            out, off = [], 0
            for L in lengths:
                out.append(host[off:off + int(L)].copy())
                off += int(L)
            self._dtrajs = out
            self._previous_stride = 1
        self._host_frames = None
        if self._converged:
            self.logger.debug("Cluster centers converged after %i steps.", len(self.inertias_))
        else:
            self.logger.warning("Algorithm did not reach convergence criterion"
                                " of %g in %i iterations. Consider increasing max_iter.",
                                self.tolerance, self.max_iter)
        return self

    def _resident_dtrajs(self, ctx, X, centers, k, metric, lengths, n_total, rank, ws, final_labels=None):
        """labels of the resident shard against `centers` (the Lloyd session's own last assignment when given, else
        b2k_dev_assign), all-gathered over the ranks and split per trajectory: the same values
        AbstractClustering.assign would produce from the host copy of the frames."""
        n_local, d = X.shape
        lab = torch.empty(max(n_local, 1), dtype=torch.int32, device=X.device)
        if final_labels is not None and final_labels.numel() == n_local:
            lab[:n_local] = final_labels
        elif n_local:
            _lib.check(ctx.lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n_local, d,
                                              C.c_void_p(centers.data_ptr()), k, metric, C.c_void_p(lab.data_ptr()), None))
        lab = lab[:n_local]
        if ws > 1:
            lab = staging.all_gather_shards(lab, n_total, rank, ws)
        host = lab.cpu().numpy()
        out, off = [], 0
        for L in lengths:
            out.append(host[off:off + int(L)].copy())
            off += int(L)
        return out

    def _callback(self, fn):
        cb = _lib.CALLBACK(lambda _u: fn())
        self._dev_cb_keepalive = cb
        return cb

    def _kmpp_sharded(self, ctx, X, d, k, metric, n_total, centers, cb):
        """k-means++ over frames sharded across the ranks: the library runs the rounds, this side only
        all-reduces the two exchange buffers when asked (4 small exchanges per round, NCCL)."""
        import torch.distributed as dist
        dev = X.device
        lib = ctx.lib
        nf = int(lib.b2k_kmpp_exchange_floats(n_total, d, k))
        xf = torch.zeros(max(nf, 1), dtype=torch.float32, device=dev)
        xi = torch.zeros(32, dtype=torch.int64, device=dev)
        ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MAX, 2: dist.ReduceOp.MIN}
        stream = torch.cuda.current_stream(dev)

        def exchange(_user, which, count, op):
            try:
                dist.all_reduce((xf if which == 0 else xi)[:count], op=ops[op])
                stream.synchronize()
                return 0
            except Exception:  # never unwind through the C frame
                self.logger.exception("k-means++ exchange failed")
                return 1

        fn = _lib.EXCHANGE(exchange)
        _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp_sharded(
            ctx.handle, C.c_void_p(X.data_ptr()), X.shape[0], d, k, metric, int(self.fixed_seed), int(self._dev_lo),
            int(n_total), C.c_void_p(xf.data_ptr()), xf.numel(), C.c_void_p(xi.data_ptr()), fn, None, cb, None,
            C.c_void_p(centers.data_ptr()), None))

    def _rows_by_global_index(self, idx, d, dev):
        """rows of the (sharded) frame array by global frame index, replicated on every rank"""
        import torch.distributed as dist
        rank, ws = staging.world()
        lo = self._dev_lo
        n_local = self._dev_frames.shape[0]
        idx_t = torch.as_tensor(idx, dtype=torch.int64, device=dev)
        rows = torch.zeros((len(idx), d), dtype=torch.float32, device=dev)
        mine = (idx_t >= lo) & (idx_t < lo + n_local)
        if mine.any():
            rows[mine] = self._dev_frames[idx_t[mine] - lo]
        if ws > 1:
            dist.all_reduce(rows)  # every row is non-zero on exactly one rank
        return rows

    def _lloyd(self, ctx, X, centers, k, metric, n_total, rank, ws):
        import torch.distributed as dist
        lib = ctx.lib
        dev = X.device
        n_local, d = X.shape
        absmax = C.c_float(0)
        _lib.check(lib.b2k_dev_absmax(ctx.handle, C.c_void_p(X.data_ptr()), n_local * d, C.byref(absmax)))
        am = torch.tensor([absmax.value, float(centers.abs().max())], dtype=torch.float32, device=dev)
        if ws > 1:
            dist.all_reduce(am, op=dist.ReduceOp.MAX)
        absmax_g = float(am.max())
        if not math.isfinite(absmax_g):
            raise _lib.InvalidDataInStreamException("Found invalid values (NaN/inf) in the input frames")
        sess = C.c_void_p()
        _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), n_local, d, k, metric, n_total,
                                            C.c_float(absmax_g), C.byref(sess)))
        try:
            acc_len = int(lib.b2k_dev_lloyd_acc_len(sess))
            acc = torch.zeros(acc_len, dtype=torch.int64, device=dev)
            labels = torch.empty(max(n_local, 1), dtype=torch.int32, device=dev)
            cur = centers.contiguous().clone()
            nxt = torch.empty_like(cur)
            it, converged, prev = 0, False, np.float32(0)
            inertias = []
            tol = np.float32(self.tolerance)
            while True:
                # (no per-iteration labels in frame order: the session keeps them, kmeans.py:254-258 returns centers only)
                _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), None,
                                                               C.c_void_p(acc.data_ptr())))
                if ws > 1:
                    dist.all_reduce(acc[:acc_len - 1])
                _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                                      C.c_void_p(nxt.data_ptr())))
                _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), None, C.c_void_p(acc.data_ptr())))
                if ws > 1:
                    dist.all_reduce(acc[acc_len - 1:])
                cost = np.float32(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[acc_len - 1].item())))
                cur, nxt = nxt, cur
                inertias.append(cost)
                rel = np.float32(abs(cost - prev) / cost) if cost != 0 else np.float32(0)
                prev = cost
                if rel <= tol:
                    converged = True
                elif self.show_progress:
                    # deeptime cluster_loop calls callback_loop after every iteration that did not converge
                    # (kmeans.py:246,254-255 registers it as progress stage 1)
                    self._progress_update(1, stage=1)
                it += 1
                if not (it < self.max_iter and not converged):
                    break
            # dtrajs: the assignment to the FINAL centers, taken by the session itself while it is alive (its fp16 operand,
            # the sorted copy of the frames and the center lists are all in place) instead of a fresh one-shot assign
            # that would rebuild them; the exact integer sums this call also produces are simply not used
            if n_local:
                _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()),
                                                               C.c_void_p(labels.data_ptr()), C.c_void_p(acc.data_ptr())))
            self._dev_final_labels = labels[:n_local]
            return cur, converged, inertias
        finally:
            lib.b2k_dev_lloyd_destroy(sess)


class MiniBatchKmeansClustering(KmeansClustering):
    """Mini-batch k-means (pyemma/coordinates/clustering/kmeans.py:341-447).

    Every pass draws, per trajectory, floor(len/total * ceil(total*batch_size)) frame indices without replacement
    with `np.random.choice` (sorted), gathers those frames and hands them to deeptime's
    `MiniBatchKMeans.partial_fit` (:428): ONE Lloyd step on the batch (`kmeans.cluster`) followed by
    `kmeans.cost_function`, i.e. the batch is re-assigned to the NEW centers and the squared distances are summed;
    the pass loop stops when the relative change of that cost is <= tolerance (:432-440).  Here the batch is
    uploaded once per pass and both halves run in libb2k (Lloyd session + assign + cost).  The first pass without
    given centers seeds them with k-means++ on the batch, like `_pick_initial_centers`.
    """

    def __init__(self, n_clusters, max_iter=5, metric="euclidean", tolerance=1e-5, init_strategy="kmeans++",
                 batch_size=0.2, oom_strategy="memmap", fixed_seed=False, stride=None, n_jobs=None, skip=0,
                 clustercenters=None, keep_data=False):
        if stride is not None:
            raise ValueError("stride is a dummy value in MiniBatch Kmeans")
        if batch_size > 1:
            raise ValueError("batch_size should be less or equal to 1, but was %s" % batch_size)
        if keep_data:
            raise ValueError("keep_data is a dummy value in MiniBatch Kmeans")
        super().__init__(n_clusters, max_iter, metric, tolerance, init_strategy, False, oom_strategy, stride=stride,
                         n_jobs=n_jobs, skip=skip, clustercenters=clustercenters, keep_data=False)
        if fixed_seed is not False:
            self.fixed_seed = fixed_seed  # the reference always seeds randomly (:361 passes False); kept settable
        self.batch_size = batch_size

    def _draw_mini_batch_sample(self):
        """kmeans.py:369-384 -- (n_samples, 2) array of (trajectory, sorted frame index)"""
        ra = np.empty((self._n_samples, 2), dtype=int)
        offset = 0
        for idx, traj_len in enumerate(self._traj_lengths):
            m = self._n_samples_traj[idx]
            ra[offset:offset + m, 0] = idx
            ra[offset:offset + m, 1] = np.sort(np.random.choice(traj_len, m, replace=False))
            offset += m
        return ra

    def _init_batches(self, iterable):
        """kmeans.py:386-399"""
        self._traj_lengths = [int(l) for l in iterable.trajectory_lengths(skip=self.skip)]
        self._total_length = sum(self._traj_lengths)
        samples = int(math.ceil(self._total_length * self.batch_size))
        self._n_samples = 0
        self._n_samples_traj = {}
        for idx, traj_len in enumerate(self._traj_lengths):
            m = int(math.floor(traj_len / float(self._total_length) * samples))
            self._n_samples_traj[idx] = m
            self._n_samples += m

    def _estimate(self, iterable, **kw):
        if not hasattr(iterable, "ra_gather"):
            raise NotImplementedError("mini-batch k-means needs a random-access data source (DataInMemory)")
        self.stride = None
        self._init_batches(iterable)
        if not self.n_clusters:
            self.n_clusters = min(int(math.sqrt(self._total_length)), 5000)
        k = int(self.n_clusters)
        if self._n_samples < k:
            raise ValueError("mini batch of %d frames is smaller than n_clusters=%d" % (self._n_samples, k))
        ctx = _lib.context()
        lib = ctx.lib
        dev = staging.device(ctx)
        ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        metric = _lib.metric_id(self.metric)
        d = iterable.dimension()
        centers = None
        if self._check_resume_iteration():
            if len(self.clustercenters) != k:
                raise RuntimeError("Passed clustercenters do not match n_clusters: {} vs. {}".format(
                    len(self.clustercenters), k))
            centers = torch.from_numpy(np.array(self.clustercenters, dtype=np.float32)).to(dev)
            self.initial_centers_ = np.array(self.clustercenters, dtype=np.float32)
        self._converged = False
        inertias = []
        self._draw_mini_batch_sample()  # the reference draws one sample to open its iterator (:413) and discards it
        i_pass, prev_cost = 0, 0.0
        pinned = torch.empty((self._n_samples, d), dtype=torch.float32, pin_memory=True)
        X = torch.empty((self._n_samples, d), dtype=torch.float32, device=dev)
        labels = torch.empty(self._n_samples, dtype=torch.int32, device=dev)
        while not (self._converged or i_pass + 1 > self.max_iter):
            ra = self._draw_mini_batch_sample()
            np.copyto(pinned.numpy(), iterable.ra_gather(ra, skip=self.skip), casting="unsafe")
            X.copy_(pinned, non_blocking=True)
            n = self._n_samples
            if centers is None:  # _pick_initial_centers on the first batch
                centers = torch.empty((k, d), dtype=torch.float32, device=dev)
                if self.init_strategy == "uniform":
                    idx = np.random.RandomState(self.fixed_seed).randint(0, n, size=k)
                    centers.copy_(X[torch.as_tensor(idx, device=dev)])
                else:
                    _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(
                        ctx.handle, C.c_void_p(X.data_ptr()), n, d, k, metric, int(self.fixed_seed), _lib.KMPP_BLOCKED,
                        _lib.CALLBACK(0), None, C.c_void_p(centers.data_ptr()), None))
                self.initial_centers_ = centers.cpu().numpy()
            absmax = C.c_float(0)
            _lib.check(lib.b2k_dev_absmax(ctx.handle, C.c_void_p(X.data_ptr()), n * d, C.byref(absmax)))
            amax = max(absmax.value, float(centers.abs().max()))
            if not math.isfinite(amax):
                raise _lib.InvalidDataInStreamException("Found invalid values (NaN/inf) in the input frames")
            sess = C.c_void_p()
            _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), n, d, k, metric, n,
                                                C.c_float(amax), C.byref(sess)))
            try:
                acc = torch.zeros(int(lib.b2k_dev_lloyd_acc_len(sess)), dtype=torch.int64, device=dev)
                newc = torch.empty_like(centers)
                _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(centers.data_ptr()),
                                                               C.c_void_p(labels.data_ptr()), C.c_void_p(acc.data_ptr())))
                _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(centers.data_ptr()),
                                                      C.c_void_p(newc.data_ptr())))
                # cost_function: assignments against the NEW centers, then the squared distances
                _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, d, C.c_void_p(newc.data_ptr()), k,
                                              metric, C.c_void_p(labels.data_ptr()), None))
                _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(newc.data_ptr()), C.c_void_p(labels.data_ptr()),
                                                  C.c_void_p(acc.data_ptr())))
                cost = float(np.float32(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[-1].item()))))
            finally:
                lib.b2k_dev_lloyd_destroy(sess)
            centers = newc
            self.clustercenters = centers.cpu().numpy()
            inertias.append(cost)
            rel_change = abs(cost - prev_cost) / cost if cost != 0.0 else 0.0
            prev_cost = cost
            if rel_change <= self.tolerance:
                self._converged = True
                self.logger.info("Cluster centers converged after %i steps.", i_pass + 1)
            i_pass += 1
        self.inertias_ = np.asarray(inertias, dtype=np.float32)
        if not self._converged:
            self.logger.info("Algorithm did not reach convergence criterion"
                             " of %g in %i iterations. Consider increasing max_iter.", self.tolerance, self.max_iter)
        return self
