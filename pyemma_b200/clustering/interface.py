"""AbstractClustering: assign / dtrajs / transform on top of libb2k.

Mirrors pyemma/coordinates/clustering/interface.py:44-302 (reference @ 3327f28) together with the
slices of its base classes the hot path relies on:
  Estimator.estimate/fit/get_params/set_params   pyemma/_base/estimator.py:392-500
  StreamingEstimator.estimate (array -> source)   coordinates/data/_base/streaming_estimator.py:33-48
  StreamingEstimationTransformer.get_output       coordinates/data/_base/transformer.py:214-226
  DataSource.get_output (chunk loop + scatter)    coordinates/data/_base/datasource.py:334-422
  NJobsMixIn                                      pyemma/_base/parallel.py:2-73
What changes: `_transform_array` (interface.py:157-167, a deeptime ClusterModel.transform per
chunk) becomes ONE libb2k call per trajectory that streams the frames through pinned staging
onto the GPU (b2k_assign), so there is no Python work per chunk.
"""
import inspect
import logging
import os
import warnings

import numpy as np

from .. import _lib, staging
from ..data import DataInMemory, as_source


class NotConvergedWarning(RuntimeError):
    """pyemma/coordinates/data/_base/streaming_estimator.py:26"""


def get_n_jobs(logger=None):
    """pyemma/_base/parallel.py:2-38"""
    def _from_env(var):
        e = os.getenv(var, None)
        if e:
            try:
                return int(e)
            except ValueError as ve:
                if logger is not None:
                    logger.warning("could not parse env variable '%s'. Value='%s'. Error=%s.", var, e, ve)
        return None

    slurm, pyemma_n = _from_env("SLURM_CPUS_ON_NODE"), _from_env("PYEMMA_NJOBS")
    if slurm and pyemma_n:
        warnings.warn("two settings for n_jobs from environment: PYEMMA_NJOBS and SLURM_CPUS_ON_NODE. "
                      "Respecting the SLURM setting to avoid overprovisioning resources.")
    val = slurm or pyemma_n
    if not val:
        try:
            import psutil
            val = psutil.cpu_count(logical=True) or 1
        except ImportError:
            val = os.cpu_count() or 1
    return val


def index_states(dtrajs):
    """pyemma/util/discrete_trajectories.py:355-408 (all states): for every state the (traj, frame) pairs."""
    dtrajs = [np.asarray(d) for d in dtrajs]
    n_states = int(max((d.max() for d in dtrajs if d.size), default=-1)) + 1
    res = [[] for _ in range(n_states)]
    for i, d in enumerate(dtrajs):
        order = np.argsort(d, kind="stable")
        sd = d[order]
        bounds = np.searchsorted(sd, np.arange(n_states + 1))
        for s in range(n_states):
            t = order[bounds[s]:bounds[s + 1]]
            if t.size:
                res[s].append(np.stack([np.full(t.size, i, dtype=int), t.astype(int)], axis=1))
    out = np.empty(n_states, dtype=object)
    for s in range(n_states):
        out[s] = np.concatenate(res[s]) if res[s] else np.zeros((0, 2), dtype=int)
    return out


def sample_indexes_by_state(indexes, nsample, replace=True):
    """pyemma/util/discrete_trajectories.py: sample rows of each state's index table."""
    res = []
    for ind in indexes:
        if len(ind) == 0:
            res.append(np.zeros((0, 2), dtype=int))
            continue
        if replace:
            sel = np.random.choice(len(ind), nsample, replace=True)
        else:
            sel = np.random.choice(len(ind), min(nsample, len(ind)), replace=False)
        res.append(ind[sel])
    return res


class AbstractClustering:
    """Common interface of the clustering estimators (interface.py:44)."""

    def __init__(self, metric="euclidean", n_jobs=None):
        self.logger = logging.getLogger("pyemma_b200.%s[%d]" % (type(self).__name__, id(self) % 10000))
        self.metric = metric
        self._clustercenters = None
        self._previous_stride = -1
        self._dtrajs = []
        self._overwrite_dtrajs = False
        self._index_states = []
        self.n_jobs = n_jobs
        self._data_producer = None
        self._estimated = False
        self._chunksize = None
        self._in_memory = False
        self._Y = None
        self.show_progress = False

    # ---- progress reporting (pyemma/_base/progress/reporter.py: _progress_register / _progress_update) --------
    # The reference drives tqdm bars; here the counters are kept on the estimator (`progress_`: stage ->
    # [done, total, description]) and an optional user hook `progress_callback(stage, done, total)` is called.
    def _progress_register(self, amount_of_work, description="", stage=0):
        if not hasattr(self, "progress_"):
            self.progress_ = {}
        self.progress_[stage] = [0, int(amount_of_work), description]

    def _progress_update(self, numerator_increment, stage=0):
        if not hasattr(self, "progress_") or stage not in self.progress_:
            return
        ent = self.progress_[stage]
        ent[0] += int(numerator_increment)
        hook = getattr(self, "progress_callback", None)
        if hook is not None:
            hook(stage, ent[0], ent[1])

    # ---- sklearn-style parameter handling (estimator.py:460-500) ----------------------------
    @classmethod
    def _get_param_names(cls):
        sig = inspect.signature(cls.__init__)
        return sorted(p.name for p in sig.parameters.values() if p.name != "self" and p.kind != p.VAR_KEYWORD)

    def get_params(self, deep=True):
        return {k: getattr(self, k, None) for k in self._get_param_names()}

    def set_params(self, **params):
        valid = set(self._get_param_names())
        for k, v in params.items():
            if k not in valid and not hasattr(type(self), k):
                raise ValueError("Invalid parameter %s for estimator %s." % (k, type(self).__name__))
            setattr(self, k, v)
        return self

    # ---- metric / n_jobs ------------------------------------------------------------------------
    @property
    def metric(self):
        return self._metric

    @metric.setter
    def metric(self, val):
        _lib.metric_id(val)  # unknown metric -> ValueError (tests/test_regspace.py:100-105)
        self._metric = val

    @property
    def n_jobs(self):
        """Kept for API parity (parallel.py:41-73).  On the GPU path it selects nothing."""
        if not hasattr(self, "_n_jobs"):
            self._n_jobs = get_n_jobs(self.logger)
        return self._n_jobs

    @n_jobs.setter
    def n_jobs(self, val):
        if val is not None and val == 0:
            raise ValueError("n_jobs must not be 0.")
        elif val is not None and val < 0:
            warnings.warn("Negative n_jobs will likely raise in future versions, use None instead.",
                          DeprecationWarning)
            val = None
        if val is None:
            val = get_n_jobs(getattr(self, "logger", None))
        self._n_jobs = int(val)

    # ---- model params ---------------------------------------------------------------------------
    @property
    def clustercenters(self):
        """Cluster centers, always fp32 C-contiguous (interface.py:77-86)."""
        return self._clustercenters

    @clustercenters.setter
    def clustercenters(self, val):
        self._clustercenters = np.asarray(val, dtype="float32", order="C")[:] if val is not None else None
        # deliberate deviation: new centers invalidate cached dtrajs
        self._dtrajs = []
        self._index_states = []
        self._Y = None

    cluster_centers_ = clustercenters  # sk-learn alias (interface.py:78)

    @property
    def overwrite_dtrajs(self):
        return self._overwrite_dtrajs

    @overwrite_dtrajs.setter
    def overwrite_dtrajs(self, value):
        self._overwrite_dtrajs = value

    # ---- data producer / chunking (transformer.py:102-193) -------------------------------------
    @property
    def data_producer(self):
        return self._data_producer

    @data_producer.setter
    def data_producer(self, dp):
        if dp is not self._data_producer:
            self._dtrajs = []
            self._index_states = []
            self._Y = None
        self._data_producer = dp

    @property
    def chunksize(self):
        if self._data_producer is not None and self._chunksize is None:
            return self._data_producer.chunksize
        return self._chunksize if self._chunksize is not None else 1000

    @chunksize.setter
    def chunksize(self, value):
        if value is not None and int(value) < 0:
            raise ValueError("chunksize has to be non-negative")
        self._chunksize = None if value is None else int(value)
        if self._data_producer is not None and value is not None:
            self._data_producer.chunksize = int(value)

    @property
    def in_memory(self):
        return self._in_memory

    @in_memory.setter
    def in_memory(self, value):
        self._in_memory = bool(value)
        if not value:
            self._Y = None

    def number_of_trajectories(self, stride=None):
        return self.data_producer.number_of_trajectories()

    ntraj = property(number_of_trajectories)

    def trajectory_lengths(self, stride=1, skip=0):
        return self.data_producer.trajectory_lengths(stride=stride, skip=skip)

    def trajectory_length(self, itraj, stride=1, skip=0):
        return self.data_producer.trajectory_length(itraj, stride=stride, skip=skip)

    def n_frames_total(self, stride=1, skip=0):
        return int(np.sum(self.trajectory_lengths(stride, skip)))

    def dimension(self):
        """output dimension of a clustering (always 1; interface.py:169-171)"""
        return 1

    def output_type(self):
        return np.int32()

    def describe(self):
        return "[%s]" % type(self).__name__

    # ---- estimation (estimator.py:392-422, streaming_estimator.py:33-48) -----------------------
    def estimate(self, X, **params):
        chunksize = params.pop("chunksize", None)
        source = as_source(X)
        if chunksize is not None:
            source.chunksize = chunksize
        self.data_producer = source
        if params:
            self.set_params(**params)
        try:
            self._model = self._estimate(source)
        except NotConvergedWarning as ncw:
            # swallowed upstream too (streaming_estimator.py:43-47)
            self.logger.info("Presumably finished estimation. Message: %s", ncw)
            self._model = self
        self._estimated = True
        return self._model

    def fit(self, X, y=None, **kw):
        self.estimate(X, **kw)
        return self

    def fit_predict(self, X, y=None):
        """sklearn ClusterMixin (_ext/sklearn/base.py:348-366)"""
        self.fit(X)
        return self.dtrajs

    def _estimate(self, iterable, **kw):
        raise NotImplementedError

    # ---- the hot path: Voronoi assignment ------------------------------------------------------
    def _transform_array(self, X):
        """closest center index per frame as an (n, 1) int32 column (interface.py:157-167)."""
        if self.clustercenters is None or len(self.clustercenters) == 0:
            raise RuntimeError("no cluster centers: estimate first")
        X = np.asarray(X)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        if X.ndim != 2 or X.shape[1] != self.clustercenters.shape[1]:
            raise ValueError("input data has wrong shape %s for centers of dimension %d"
                             % (X.shape, self.clustercenters.shape[1]))
        X = np.require(X, dtype=np.float32, requirements="C")
        # NaN / inf frames are rejected by the library while the chunk is on the device (InvalidDataInStreamException);
        # a host-side np.isfinite pass over 1e8 floats would cost more than the assignment itself
        dtraj = _lib.assign(X, self.clustercenters, self.metric)
        return dtraj[:, None]

    def transform(self, X):
        """Transformer.transform (transformer.py:42-79): array -> (T,1); list -> list of (T_i,1)."""
        if isinstance(X, np.ndarray):
            if X.ndim in (1, 2):
                return self._transform_array(X)
            raise ValueError("Input has wrong number of dimensions (%d)" % X.ndim)
        if isinstance(X, (list, tuple)):
            return [self._transform_array(np.asarray(x)) for x in X]
        raise ValueError("Input has wrong type %s" % type(X))

    def get_output(self, dimensions=slice(0, None), stride=1, skip=0, chunk=None):
        """list of (T_i, 1) int32 matrices, one per trajectory (datasource.py:334-422).

        The reference pulls chunks through a Python loop and scatters them by it.pos; here every
        trajectory's strided frames go to libb2k in one call, which chunks internally through the
        pinned double buffer (chunk only bounds the staging size)."""
        if not self._estimated:
            raise RuntimeError("estimate first")
        if self._in_memory and self._Y is not None and stride == 1 and skip == 0:
            return self._Y
        src = self.data_producer
        if src is None:
            raise RuntimeError("no data producer set")
        out = []
        rank, ws = staging.world()
        if isinstance(src, DataInMemory) and ws > 1 and getattr(self, "distributed_assign", False):
            # explicit opt-in: every rank of the torch.distributed job must make this call (it is a collective)
            out = self._get_output_sharded(src, stride, skip, rank, ws)
        elif isinstance(src, DataInMemory):
            for x in src.data:
                out.append(self._transform_array(x[skip::stride]))
        else:
            lengths = src.trajectory_lengths(stride=stride, skip=skip)
            out = [np.empty((int(l), 1), dtype=np.int32) for l in lengths]
            cs = self.chunksize if chunk is None else chunk
            with src.iterator(stride=stride, skip=skip, chunk=cs, return_trajindex=True) as it:
                for itraj, X in it:
                    out[itraj][it.pos:it.pos + len(X)] = self._transform_array(X)
        if self._in_memory and stride == 1 and skip == 0:
            self._Y = out
        return out

    def _get_output_sharded(self, src, stride, skip, rank, ws):
        """assign / dtrajs under torchrun, opt-in through `distributed_assign = True`: frames are
        independent, so every rank assigns one contiguous range of the (strided, concatenated) frames and the int32
        labels of the shards are combined with ONE all-gather (N/ws labels per rank); every rank returns the complete
        dtrajs.  A collective: all ranks must call it, and a failure on one rank is raised on all of them."""
        import torch
        import torch.distributed as dist
        views = [x[skip::stride] for x in src.data]
        lengths = [len(v) for v in views]
        total = int(sum(lengths))
        lo, hi = staging.shard_bounds(total, rank, ws)
        dev = staging.device()
        mine = torch.empty(max(hi - lo, 0), dtype=torch.int32, device=dev)
        err, off = None, 0
        try:
            for v, L in zip(views, lengths):
                a, b = max(lo, off), min(hi, off + L)
                if a < b:
                    lab = self._transform_array(v[a - off:b - off])[:, 0]
                    mine[a - lo:b - lo] = torch.from_numpy(np.ascontiguousarray(lab)).to(dev)
                off += L
        except Exception as e:  # keep the collective sequence identical on every rank, then raise everywhere
            err = e
        flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            raise err if err is not None else RuntimeError("sharded assign failed on another rank")
        host = staging.all_gather_shards(mine, total, rank, ws).cpu().numpy()
        out, off = [], 0
        for L in lengths:
            out.append(host[off:off + L].reshape(-1, 1).copy())
            off += L
        return out

    def iterator(self, stride=1, skip=0, chunk=None, return_trajindex=True):
        """A clustering is itself a data source: yields (itraj, (n,1) int32) chunks (transformer.py:169-175)."""
        cs = self.chunksize if chunk is None else chunk
        inner = self.data_producer.iterator(stride=stride, skip=skip, chunk=cs, return_trajindex=True)

        class _Encapsulated:
            def __init__(s):
                s.pos, s.last_chunk, s.last_chunk_in_traj, s.current_trajindex = 0, False, False, 0

            def __enter__(s):
                return s

            def __exit__(s, *e):
                return False

            def __iter__(s):
                return s

            def __next__(s):
                itraj, X = next(inner)
                s.pos, s.last_chunk, s.last_chunk_in_traj = inner.pos, inner.last_chunk, inner.last_chunk_in_traj
                s.current_trajindex = itraj
                Y = self._transform_array(X)
                return (itraj, Y) if return_trajindex else Y

            def n_chunks(s):
                return inner.n_chunks()

        return _Encapsulated()

    def assign(self, X=None, stride=1):
        """interface.py:176-231"""
        if X is None:
            if self._previous_stride == stride and len(self._dtrajs) > 0:
                return self._dtrajs
            self._previous_stride = stride
            skip = self.skip if hasattr(self, "skip") else 0
            mapped = self.get_output(stride=stride, chunk=self.chunksize, skip=skip)
            self._dtrajs = [np.transpose(m)[0] for m in mapped]
            return self._dtrajs
        if stride != 1:
            raise ValueError("assign accepts either X or stride parameters, but not both. If you want to map "
                             "only a subset of your data, extract the subset yourself and pass it as X.")
        mapped = self.transform(X)
        if isinstance(mapped, np.ndarray):
            return np.transpose(mapped)[0]
        return [np.transpose(m)[0] for m in mapped]

    @property
    def dtrajs(self):
        """Discrete trajectories (interface.py:101-106)."""
        if len(self._dtrajs) == 0:
            self._dtrajs = self.assign(stride=1)
        return self._dtrajs

    @property
    def index_clusters(self):
        if len(self._dtrajs) == 0:
            self._dtrajs = self.assign()
        if len(self._index_states) == 0:
            self._index_states = index_states(self._dtrajs)
        return self._index_states

    def sample_indexes_by_cluster(self, clusters, nsample, replace=True):
        if len(self._index_states) == 0:
            self._index_states = index_states(self.dtrajs)
        return sample_indexes_by_state(self._index_states[clusters], nsample, replace=replace)

    def save_dtrajs(self, trajfiles=None, prefix="", output_dir=".", output_format="ascii", extension=".dtraj"):
        """interface.py:233-302"""
        if extension[0] != ".":
            extension = "." + extension
        if output_format == "ascii":
            def write_dtraj(fn, dt):
                with open(fn, "w") as f:
                    np.asarray(dt).tofile(f, sep="\n", format="%d")
        else:
            def write_dtraj(fn, dt):
                np.save(fn, np.asarray(dt))
        names = []
        if trajfiles is not None:
            for f in trajfiles:
                base = os.path.splitext(os.path.basename(f))[0]
                names.append(("%s_%s%s" % (prefix, base, extension)) if prefix else ("%s%s" % (base, extension)))
        else:
            for i in range(len(self.dtrajs)):
                names.append(("%s_%i%s" % (prefix, i, extension)) if prefix else (str(i) + extension))
        assert len(self.dtrajs) == len(names)
        os.makedirs(output_dir, exist_ok=True)
        for name, dt in zip(names, self.dtrajs):
            dest = os.path.join(output_dir, name)
            if os.path.exists(dest) and not self.overwrite_dtrajs:
                raise EnvironmentError('Attempted to write dtraj "%s" which already existed. To automatically'
                                       " overwrite existing files, set source.overwrite_dtrajs=True." % dest)
            write_dtraj(dest, dt)

    # pickle-able state (HDF5 serialization is out of scope)
    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop("logger", None)
        for k in [k for k in st if k.startswith("_dev_")]:
            st.pop(k)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self.logger = logging.getLogger("pyemma_b200.%s" % type(self).__name__)
