"""Minimal in-memory data source + chunk iterator with the reference's streaming semantics.

Mirrors (paths relative to /root/reference/pyemma/coordinates/data):
  * DataInMemory                         data_in_memory.py:34-130   (1-D -> (N,1), >2-D flattened,
                                                                     all trajectories share ndim)
  * trajectory_length under stride/skip  _base/datasource.py:268    max((len-skip-1)//stride+1, 0)
  * n_chunks                             _base/datasource.py:271-285
  * chunk contents                       data_in_memory.py:255-277  data[skip+t*stride : skip+(t+chunk)*stride : stride]
  * iterator protocol                    _base/datasource.py:786-797,1000-1080 (pos, last_chunk,
                                                                     last_chunk_in_traj, return_trajindex,
                                                                     context manager)
  * default chunksize                    _base/iterable.py:43-61    256 MB / (dim * itemsize) frames
  * NaN/inf guard                        _base/datasource.py:1067-1075
Readers, featurizers and random-access strides are out of scope.
"""
import numpy as np

from ._lib import InvalidDataInStreamException

DEFAULT_CHUNK_BYTES = 256 * 1024 * 1024  # pyemma.cfg:40 default_chunksize = 256m
FALLBACK_CHUNKSIZE = 1000                # _base/iterable.py:27


def ensure_traj_list(X):
    """pyemma/util/types.py:473-485: ndarray -> [ndarray]; list/tuple of arrays kept."""
    if isinstance(X, np.ndarray):
        return [X]
    if isinstance(X, (list, tuple)):
        if len(X) == 0:
            raise ValueError("empty trajectory list")
        out = []
        for x in X:
            x = np.asarray(x)
            if not np.issubdtype(x.dtype, np.number):
                raise ValueError("trajectory must be numeric")
            out.append(x)
        return out
    raise ValueError("input data is neither an ndarray nor a list of ndarrays: %r" % type(X))


class DataInMemory:
    """List of trajectories held in host memory (data_in_memory.py:34)."""

    def __init__(self, data, chunksize=None):
        trajs = ensure_traj_list(data)
        fixed = []
        for x in trajs:
            if x.ndim == 1:
                x = x.reshape(-1, 1)                      # data_in_memory.py:96-97
            elif x.ndim > 2:
                x = x.reshape(x.shape[0], -1)            # data_in_memory.py:98-104
            fixed.append(x)
        ndims = {x.shape[1] for x in fixed}
        if len(ndims) != 1:
            raise ValueError("input data has different dimensions: %s" % sorted(ndims))  # :120-124
        self.data = fixed
        self._ndim = fixed[0].shape[1]
        self._lengths = [len(x) for x in fixed]
        self._chunksize = chunksize
        # per-chunk host-side NaN/inf check of the iterator (datasource.py:1067-1075).  Off by default like the reference
        # (pyemma.cfg: coordinates_check_output = False; its test-suite switches it on, conftest.py:14): the estimators
        # reject non-finite frames on the device anyway (absmax of the gathered frames, finite flag of staged assigns)
        self.check_output = False

    # -- DataSource surface -------------------------------------------------------------------
    def dimension(self):
        return self._ndim

    ndim = property(dimension)

    def output_type(self):
        return self.data[0].dtype.type()

    def number_of_trajectories(self, stride=None):
        return len(self.data)

    ntraj = property(number_of_trajectories)

    def trajectory_length(self, itraj, stride=1, skip=0):
        return max((self._lengths[itraj] - skip - 1) // int(stride) + 1, 0)

    def trajectory_lengths(self, stride=1, skip=0):
        return np.array([self.trajectory_length(i, stride, skip) for i in range(len(self.data))], dtype=int)

    def n_frames_total(self, stride=1, skip=0):
        return int(self.trajectory_lengths(stride, skip).sum())

    @property
    def default_chunksize(self):
        itemsize = self.data[0].dtype.itemsize
        dim = max(self._ndim, 1)
        cs = DEFAULT_CHUNK_BYTES // (itemsize * dim)
        return int(cs) if cs > 0 else FALLBACK_CHUNKSIZE

    @property
    def chunksize(self):
        return self.default_chunksize if self._chunksize is None else self._chunksize

    @chunksize.setter
    def chunksize(self, value):
        if value is not None and int(value) < 0:
            raise ValueError("chunksize has to be non-negative")
        self._chunksize = None if value is None else int(value)

    def n_chunks(self, chunksize, stride=1, skip=0):
        if chunksize == 0:
            return len(self.data)
        return int(sum(-(-l // chunksize) for l in self.trajectory_lengths(stride, skip)))

    def iterator(self, stride=1, skip=0, chunk=None, return_trajindex=True):
        return DataIterator(self, stride=stride, skip=skip, chunk=self.chunksize if chunk is None else chunk,
                            return_trajindex=return_trajindex)

    def get_output(self, stride=1, skip=0, chunk=None):
        return [x[skip::stride] for x in self.data]

    def ra_gather(self, ra_stride, skip=0):
        """frames of a random-access stride: an (n, 2) int array of (trajectory, frame) pairs, sorted, `skip` added to
        the frame column (data/_base/datasource.py:691-712) -> (n, dim) array in that order"""
        ra = np.asarray(ra_stride)
        if ra.ndim != 2 or ra.shape[1] != 2:
            raise ValueError("random access stride must be an (n, 2) array of (trajectory, frame) indices")
        out = np.empty((len(ra), self._ndim), dtype=self.data[0].dtype)
        for itraj in np.unique(ra[:, 0]):
            sel = ra[:, 0] == itraj
            frames = ra[sel, 1] + skip
            if len(frames) and (frames.min() < 0 or frames.max() >= self._lengths[int(itraj)]):
                raise IndexError("random access stride points outside trajectory %d" % int(itraj))
            out[sel] = self.data[int(itraj)][frames]
        if self.check_output and np.issubdtype(out.dtype, np.floating) and not np.all(np.isfinite(out)):
            raise InvalidDataInStreamException("Found invalid values in the random-access frames")
        return out


class DataIterator:
    """Chunk iterator over a DataInMemory (uniform stride only)."""

    def __init__(self, source, stride=1, skip=0, chunk=0, return_trajindex=True):
        if not isinstance(stride, (int, np.integer)) or stride < 1:
            raise ValueError("only uniform integer strides >= 1 are supported on this path")
        self.source = source
        self.stride, self.skip, self.chunksize = int(stride), int(skip), int(chunk)
        self.return_trajindex = return_trajindex
        self._itraj = 0
        self._t = 0          # position inside the current trajectory, in strided coordinates
        self.pos = 0         # strided index of the first frame of the chunk just yielded
        self.current_trajindex = 0
        self._lengths = source.trajectory_lengths(self.stride, self.skip)
        self._remaining = source.n_chunks(self.chunksize, self.stride, self.skip)
        self.last_chunk = False
        self.last_chunk_in_traj = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        return self

    def n_chunks(self):
        return self.source.n_chunks(self.chunksize, self.stride, self.skip)

    def __next__(self):
        src = self.source
        while self._itraj < len(src.data) and self._t >= self._lengths[self._itraj]:
            self._itraj += 1
            self._t = 0
        if self._itraj >= len(src.data):
            raise StopIteration
        itraj, t = self._itraj, self._t
        L = int(self._lengths[itraj])
        traj = src.data[itraj]
        if self.chunksize == 0:
            X = traj[self.skip::self.stride]
            n = L
        else:
            n = min(self.chunksize, L - t)
            a = self.skip + t * self.stride
            X = traj[a:a + n * self.stride:self.stride]
        self.pos = t
        self.current_trajindex = itraj
        self._t = t + n
        self.last_chunk_in_traj = self._t >= L
        rest = [l for l in self._lengths[itraj + 1:] if l > 0]
        self.last_chunk = self.last_chunk_in_traj and len(rest) == 0
        if src.check_output and np.issubdtype(X.dtype, np.floating) and not np.all(np.isfinite(X)):
            raise InvalidDataInStreamException(
                "Found invalid values in chunk in trajectory index %d at chunk [%d, %d]" % (itraj, t, t + n))
        return (itraj, X) if self.return_trajindex else X


def as_source(X, chunksize=None):
    """streaming_estimator.py:33-40: arrays / lists are wrapped into DataInMemory."""
    if isinstance(X, DataInMemory):
        return X
    if hasattr(X, "iterator") and hasattr(X, "trajectory_lengths") and hasattr(X, "dimension"):
        return X
    return DataInMemory(X, chunksize=chunksize)
