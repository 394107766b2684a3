"""Chunk hand-off: host chunks -> pinned staging -> device, and the device-resident frame shard.

Replaces the host-side gather of the reference (KmeansClustering._collect_data,
pyemma/coordinates/clustering/kmeans.py:326-338, which copies every chunk with an fp64->fp32 cast
into one host (N,d) array, and _init_in_memory_chunks :170-200): here the "in-memory array" lives
in HBM.  Every chunk is cast to fp32 straight into one of two pinned staging tensors and sent
with an async copy on a side stream, so the cast of chunk c+1 overlaps the DMA of chunk c.

PyTorch is used for what it is good at here -- device / pinned allocations, streams and (in
distributed runs) the NCCL process group; all arithmetic is in libb2k.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def device(ctx=None):
    ctx = ctx or _lib.context()
    if not torch.cuda.is_available():
        raise RuntimeError("pyemma_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", ctx.device)


def world():
    """(rank, world_size) of the torch.distributed job, (0, 1) when not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def free_hbm(need=0, ctx=None):
    """Free bytes of the context's device.  libb2k keeps the working buffers of finished sessions in a block cache
    (api.cu, option "cache_mb") that only its own allocations can draw on; when `need` bytes would not fit beside it,
    the cache goes back to the driver first, so a torch allocation of that size sees the room."""
    ctx = ctx or _lib.context()
    dev = device(ctx)
    free, _tot = torch.cuda.mem_get_info(dev)
    if need > 0.9 * free and ctx.get_stat("cache_bytes") > 0:
        ctx.set_option("cache_release", 1)
        free, _tot = torch.cuda.mem_get_info(dev)
    return free


SHARD_ALIGN = 1024  # frames; a height-10 node of the k-means++ sum tree never straddles two shards


def shard_bounds(n_total, rank, world_size):
    """Contiguous frame range [lo, hi) owned by `rank` (frames are independent: SURVEY 8e).  Shards are equal up
    to SHARD_ALIGN frames and every shard starts at a multiple of SHARD_ALIGN (the sharded k-means++ needs that);
    trailing ranks of a tiny data set may own nothing."""
    n_total, world_size = int(n_total), int(world_size)
    per = -(-n_total // world_size)                       # ceil
    per = -(-per // SHARD_ALIGN) * SHARD_ALIGN if world_size > 1 else per
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def all_gather_shards(local, n_total, rank, world_size):
    """Concatenate the per-rank pieces of a frame-sharded 1-D device tensor (shard_bounds layout) on every rank:
    one all-gather of ceil(n/ws) elements per rank instead of an all-reduce over a zero-filled length-n buffer."""
    import torch.distributed as dist
    lo0, hi0 = shard_bounds(n_total, 0, world_size)
    per = max(hi0 - lo0, 1)
    send = torch.zeros(per, dtype=local.dtype, device=local.device)
    send[:local.numel()] = local
    recv = torch.empty(per * world_size, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send)
    return recv[:n_total]  # shards are contiguous, equal-sized except the trailing ones: padding only sits at the end


class PinnedStager:
    """Two pinned fp32 staging slots + a copy stream."""

    def __init__(self, rows, dim, dev):
        self.rows, self.dim, self.dev = int(rows), int(dim), dev
        self.slots = [torch.empty((self.rows, self.dim), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self.views = [s.numpy() for s in self.slots]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.stream = torch.cuda.Stream(device=dev)
        self.i = 0
        self.bytes_h2d = 0

    def send(self, X, dst):
        """cast+copy the host chunk X (n,dim) into device tensor view dst (n,dim), asynchronously."""
        n = X.shape[0]
        off = 0
        while off < n:
            m = min(self.rows, n - off)
            s = self.i & 1
            self.events[s].synchronize()  # slot free again?
            np.copyto(self.views[s][:m], X[off:off + m], casting="unsafe")
            with torch.cuda.stream(self.stream):
                dst[off:off + m].copy_(self.slots[s][:m], non_blocking=True)
                self.events[s].record(self.stream)
            self.bytes_h2d += m * self.dim * 4
            self.i += 1
            off += m

    def finish(self):
        self.stream.synchronize()


def gather_frames(source, stride=1, skip=0, chunksize=None, ctx=None, rank=0, world_size=1, progress=None, to_host=False):
    """All (strided, skipped) frames of `source` owned by this rank as ONE fp32 (n_local, d) CUDA tensor -- or, with
    to_host (the tier below HBM), as one PINNED host tensor that the library streams through the device.

    Returns (tensor, n_total, lo) where lo is the global index of the first local frame."""
    dev = device(ctx)
    d = source.dimension()
    lengths = source.trajectory_lengths(stride=stride, skip=skip)
    n_total = int(np.sum(lengths))
    lo, hi = shard_bounds(n_total, rank, world_size)
    n_local = hi - lo
    if to_host:
        out = torch.empty((n_local, d), dtype=torch.float32, pin_memory=True)
        view = out.numpy()
        if getattr(source, "project_host_chunk", None) is not None:
            raise NotImplementedError("fused projection sources need the resident path")
        cs = source.chunksize if chunksize is None else chunksize
        t = 0
        with source.iterator(stride=stride, skip=skip, chunk=cs, return_trajindex=True) as it:
            for _itraj, X in it:
                a, b = t, t + len(X)
                t = b
                if progress is not None:
                    progress()
                if b <= lo or a >= hi:
                    continue
                xa, xb = max(a, lo) - a, min(b, hi) - a
                np.copyto(view[a + xa - lo:a + xb - lo], X[xa:xb], casting="unsafe")
        return out, n_total, lo
    out = torch.empty((n_local, d), dtype=torch.float32, device=dev)
    cs = source.chunksize if chunksize is None else chunksize
    stage_rows = max(1, min(n_local if n_local else 1, (64 << 20) // max(4 * d, 1)))
    stager = PinnedStager(stage_rows, d, dev)
    # a LinearProjection source: project every RAW chunk on the device while it is staged (transform.py)
    fused = getattr(source, "project_host_chunk", None)
    it_source = source.raw if fused is not None else source
    t = 0
    with it_source.iterator(stride=stride, skip=skip, chunk=cs, return_trajindex=True) as it:
        for _itraj, X in it:
            a, b = t, t + len(X)
            t = b
            if progress is not None:
                progress()
            if b <= lo or a >= hi:
                continue
            xa, xb = max(a, lo) - a, min(b, hi) - a
            part, dst = X[xa:xb], out[a + xa - lo:a + xb - lo]
            if fused is not None:
                fused(part, dst, ctx)
            elif part.dtype == np.float32 and part.flags.c_contiguous and part.shape[0] > 0:
                # fp32 chunks need no cast: libb2k's own staging (several bounce-copy threads for pageable memory)
                c = ctx or _lib.context()
                _lib.check(c.lib.b2k_upload(c.handle, C.c_void_p(part.ctypes.data), C.c_void_p(dst.data_ptr()),
                                            part.nbytes))
            else:
                stager.send(part, dst)
    stager.finish()
    torch.cuda.current_stream(dev).wait_stream(stager.stream)
    return out, n_total, lo
