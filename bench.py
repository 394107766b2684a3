#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the k-means / assign hot path (BASELINE.json configs 1-5).

    python bench.py [--workload cfg1|cfg2|cfg3|cfg4|cfg5] --gpus N --steps K --warmup W [--impl reference]

Default workload = cfg2 (BASELINE.json configs[1], the configuration the metric is quoted on): TICA-like features,
1e7 frames x 10 dims fp32 PER GPU, k=1000, euclidean.  One "step":

  cfg2/cfg3/cfg4  one Lloyd iteration over all resident frames: assign (argmin over k centers) + centroid update +
                  cost, what deeptime kmeans.cluster_loop does per iteration (pyemma/coordinates/clustering/
                  kmeans.py:254-258).  cfg3 = 1.25e7 x 64 per GPU (1e8 over 8), k=2000; cfg4 = 2e7 x 256, k=5000 (its
                  line also carries the k-means++ seeding time, HBM-bound).
  cfg1            one whole fit: k-means++ (k=100, fixed_seed) + 10 Lloyd iterations on 1e5 x 2 three-well frames.
  cfg5            one metric='minRMSD' assign pass of 1e6 frames x 300 atoms against 1000 regspace centers (QCP kernel);
                  the line also carries the regspace dmin sweep.

`value` = frames assigned per second over the whole job (all GPUs), inputs resident in HBM.
`e2e`   = the same step through the C-ABI host-pointer entry points with the frames in PINNED HOST memory every step
          (H2D chunk by chunk on copy streams, labels D2H, the cost word read on the host).
Multi-GPU: frames shard over ranks (weak scaling), one int64 all-reduce of [k*d sums | k counts] + one cost word per
Lloyd iteration over NCCL; at N > 1 a small fixed global problem is first clustered on all ranks and compared with
hashes computed from the CPU oracle (`parity` in the line).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg1": dict(kind="fit", n=100_000, d=2, k=100, steps=40,
                 text="cfg1: three-well 1e5x2 fp32, k=100, k-means++ (fixed_seed) + 10 Lloyd iterations per step"),
    "cfg2": dict(kind="lloyd", n=10_000_000, d=10, k=1000, blobs=20, steps=1000,
                 text="cfg2: TICA-like 1e7x10 fp32 per GPU, k=1000, one Lloyd iteration per step"),
    "cfg3": dict(kind="lloyd", n=12_500_000, d=64, k=2000, blobs=50, steps=300,
                 text="cfg3: pairwise-distance-like 1.25e7x64 fp32 per GPU (1e8 over 8 GPUs), k=2000, one Lloyd iteration per step"),
    "cfg4": dict(kind="lloyd", n=20_000_000, d=256, k=5000, blobs=200, steps=40,
                 text="cfg4: 2e7x256 fp32, k=5000, one Lloyd iteration per step (+ k-means++ seeding, reported beside it)"),
    "cfg5": dict(kind="rmsd", n=1_000_000, d=900, k=1000, steps=8,
                 text="cfg5: 1e6 frames x 300 atoms fp32, metric=minRMSD, one assign pass against 1000 regspace centers per step"),
}
W = dict(WORKLOADS["cfg2"], name="cfg2")


def select_workload(name):
    global W
    W = dict(WORKLOADS[name], name=name)


def metric_text():
    if W["kind"] == "fit":
        return "frames assigned/s (k-means++ + 10 Lloyd iterations per fit; 1e5x2 fp32, k=100)"
    if W["kind"] == "rmsd":
        return "frames assigned/s (minRMSD assign pass; 1e6 frames x 300 atoms fp32, k=1000)"
    return ("frames assigned/s (Lloyd iteration: assign + centroid update + cost; %gx%d fp32 per GPU, k=%d)"
            % (W["n"], W["d"], W["k"]))


# ---------------------------------------------------------------------------------------------- synthetic data
def synth_params(seed=2):
    """TICA-like synthetic mixture: `blobs` metastable blobs, per-dim variance decaying 1,.8,.6..."""
    import numpy as np
    rng = np.random.RandomState(seed)
    d = W["d"]
    scale = np.sqrt(np.maximum(1.0 - 0.2 * np.arange(d), 0.05)).astype(np.float32)
    means = (rng.randn(W["blobs"], d) * 1.5).astype(np.float32) * scale
    return means, scale * 0.6


def synth_device(n, rank, dev):
    import torch
    means, sig = synth_params()
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    sig_d, means_d = torch.from_numpy(sig).to(dev), torch.from_numpy(means).to(dev)
    X = torch.empty((n, W["d"]), device=dev, dtype=torch.float32)
    step = max(1, (1 << 28) // W["d"])  # 1 GB pieces: no second full-size temporary
    for a in range(0, n, step):
        b = min(n, a + step)
        lab = torch.randint(0, W["blobs"], (b - a,), generator=g, device=dev)
        X[a:b].normal_(generator=g).mul_(sig_d).add_(means_d[lab])
    return X


def synth_host(n, seed):
    import numpy as np
    means, sig = synth_params()
    rng = np.random.RandomState(seed)
    out = np.empty((n, W["d"]), np.float32)
    for a in range(0, n, 1 << 20):
        b = min(n, a + (1 << 20))
        lab = rng.randint(0, W["blobs"], b - a)
        out[a:b] = means[lab] + sig * rng.standard_normal((b - a, W["d"])).astype(np.float32)
    return out


def three_well(n, seed):
    """2-D three-well trajectory with rare jumps (BASELINE configs[0])"""
    import numpy as np
    rng = np.random.RandomState(seed)
    cen = np.array([[-1.5, 0.0], [0.0, 1.2], [1.5, 0.0]])
    jump = np.flatnonzero(rng.rand(n) < 0.01)
    tgt = rng.randint(0, 3, n)
    s = np.zeros(n, dtype=np.int64)
    bounds = np.concatenate([[0], jump, [n]])
    state = 0
    for a, b in zip(bounds[:-1], bounds[1:]):
        if a > 0:
            state = tgt[a]
        s[a:b] = state
    return (cen[s] + 0.35 * rng.randn(n, 2)).astype(np.float32)


def conformations(n, n_atoms, n_templates, seed, noise=0.05):
    """randomly rotated / translated noisy copies of `n_templates` random structures (host, fp32)"""
    import numpy as np
    rng = np.random.RandomState(seed)
    T = rng.uniform(-2, 2, size=(n_templates, n_atoms, 3))
    out = np.empty((n, n_atoms * 3), np.float32)
    for s in range(0, n, 20000):
        e = min(n, s + 20000)
        q = rng.standard_normal((e - s, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = np.stack([np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
                      np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
                      np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1)], 1)
        P = T[rng.randint(0, n_templates, e - s)] + noise * rng.standard_normal((e - s, n_atoms, 3))
        P = np.einsum("nij,nkj->nki", R, P) + rng.uniform(-5, 5, (e - s, 1, 3))
        out[s:e] = P.reshape(e - s, -1)
    return out


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe): ONE `nvidia-smi -lms <period>` child
    process started before the region and read back after it.  Every query -- from a thread of this process, through NVML
    or from a child -- holds up kernel launches on this system for ~0.1 s (measured on the launch-bound cfg1 fit: 10.2 ms
    without sampling, 13.9 / 24.7 ms with 7 / 11 samples in the region), so the launch-bound workload samples once a
    second and the kernel-bound ones ten times."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=250):
        self.index, self.proc, self.period_ms = index, None, int(period_ms)
        self.rows, self.reader = [], None   # (arrival time, fields)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except Exception:
            self.proc = None
            return

        def pump():  # only reads the pipe: no driver call from this process
            for line in self.proc.stdout:
                if line.strip():
                    self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

        self.reader = threading.Thread(target=pump, daemon=True)
        self.reader.start()

    def stop(self, windows):
        """windows: [(name, t0, t1)] in time.monotonic(); the first one that holds >= 3 samples is reported (the timed
        region; for very short regions the timed region plus the e2e leg that follows it under the same load)"""
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            if self.reader is not None:
                self.reader.join(timeout=2)
        rows, used = [], None
        for name, t0, t1 in windows:
            rows = [r for (t, r) in self.rows if t0 <= t <= t1]
            used = name
            if len(rows) >= 3:
                break
        sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in rows for i in range(4) if len(r) > 2 + i and r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "window": used,
                "source": "nvidia-smi -lms %d (child process)" % self.period_ms}


# ---------------------------------------------------------------------------------------------- CPU oracle legs
def cpu_sample_frames():
    """frames per CPU step: the workload's full N where one step takes seconds on a 16-core host (cfg1, cfg2),
    else a sample with full k and d sized for ~5 s per step"""
    if W["kind"] == "fit":
        return W["n"]
    if W["kind"] == "rmsd":
        return 30_000
    return int(min(W["n"], max(20_000, 2.0e11 / (W["k"] * W["d"]))))


def cpu_oracle_rate(n_sample, steps, warmup, threads, budget_s=1e9):
    """frames/s of the oracle's step on the host cores; stops early once `budget_s` is spent (>= 1 timed step)"""
    import numpy as np
    from oracle import oracle as O
    t_start = time.perf_counter()
    times = []
    if W["kind"] == "fit":
        X = three_well(n_sample, 1)
        iters = []
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            c0 = O.kmpp_init(X, W["k"], 42, n_threads=threads, scan="blocked")
            cen, code, it, inert = O.cluster_loop(X, c0, 10, 1e-5, n_threads=threads)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                iters.append(it)
            if time.perf_counter() - t_start > budget_s and times:
                break
        return n_sample * sum(iters) / sum(times), sum(times) / len(times), len(times)
    if W["kind"] == "rmsd":
        X = conformations(n_sample, 300, 1200, 5)
        Cn = X[np.random.RandomState(5).choice(n_sample, W["k"], replace=False)].copy()
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            O.assign(X, Cn, "minRMSD", n_threads=threads)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
            if time.perf_counter() - t_start > budget_s and times:
                break
        return n_sample * len(times) / sum(times), sum(times) / len(times), len(times)
    X = synth_host(n_sample, 99)
    cen = X[np.random.RandomState(5).choice(n_sample, W["k"], replace=False)].copy()
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        newc, labels = O.kmeans_cluster(X, cen, n_threads=threads)
        O.cost(X, newc, labels, n_threads=threads)
        dt = time.perf_counter() - t0
        cen = newc
        if s >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and times:
            break
    return n_sample * len(times) / sum(times), sum(times) / len(times), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.ref_frames or cpu_sample_frames()
    rate, sec, done = cpu_oracle_rate(n_sample, args.steps, args.warmup, threads, budget_s=args.ref_budget_s)
    from oracle import oracle as O
    line = {
        "impl": "reference", "metric": metric_text(), "value": rate, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": W["text"], "frames_per_step_sample": n_sample, "d": W["d"], "k": W["k"],
                   "same_frames_as_product_arm": n_sample == W["n"], "steps_requested": args.steps},
        "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d frames of the %s workload per step, full k and d; %s"
                                   % (n_sample, W["name"], O.build_info())},
        "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- multi-GPU parity
def distributed_parity(ctx, dist, rank, ws, dev):
    """A fixed small global problem run sharded over all ranks (k-means++ sharded scan, Lloyd iterations with the
    int64 all-reduce, sharded assign), compared with hashes of the CPU oracle's result on the unsharded data (rank 0)."""
    import numpy as np
    import torch
    from pyemma_b200 import _lib
    lib = ctx.lib
    n_tot, d, k, iters = 64 * 1024, 8, 96, 4
    rng = np.random.RandomState(77)
    cen = rng.uniform(-4, 4, size=(12, d))
    X = (cen[rng.randint(0, 12, n_tot)] + 0.5 * rng.randn(n_tot, d)).astype(np.float32)
    per = n_tot // ws // 1024 * 1024  # shards start at multiples of 1024 frames (sharded k-means++ contract)
    lo = rank * per
    hi = n_tot if rank == ws - 1 else lo + per
    dX = torch.from_numpy(X[lo:hi]).to(dev)
    n_loc = hi - lo

    # --- sharded k-means++ (the library runs the rounds; this side all-reduces the two exchange buffers when asked) ---
    nf = int(lib.b2k_kmpp_exchange_floats(n_tot, d, k))
    xf = torch.zeros(max(nf, 1), dtype=torch.float32, device=dev)
    xi = torch.zeros(32, dtype=torch.int64, device=dev)
    ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MAX, 2: dist.ReduceOp.MIN}
    stream = torch.cuda.current_stream(dev)

    def exchange(_user, which, count, op):
        try:
            dist.all_reduce((xf if which == 0 else xi)[:count], op=ops[op])
            stream.synchronize()
            return 0
        except Exception:
            return 1

    fn = _lib.EXCHANGE(exchange)
    dC = torch.empty((k, d), dtype=torch.float32, device=dev)
    chosen = (C.c_int64 * k)()
    _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp_sharded(
        ctx.handle, C.c_void_p(dX.data_ptr()), n_loc, d, k, 0, 42, lo, n_tot, C.c_void_p(xf.data_ptr()), xf.numel(),
        C.c_void_p(xi.data_ptr()), fn, None, _lib.CALLBACK(0), None, C.c_void_p(dC.data_ptr()), chosen))
    picks = np.array(list(chosen), np.int64)

    # --- Lloyd iterations, all-reduce of the exact int64 sums ---
    am = torch.tensor([float(np.abs(X[lo:hi]).max())], device=dev)
    dist.all_reduce(am, op=dist.ReduceOp.MAX)
    sess = C.c_void_p()
    _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(dX.data_ptr()), n_loc, d, k, 0, n_tot,
                                        C.c_float(float(am.item())), C.byref(sess)))
    acc_len = int(lib.b2k_dev_lloyd_acc_len(sess))
    acc = torch.zeros(acc_len, dtype=torch.int64, device=dev)
    lab = torch.empty(n_loc, dtype=torch.int32, device=dev)
    cur, nxt = dC, torch.empty_like(dC)
    for _ in range(iters):
        _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), C.c_void_p(lab.data_ptr()),
                                                       C.c_void_p(acc.data_ptr())))
        dist.all_reduce(acc[:acc_len - 1])
        _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                              C.c_void_p(nxt.data_ptr())))
        cur, nxt = nxt, cur
    lib.b2k_dev_lloyd_destroy(sess)
    # --- sharded assign (dtrajs) with the final centers, labels gathered on every rank ---
    _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(dX.data_ptr()), n_loc, d, C.c_void_p(cur.data_ptr()), k, 0,
                                  C.c_void_p(lab.data_ptr()), None))
    sizes = [per] * (ws - 1) + [n_tot - per * (ws - 1)]
    parts = [torch.empty(s, dtype=torch.int32, device=dev) for s in sizes]
    dist.all_gather(parts, lab)
    dtraj = torch.cat(parts).cpu().numpy()
    centers = cur.cpu().numpy()
    # --- sharded metric='minRMSD' assign: 4096 conformations of 20 atoms against 64 of them ---
    Y = conformations(4096, 20, 12, 9)
    Cy = Y[::64].copy()
    ylo, yhi = rank * (4096 // ws), (4096 if rank == ws - 1 else (rank + 1) * (4096 // ws))
    dY, dCy = torch.from_numpy(Y[ylo:yhi]).to(dev), torch.from_numpy(Cy).to(dev)
    ylab = torch.empty(yhi - ylo, dtype=torch.int32, device=dev)
    _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(dY.data_ptr()), yhi - ylo, 60, C.c_void_p(dCy.data_ptr()), 64, 1,
                                  C.c_void_p(ylab.data_ptr()), None))
    ysizes = [4096 // ws] * (ws - 1) + [4096 - (4096 // ws) * (ws - 1)]
    yparts = [torch.empty(s, dtype=torch.int32, device=dev) for s in ysizes]
    dist.all_gather(yparts, ylab)
    ydtraj = torch.cat(yparts).cpu().numpy()
    if rank != 0:
        return None
    from oracle import oracle as O
    rc0, ridx = O.kmpp_init(X, k, 42, scan="blocked", n_threads=os.cpu_count() or 1, return_indices=True)
    rcen = rc0
    for _ in range(iters):
        rcen, _ = O.kmeans_cluster(X, rcen, n_threads=os.cpu_count() or 1, acc="f64")
    rdtraj = O.assign(X, centers, n_threads=os.cpu_count() or 1)
    h = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    return {"problem": "%d x %d, k=%d, %d Lloyd iterations, %d ranks" % (n_tot, d, k, iters, ws),
            "kmpp_picks": bool(np.array_equal(picks, ridx)), "kmpp_picks_sha1": h(picks), "oracle_picks_sha1": h(ridx),
            "centers_max_rel_err_vs_f64_oracle": float(np.abs(centers - rcen).max() / np.abs(rcen).max()),
            "centers": bool(np.abs(centers - rcen).max() <= 1e-5 * np.abs(rcen).max()),
            "dtrajs": bool(np.array_equal(dtraj, rdtraj)), "dtrajs_sha1": h(dtraj), "oracle_dtrajs_sha1": h(rdtraj),
            "minrmsd_dtrajs": bool(np.array_equal(ydtraj, O.assign(Y, Cy, "minRMSD", n_threads=os.cpu_count() or 1)))}


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (0: the workload's default, sized for >= 2 s)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU (0: the workload's)")
    ap.add_argument("--ref-frames", type=int, default=0, help="frames per step of the CPU reference arm (0: automatic)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="the reference arm stops after this many seconds")
    ap.add_argument("--cpu-frames", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity problem (N > 1)")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="b2k_ctx_set_option before the run (experiments; repeatable)")
    ap.add_argument("--engine", default="auto", choices=["auto", "direct", "screen"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--stage-mb", type=int, default=0, help="pinned staging chunk of the e2e leg in MB (0: library default)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.steps <= 0:
        args.steps = 20 if args.impl == "reference" else W["steps"]
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from pyemma_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.context(local_rank)
    lib = ctx.lib
    ctx.set_option("assign_engine", {"auto": 0, "direct": 1, "screen": 2}[args.engine])
    for opt in args.option:
        name, _, val = opt.partition("=")
        ctx.set_option(name, int(val))
    stream = torch.cuda.current_stream(dev)
    ctx.set_stream(stream.cuda_stream)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    parity = None
    if ws > 1 and not args.no_parity:
        parity = distributed_parity(ctx, dist, rank, ws, dev)

    n = args.frames or W["n"]
    D, K = W["d"], W["k"]
    extra = {}
    costs = []

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- workload set-up: step(), e2e_step()
    if W["kind"] == "lloyd":
        X = synth_device(n, rank, dev)
        cur = X[:K].clone()  # initial centers: the first K frames of rank 0's shard (identical on all ranks)
        if ws > 1:
            dist.broadcast(cur, 0)
        nxt = torch.empty_like(cur)
        absmax = C.c_float(0)
        _lib.check(lib.b2k_dev_absmax(ctx.handle, C.c_void_p(X.data_ptr()), n * D, C.byref(absmax)))
        am = torch.tensor([absmax.value], device=dev)
        if ws > 1:
            dist.all_reduce(am, op=dist.ReduceOp.MAX)
        sess = C.c_void_p()
        _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), n, D, K, _lib.EUCLIDEAN, n * ws,
                                            C.c_float(float(am.item())), C.byref(sess)))
        acc_len = int(lib.b2k_dev_lloyd_acc_len(sess))
        acc = torch.zeros(acc_len, dtype=torch.int64, device=dev)
        labels = torch.empty(n, dtype=torch.int32, device=dev)
        state = {"cur": cur, "nxt": nxt}

        def finish_step():
            c, nx = state["cur"], state["nxt"]
            if ws > 1:
                dist.all_reduce(acc[:acc_len - 1])
            _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(c.data_ptr()),
                                                  C.c_void_p(nx.data_ptr())))
            _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nx.data_ptr()), step_labels[0], C.c_void_p(acc.data_ptr())))
            if ws > 1:
                dist.all_reduce(acc[acc_len - 1:])
            # the loop's convergence test needs the cost on the host every iteration
            costs.append(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[acc_len - 1].item())))
            state["cur"], state["nxt"] = nx, c

        # the resident loop does not ask for per-iteration labels in its frame order (deeptime's cluster_loop returns
        # centers; the session keeps the labels and b2k_dev_lloyd_get_labels hands them out -- read once after the region)
        step_labels = [None]

        def step():
            step_labels[0] = None
            _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(state["cur"].data_ptr()), None,
                                                           C.c_void_p(acc.data_ptr())))
            finish_step()

        hx = hl = None

        def e2e_setup():
            nonlocal hx, hl
            hx = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
            hx.copy_(X)
            hl = torch.empty(n, dtype=torch.int32, pin_memory=True)

        def e2e_step():
            step_labels[0] = C.c_void_p(labels.data_ptr())
            _lib.check(lib.b2k_stage_lloyd_assign_accumulate(sess, C.c_void_p(hx.data_ptr()),
                                                             C.c_void_p(state["cur"].data_ptr()), C.c_void_p(X.data_ptr()),
                                                             C.c_void_p(labels.data_ptr()), C.c_void_p(hl.data_ptr()),
                                                             C.c_void_p(acc.data_ptr())))
            finish_step()

        e2e_bytes = (n * D * 4, n * 4 + 8)
        e2e_api = ("b2k_stage_lloyd_assign_accumulate (pinned host frames -> HBM chunk by chunk, each chunk assigned and "
                   "summed while the next is on the bus, labels back) + b2k_dev_lloyd_finalize/cost")
        frames_per_step = n
        # first iteration of the first session of the process (no pruning yet; includes building the fp16 operand, the
        # term probe, the cold cudaMallocs of the working buffers and CUDA's lazy module loading: a later session of the
        # same shape draws on the library's block cache and takes ~4 ms at cfg2, tools/fit_probe.py)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        step()
        f1.record(stream)
        barrier()
        extra["first_iteration_ms"] = f0.elapsed_time(f1)

    elif W["kind"] == "fit":
        import pyemma_b200 as coor
        Xh = three_well(n, 1 + rank)
        X = torch.from_numpy(Xh).to(dev)
        cen = torch.empty((K, D), dtype=torch.float32, device=dev)
        code, iters_c = C.c_int(0), C.c_int(0)
        inert = (C.c_float * 16)()
        fit_iters = []

        def step():
            _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(ctx.handle, C.c_void_p(X.data_ptr()), n, D, K, 0, 42,
                                                            _lib.KMPP_BLOCKED, _lib.CALLBACK(0), None,
                                                            C.c_void_p(cen.data_ptr()), None))
            _lib.check(lib.b2k_dev_kmeans_cluster_loop(ctx.handle, C.c_void_p(X.data_ptr()), n, D, C.c_void_p(cen.data_ptr()),
                                                       K, 0, 10, C.c_float(1e-5), _lib.CALLBACK(0), None, C.byref(code),
                                                       C.byref(iters_c), inert, 16, None))
            fit_iters.append(iters_c.value)
            costs.append(float(inert[max(iters_c.value - 1, 0)]))

        e2e_iters = []

        def e2e_setup():
            pass

        def e2e_step():
            km = coor.cluster_kmeans(Xh, k=K, max_iter=10, fixed_seed=42, kmpp_scan="blocked", tolerance=1e-5)
            e2e_iters.append(len(km.inertias_))

        e2e_bytes = (n * D * 4, K * D * 4)
        e2e_api = "pyemma_b200.cluster_kmeans(X_host, k=100, max_iter=10, fixed_seed=42) (estimator API: gather, k-means++, Lloyd loop)"
        frames_per_step = None  # n * iterations of the fit, counted after the run

    else:  # rmsd
        Xh = conformations(n, 300, 1200, 5 + rank)
        X = torch.from_numpy(Xh).to(dev)
        # regspace dmin sweep (replicas only at N > 1: center discovery is sequential in the frame order)
        sweep = []
        cen_np = None
        for dmin in (2.5, 0.8, 0.4):
            h = _lib.RegspaceHandle(D, dmin, K, "minRMSD", ctx)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            hit = False
            try:
                h.partial_fit_dev(X.data_ptr(), n)
            except _lib.MaxCentersReachedException:
                hit = True
            ctx.sync()
            sweep.append({"dmin": dmin, "centers": h.n_centers, "max_centers_hit": hit, "seconds": time.perf_counter() - t0})
            if h.n_centers == K:
                cen_np = h.centers()
            h.close()
        extra["regspace_sweep"] = sweep
        if cen_np is None:
            cen_np = Xh[np.random.RandomState(5).choice(n, K, replace=False)].copy()
        cen = torch.from_numpy(cen_np).to(dev)
        labels = torch.empty(n, dtype=torch.int32, device=dev)

        def step():
            _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, D, C.c_void_p(cen.data_ptr()), K, 1,
                                          C.c_void_p(labels.data_ptr()), None))

        hx = hl = None

        def e2e_setup():
            nonlocal hx, hl
            hx = torch.from_numpy(Xh).pin_memory()
            hl = torch.empty(n, dtype=torch.int32).pin_memory()

        def e2e_step():
            _lib.check(lib.b2k_assign(ctx.handle, C.c_void_p(hx.data_ptr()), n, D, C.c_void_p(cen_np.ctypes.data), K, 1,
                                      C.c_void_p(hl.data_ptr())))

        e2e_bytes = (n * D * 4 + K * D * 4, n * 4)
        e2e_api = "b2k_assign(metric=minRMSD) (pinned host frames streamed through the staging slots, labels back)"
        frames_per_step = n

    # ---------------------------------------------------------------- timed region
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank, 1000 if W["kind"] == "fit" else 100) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)  # the child is up (and has taken its first sample) before the timed region starts
    barrier()            # every rank (the sampler only runs on rank 0)
    ctx.set_option("profile", 1)
    l0 = _lib.launch_count()
    if W["kind"] == "fit":
        fit_iters.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.monotonic()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    t_region1 = time.monotonic()
    launches = _lib.launch_count() - l0
    if W["kind"] == "lloyd":  # the labels of the last timed iteration, in the caller's frame order
        _lib.check(lib.b2k_dev_lloyd_get_labels(sess, C.c_void_p(labels.data_ptr())))
        extra["last_labels_sum"] = int(labels.to(torch.int64).sum().item())
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if ws > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = None  # read after the e2e leg (a short timed region may hold fewer than three samples)
    if W["kind"] == "fit":
        frames_done = n * sum(fit_iters)
        extra["lloyd_iterations_per_fit"] = sum(fit_iters) / max(len(fit_iters), 1)
        extra["fits_per_s"] = args.steps / (total_ms * 1e-3)
    else:
        frames_done = frames_per_step * args.steps
    value = frames_done * ws / (total_ms * 1e-3)

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_hbm = peaks.get("hbm_gbs", 6500.0)
    src = ("measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)")
    traffic_by_class = {}
    try:  # measured DRAM bytes per launch (ncu --set full) of the kernel classes captured for this exact shape, else null
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic_by_class = tr.get("%s n=%d d=%d k=%d" % (W["name"], n, D, K), {})
    except Exception:
        pass
    traffic = traffic_by_class.get("screen")
    gemm_launches = ctx.get_stat("screen_gemm_launches")
    gemm_ms = ctx.get_stat("screen_gemm_ms_total") / gemm_launches if gemm_launches else None
    classes = {}
    for cls in ("verify", "sums", "cost", "lists"):
        cn = ctx.get_stat("prof_n_" + cls)
        if cn:
            classes[cls] = ctx.get_stat("prof_ms_" + cls) / args.steps
    ctx.set_option("profile", 0)
    if W["kind"] == "lloyd" and gemm_ms:
        step_ms = total_ms / args.steps
        flops = 2.0 * K * D * n  # algorithmic: SURVEY 8d "2*k*d flop per frame" x frames per launch
        pruned = ctx.get_stat("prune_steps") > 0
        mean_list = ctx.get_stat("prune_mean_list")
        terms = ctx.get_stat("screen_terms_used")
        kc = terms * D + 3
        k_eff = -(-kc // 16) * 16
        cols = mean_list if pruned else -(-K // 256) * 256
        issued = 2.0 * cols * k_eff * n  # MMA flops the kernel really issues (listed centers x operand columns)
        breakdown = dict(classes)
        breakdown["screen"] = gemm_ms * gemm_launches / args.steps
        breakdown["other"] = max(step_ms - sum(breakdown.values()), 0.0)
        extra["step_breakdown_ms"] = breakdown
        extra["pruning"] = {
            "active": bool(pruned), "mean_centers_per_tile_list": mean_list, "k": K, "sorts": ctx.get_stat("prune_sorts"),
            "incremental_sum_steps": ctx.get_stat("delta_steps"), "list_reuse_steps": ctx.get_stat("list_reuse_steps"),
            "labels_changed_last_step_frac": max(ctx.get_stat("labels_changed"), 0.0) / n,
            "note": ("after its first iteration the session keeps the frames sorted by label; every 128-frame tile is "
                     "screened against the centers the triangle inequality cannot exclude (exact: labels, sums and "
                     "costs are bit-identical to the unpruned iteration, tests/test_gpu_prune.py); once at most an eighth of the "
                     "labels changed in the previous iteration the exact integer member sums are updated from the changed "
                     "frames only (incremental_sum_steps; same integers as a full pass); the first (unpruned) "
                     "iteration is reported as first_iteration_ms")}
        dominant = max(breakdown, key=lambda c: breakdown[c] if c != "other" else -1.0)
        screen_roof = {"bound": "tensor", "achieved": flops / (gemm_ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                       "frac": flops / (gemm_ms * 1e-3) / 1e12 / peak_tf, "traffic": traffic,
                       "kernel": "b2k::screen_gemm_%skernel (tcgen05 distance screen, %d launches timed with CUDA events "
                                 "inside the timed region)" % ("listed_" if pruned else "", int(gemm_launches)),
                       "kernel_ms": gemm_ms, "peak_source": "bf16_tflops_sustained, " + src,
                       "algorithmic_flops_per_launch": flops, "operand_terms": terms,
                       "issued_mma_flops_per_launch": issued, "issued_frac": issued / (gemm_ms * 1e-3) / 1e12 / peak_tf,
                       "hbm": {"bytes_per_launch": n * (-(-kc // 64) * 64) * 2.0,
                               "achieved_gbs": n * (-(-kc // 64) * 64) * 2.0 / (gemm_ms * 1e-3) / 1e9, "peak_gbs": peak_hbm},
                       "note": ("achieved = SURVEY 8d algorithmic flops (2*k*d per frame) / kernel time: with exact pruning "
                                "the kernel only meets the listed centers, so this can exceed what the MMA pipe issues "
                                "(issued_frac) and, on clustered wide-row data, the dense tensor peak itself")
                               if pruned else "3-term fp16 operand split issues 3x the algorithmic flops"}
        if dominant == "screen":
            roofline = screen_roof
        else:
            # an HBM-streaming kernel leads the step: 4d+4 algorithmic bytes per frame (the frame row and its label)
            kms = breakdown[dominant]
            gbs = n * (4.0 * D + 4) / (kms * 1e-3) / 1e9
            names = {"verify": "exact verify of the screen's candidates (screen_verify_*_kernel + exact fallback)",
                     "sums": "member sums (seg_* counting sort + seg_sum_kernel, or accumulate_delta_kernel)", "cost": "cost (labeled distances + integer sum)",
                     "lists": "per-tile center lists (prune.cu)"}
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                        "traffic": traffic_by_class.get(dominant), "kernel": names[dominant], "kernel_ms": kms,
                        "peak_source": "hbm_gbs, " + src,
                        "algorithmic_bytes_per_launch": n * (4.0 * D + 4), "screen_kernel": screen_roof}
        roofline["step_frac"] = flops / (step_ms * 1e-3) / 1e12 / peak_tf
    elif W["kind"] == "lloyd":
        achieved_tf = 2.0 * K * D * n / (total_ms / args.steps * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf, "traffic": None, "kernel": "exact fp32 engine (whole step)",
                    "kernel_ms": total_ms / args.steps, "peak_source": "bf16_tflops_sustained, " + src}
    elif W["kind"] == "fit":
        # the assign kernel of one Lloyd iteration streams 4d+4 bytes per frame; the data set (0.8 MB) lives in L2, so the
        # HBM figure only says how far the launch-latency regime is from streaming speed
        it_total = max(sum(fit_iters), 1)
        gbs = (4.0 * D + 4) * n * it_total / (total_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                    "traffic": None, "kernel": "whole fit (k-means++ rounds + Lloyd iterations; launch-latency regime, data L2 resident)",
                    "kernel_ms": total_ms / args.steps, "peak_source": "hbm_gbs, " + src,
                    "launches_per_fit": launches / args.steps}
    else:
        lane_rate = ctx.get_stat("fp32_lane_instr_per_s")
        instr = 18.0 * 300 * n * K  # 6 lane sums x 3 non-fusable fp32 ops per atom and pair
        achieved = instr / (total_ms / args.steps * 1e-3) / 1e12
        roofline = {"bound": "fp32_issue", "achieved": achieved, "peak": lane_rate / 1e12, "unit": "T lane-instr/s",
                    "frac": achieved / (lane_rate / 1e12), "traffic": None, "kernel": "b2k::rmsd_slab_kernel (QCP, whole pass)",
                    "kernel_ms": total_ms / args.steps,
                    "peak_source": "fp32 CUDA-core issue rate measured by the library on this GPU (non-fusable FMUL+FADD chains)",
                    "hbm": {"achieved_gbs": 4.0 * D * n / (total_ms / args.steps * 1e-3) / 1e9, "peak_gbs": peak_hbm}}

    # cfg4: k-means++ seeding beside the Lloyd numbers (HBM-bound: k rounds over the frames)
    if W["name"] == "cfg4" and ws == 1:
        kn = min(n, 2_000_000)
        cenk = torch.empty((K, D), dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp(ctx.handle, C.c_void_p(X.data_ptr()), kn, D, K, 0, 42,
                                                        _lib.KMPP_BLOCKED, _lib.CALLBACK(0), None,
                                                        C.c_void_p(cenk.data_ptr()), None))
        torch.cuda.synchronize(dev)
        dtk = time.perf_counter() - t0
        extra["kmeans_pp"] = {"frames": kn, "k": K, "seconds": dtk, "ms_per_round": dtk / K * 1e3,
                              "hbm_frac_naive": float(K) * kn * (4 * D + 8) / dtk / 1e9 / peak_hbm,
                              "note": "exact triangle-inequality pruning skips most frame reads, so the naive-traffic fraction may exceed 1"}

    if W["name"] == "cfg4" and ws > 1:
        # sharded k-means++: the frames of all ranks, 300 rounds timed (4 small all-reduces per round)
        kk = 300
        nf = int(lib.b2k_kmpp_exchange_floats(n * ws, D, kk))
        xf = torch.zeros(max(nf, 1), dtype=torch.float32, device=dev)
        xi = torch.zeros(32, dtype=torch.int64, device=dev)
        ops = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MAX, 2: dist.ReduceOp.MIN}

        def exchange(_user, which, count, op):
            try:
                dist.all_reduce((xf if which == 0 else xi)[:count], op=ops[op])
                stream.synchronize()
                return 0
            except Exception:
                return 1

        fn = _lib.EXCHANGE(exchange)
        cenk = torch.empty((kk, D), dtype=torch.float32, device=dev)
        barrier()
        t0 = time.perf_counter()
        _lib.check(lib.b2k_dev_kmeans_init_centers_kmpp_sharded(
            ctx.handle, C.c_void_p(X.data_ptr()), n, D, kk, 0, 42, rank * n, n * ws, C.c_void_p(xf.data_ptr()), xf.numel(),
            C.c_void_p(xi.data_ptr()), fn, None, _lib.CALLBACK(0), None, C.c_void_p(cenk.data_ptr()), None))
        barrier()
        dtk = time.perf_counter() - t0
        extra["kmeans_pp_sharded"] = {"frames_total": n * ws, "k": kk, "seconds": dtk, "ms_per_round": dtk / kk * 1e3}

    # ---------------------------------------------------------------- e2e: host buffers, copies inside the timed region
    if args.stage_mb > 0:
        ctx.set_option("stage_bytes", args.stage_mb << 20)
    e2e_setup()
    e2e_steps = max(3, min(args.steps, 8 if W["kind"] != "fit" else 20))
    e2e_step()
    e2e_step()
    if W["kind"] == "fit":
        e2e_iters.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if ws > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if sampler:
        clocks = sampler.stop([("timed region", t_region0, t_region1),
                               ("timed region + e2e leg (same load)", t_region0, time.monotonic())])
    e2e_frames = n * sum(e2e_iters) if W["kind"] == "fit" else n * e2e_steps
    e2e = {"value": e2e_frames * ws / float(dt.item()), "unit": "frames/s",
           "h2d_bytes_per_step": e2e_bytes[0], "d2h_bytes_per_step": e2e_bytes[1], "api": e2e_api,
           "steps": e2e_steps, "ms_per_step": float(dt.item()) / e2e_steps * 1e3}

    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ns = args.cpu_frames or cpu_sample_frames()
        rate, sec, done = cpu_oracle_rate(ns, 2, 1, threads, budget_s=40.0)
        from oracle import oracle as O
        cpu = {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames of the same workload (full k, d), %d timed steps of %.2f s; %s"
                         % (ns, done, sec, O.build_info())}

    if W["kind"] == "lloyd":
        lib.b2k_dev_lloyd_destroy(sess)
        if ws == 1:
            # the same iteration without pruning (what the first iteration of every session costs, setup excluded)
            ctx.set_option("prune_mode", 0)
            sess = C.c_void_p()
            _lib.check(lib.b2k_dev_lloyd_create(ctx.handle, C.c_void_p(X.data_ptr()), n, D, K, _lib.EUCLIDEAN, n * ws,
                                                C.c_float(float(am.item())), C.byref(sess)))
            step()
            torch.cuda.synchronize(dev)
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record(stream)
            for _ in range(3):
                step()
            u1.record(stream)
            torch.cuda.synchronize(dev)
            extra["unpruned_iteration_ms"] = u0.elapsed_time(u1) / 3
            lib.b2k_dev_lloyd_destroy(sess)
            ctx.set_option("prune_mode", 1)
    if rank == 0:
        line = {
            "metric": metric_text(), "value": value, "unit": "frames/s", "n_gpus": ws, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": W["text"], "frames_per_gpu": n, "d": D, "k": K, "parallelism": "frames sharded x%d" % ws,
                       "l2": ("inputs (%.0f MB per GPU) larger than L2" % (n * D * 4 / 1e6) if n * D * 4 > 130e6 else
                              "inputs (%.1f MB) fit L2: launch-latency regime, stated with the number" % (n * D * 4 / 1e6)),
                       "engine": args.engine,
                       "timed_region_s": total_ms * 1e-3},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "final_cost": costs[-1] if costs else None,
        }
        if W["kind"] == "lloyd":
            line["lloyd_iters_per_s"] = args.steps / (total_ms * 1e-3)
        line.update(extra)
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
