#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the k-means / assign hot path.

Workload (BASELINE.json configs[1]): TICA-projected features, 1e7 frames x 10 dims fp32 PER GPU,
k=1000, euclidean.  One "step" = one Lloyd iteration over all resident frames:
    assign (argmin over 1000 centers) + centroid update + cost, exactly what
    deeptime kmeans.cluster_loop does per iteration (pyemma/coordinates/clustering/kmeans.py:254-258).
`value` = frames assigned per second over the whole job (all GPUs), frames resident in HBM.
`e2e`   = the same iteration with the frames in PINNED HOST memory every step (C-ABI call
          b2k_stage_lloyd_assign_accumulate: H2D chunk by chunk, every chunk assigned and summed while the
          next one is on the bus, labels D2H) + finalize + cost; the cost word is read on the host.
Multi-GPU: frames shard over ranks (weak scaling: 1e7 frames per GPU), one int64 all-reduce of
[k*d sums | k counts] + one cost word per iteration over NCCL.

python bench.py --gpus N --steps K --warmup W [--impl reference] [--frames F]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, K = 10, 1000
FRAMES_PER_GPU = 10_000_000
N_BLOBS = 20
METRIC = "frames assigned/s (Lloyd iteration: assign + centroid update + cost; 1e7x10 fp32 per GPU, k=1000)"
WORKLOAD = "cfg2: TICA-like 1e7x10 fp32 per GPU, k=1000, one Lloyd iteration per step"


def select_workload(name):
    """cfg2 is the driver's contract (BASELINE.json configs[1]); cfg3 = configs[2] (raw pairwise-distance features,
    1e8 x 64 over 8 GPUs = 1.25e7 frames per GPU, k=2000) measures the NCCL-sharded Lloyd iteration the metric names."""
    global D, K, FRAMES_PER_GPU, N_BLOBS, METRIC, WORKLOAD
    if name == "cfg3":
        D, K, FRAMES_PER_GPU, N_BLOBS = 64, 2000, 12_500_000, 50
        METRIC = "frames assigned/s (Lloyd iteration: assign + centroid update + cost; 1.25e7x64 fp32 per GPU, k=2000)"
        WORKLOAD = "cfg3: pairwise-distance-like 1.25e7x64 fp32 per GPU (1e8 over 8 GPUs), k=2000, one Lloyd iteration per step"


def synth_params(seed=2):
    """TICA-like synthetic mixture: 20 metastable blobs, per-dim variance decaying 1,.8,.6..."""
    import numpy as np
    rng = np.random.RandomState(seed)
    scale = np.sqrt(np.maximum(1.0 - 0.2 * np.arange(D), 0.05)).astype(np.float32)
    means = (rng.randn(N_BLOBS, D) * 1.5).astype(np.float32) * scale
    return means, scale * 0.6


def synth_device(n, rank, dev):
    import torch
    means, sig = synth_params()
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    lab = torch.randint(0, N_BLOBS, (n,), generator=g, device=dev)
    X = torch.randn((n, D), generator=g, device=dev, dtype=torch.float32)
    X.mul_(torch.from_numpy(sig).to(dev)).add_(torch.from_numpy(means).to(dev)[lab])
    return X.contiguous()


def synth_host(n, seed):
    import numpy as np
    means, sig = synth_params()
    rng = np.random.RandomState(seed)
    lab = rng.randint(0, N_BLOBS, n)
    return (means[lab] + sig * rng.standard_normal((n, D)).astype(np.float32)).astype(np.float32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_oracle_rate(n_sample, steps, warmup, threads):
    """frames/s of the oracle's Lloyd iteration (labels + update + cost) on the host cores."""
    import numpy as np
    from oracle import oracle as O
    X = synth_host(n_sample, 99)
    C0 = X[np.random.RandomState(5).choice(n_sample, K, replace=False)].copy()
    times = []
    cen = C0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        newc, labels = O.kmeans_cluster(X, cen, n_threads=threads)
        O.cost(X, newc, labels, n_threads=threads)
        dt = time.perf_counter() - t0
        cen = newc
        if s >= warmup:
            times.append(dt)
    return n_sample * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = args.ref_frames
    rate, sec = cpu_oracle_rate(n_sample, args.steps, args.warmup, threads)
    from oracle import oracle as O
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_sample": n_sample, "d": D, "k": K},
        "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "%d-frame sample of the cfg2 workload per step, full k and d; %s"
                                   % (n_sample, O.build_info())},
        "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--frames", type=int, default=FRAMES_PER_GPU, help="frames per GPU")
    ap.add_argument("--ref-frames", type=int, default=2_000_000, help="frames per step of the CPU reference arm")
    ap.add_argument("--cpu-frames", type=int, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="b2k_ctx_set_option before the run (experiments; repeatable)")
    ap.add_argument("--engine", default="auto", choices=["auto", "direct", "screen"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--stage-mb", type=int, default=0, help="pinned staging chunk of the e2e leg in MB (0: library default)")
    args = ap.parse_args()
    if args.workload != "cfg2":
        select_workload(args.workload)
        if args.frames == 10_000_000:
            args.frames = FRAMES_PER_GPU
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

    n = args.frames
    X = synth_device(n, rank, dev)
    # initial centers: the first K frames of rank 0's shard (identical on all ranks)
    cur = X[:K].clone()
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
    costs = []

    def step():
        nonlocal cur, nxt
        _lib.check(lib.b2k_dev_lloyd_assign_accumulate(sess, C.c_void_p(cur.data_ptr()), C.c_void_p(labels.data_ptr()),
                                                       C.c_void_p(acc.data_ptr())))
        if ws > 1:
            dist.all_reduce(acc[:acc_len - 1])
        _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                              C.c_void_p(nxt.data_ptr())))
        _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), C.c_void_p(labels.data_ptr()),
                                          C.c_void_p(acc.data_ptr())))
        if ws > 1:
            dist.all_reduce(acc[acc_len - 1:])
        # the loop's convergence test needs the cost on the host every iteration
        costs.append(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[acc_len - 1].item())))
        cur, nxt = nxt, cur

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ctx.set_option("profile", 1)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = _lib.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if ws > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if sampler else None
    value = n * ws * args.steps / (total_ms * 1e-3)

    # ---- dominant kernel: the tcgen05 screen kernel, timed by the library with CUDA events on the launching
    #      stream around every one of its launches INSIDE the timed region above (option "profile")
    gemm_launches = ctx.get_stat("screen_gemm_launches")
    gemm_ms = ctx.get_stat("screen_gemm_ms_total") / gemm_launches if gemm_launches else None
    ctx.set_option("profile", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = ("of measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" if peaks
                else "of fallback (B200_PROFILING.md ~1400 sustained)")
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "screen_gemm_kernel n=%d d=%d k=%d" % (n, D, K)
        traffic = tr.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    if gemm_ms:
        flops = 2.0 * K * D * n  # algorithmic: SURVEY 8d "2*k*d flop per frame" x frames per launch
        achieved_tf = flops / (gemm_ms * 1e-3) / 1e12
        k_pad = (K + 255) // 256 * 256
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965
        tmem_peak = 64.0 * 148 * sm_mhz * 1e6 / 1e9  # GB/s: tcgen05.ld moves 64 B/clk/SM (measured)
        tmem_gbs = n * k_pad * 4 / (gemm_ms * 1e-3) / 1e9
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf, "traffic": traffic,
                    "kernel": "b2k::screen_gemm_kernel (tcgen05 distance screen, %d launches timed)" % int(gemm_launches),
                    "kernel_ms": gemm_ms, "peak_source": peak_src,
                    "algorithmic_flops_per_launch": flops,
                    "limiter": ("at d<=16 the kernel is bound by reading the fp32 score matrix out of TMEM "
                                "(tcgen05.ld, 64 B/clk/SM), not by the MMA pipe: see tmem_read" if D <= 16 else
                                "3-term fp16 operand split: the MMA pipe issues 3x the algorithmic flops "
                                "(rigorous margin), so frac is capped at 1/3"),
                    "tmem_read": {"achieved_gbs": tmem_gbs, "peak_gbs": tmem_peak, "frac": tmem_gbs / tmem_peak,
                                  "bytes_per_launch": n * k_pad * 4}}
    else:  # exact CUDA-core engine (--engine direct): 3 fp32 ops per pair-dimension, no FMA (reference rounding)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(3):
            _lib.check(lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, D, C.c_void_p(cur.data_ptr()), K,
                                          _lib.EUCLIDEAN, C.c_void_p(labels.data_ptr()), None))
        a1.record(stream)
        torch.cuda.synchronize(dev)
        ms = a0.elapsed_time(a1) / 3
        achieved_tf = 2.0 * K * D * n / (ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved_tf / peak_tf, "traffic": None, "kernel": "b2k::assign_small_kernel (exact fp32)",
                    "kernel_ms": ms, "peak_source": peak_src}

    # ---- e2e: the same Lloyd iteration, but the frames start in PINNED HOST memory every step:
    #      b2k_stage_lloyd_assign_accumulate (H2D chunk by chunk, each chunk assigned and its member sums added while the
    #      next is on the bus, labels D2H) -> all-reduce -> finalize -> cost -> all-reduce -> cost to the host.
    hx = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
    hx.copy_(X)
    hl = torch.empty(n, dtype=torch.int32, pin_memory=True)
    e2e_steps = max(3, min(args.steps, 8))

    def e2e_step():
        nonlocal cur, nxt
        _lib.check(lib.b2k_stage_lloyd_assign_accumulate(sess, C.c_void_p(hx.data_ptr()), C.c_void_p(cur.data_ptr()),
                                                         C.c_void_p(X.data_ptr()), C.c_void_p(labels.data_ptr()),
                                                         C.c_void_p(hl.data_ptr()), C.c_void_p(acc.data_ptr())))
        if ws > 1:
            dist.all_reduce(acc[:acc_len - 1])
        _lib.check(lib.b2k_dev_lloyd_finalize(sess, C.c_void_p(acc.data_ptr()), C.c_void_p(cur.data_ptr()),
                                              C.c_void_p(nxt.data_ptr())))
        _lib.check(lib.b2k_dev_lloyd_cost(sess, C.c_void_p(nxt.data_ptr()), C.c_void_p(labels.data_ptr()),
                                          C.c_void_p(acc.data_ptr())))
        if ws > 1:
            dist.all_reduce(acc[acc_len - 1:])
        costs.append(lib.b2k_dev_lloyd_decode_cost(sess, int(acc[acc_len - 1].item())))
        cur, nxt = nxt, cur

    if args.stage_mb > 0:
        ctx.set_option("stage_bytes", args.stage_mb << 20)
    e2e_step()
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if ws > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": n * ws * e2e_steps / float(dt.item()), "unit": "frames/s",
           "h2d_bytes_per_step": n * D * 4, "d2h_bytes_per_step": n * 4 + 8,
           "api": "b2k_stage_lloyd_assign_accumulate (pinned host frames -> HBM chunk by chunk, each chunk assigned and "
                  "summed while the next is on the bus, labels back) + b2k_dev_lloyd_finalize/cost",
           "steps": e2e_steps, "ms_per_step": float(dt.item()) / e2e_steps * 1e3}

    cpu = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec = cpu_oracle_rate(args.cpu_frames, 2, 1, threads)
        from oracle import oracle as O
        cpu = {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d-frame sample of the same workload (full k, d), 2 timed Lloyd iterations; %s"
                         % (args.cpu_frames, O.build_info())}

    lib.b2k_dev_lloyd_destroy(sess)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": ws, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_gpu": n, "d": D, "k": K, "parallelism": "frames sharded x%d" % ws,
                       "l2": "inputs (%.0f MB per GPU) larger than L2" % (n * D * 4 / 1e6),
                       "engine": args.engine},
            "lloyd_iters_per_s": args.steps / (total_ms * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "final_cost": costs[-1] if costs else None,
        }
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
