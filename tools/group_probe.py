#!/usr/bin/env python
"""A/B of the screen's candidate group size: python tools/group_probe.py  (Lloyd step at cfg2 / cfg3 / cfg4-tenth shapes)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import config_bench as cb  # noqa: E402
from pyemma_b200 import _lib  # noqa: E402

ctx = _lib.context(0)
ctx.set_stream(torch.cuda.current_stream(cb.DEV).cuda_stream)
shapes = [("cfg2", 10_000_000, 10, 1000, 20, 1.5, 0.6, False), ("cfg3", 12_500_000, 64, 2000, 50, 1.0, 0.3, True),
          ("cfg4/10", 2_000_000, 256, 5000, 200, 5.0, 1.0, False)]
for name, n, d, k, nb, spread, sigma, pos in shapes:
    X, _ = cb.device_blobs(n, d, nb, spread, sigma, 2, positive=pos)
    for g, vm in ((8, 1), (8, 0), (4, 0), (2, 0)):
        ctx.set_option("screen_group", g)
        ctx.set_option("verify_mode", vm)
        torch.manual_seed(0)
        r = cb.lloyd_and_assign(ctx, X, k, 5, name)
        print(json.dumps({"shape": name, "group": g, "verify_mode": vm, "lloyd_ms": round(r["lloyd_ms_per_iter"], 3),
                          "gemm_ms": round(r["screen_gemm_ms_per_iter"], 3),
                          "groups_per_frame": round(r["cand_groups_per_frame"], 3),
                          "centers_per_frame": round(r["cand_groups_per_frame"] * g, 3),
                          "fallback": r["fallback_frames"]}), flush=True)
    del X
    torch.cuda.empty_cache()
ctx.set_option("screen_group", 0)
ctx.set_option("verify_mode", 0)
