"""screen kernel as 2-CTA clusters -- multicast center tiles (option screen_cluster=2) or CTA-pair MMAs (cta_group::2,
screen_cluster=3) -- vs plain streaming mode: labels must be identical; prints the Lloyd step / screen kernel times."""
import os, sys, torch, ctypes as C
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import config_bench as cb
from pyemma_b200 import _lib
ctx = _lib.context(0); ctx.set_stream(torch.cuda.current_stream(cb.DEV).cuda_stream)
shapes = [(4357, 64, 2000, 50, 1.0, 0.3), (300_001, 256, 1000, 200, 5.0, 1.0), (4_000_000, 64, 2000, 50, 1.0, 0.3), (2_000_000, 256, 5000, 200, 5.0, 1.0)]
if len(sys.argv) > 1 and sys.argv[1] == "wide":   # where does the pair mode start to pay?
    shapes = [(2_000_000, 96, 2000, 100, 2.0, 0.5), (2_000_000, 128, 2000, 100, 2.0, 0.5), (2_000_000, 192, 3000, 100, 3.0, 0.7),
              (1_000_000, 512, 2000, 100, 5.0, 1.0), (500_000, 900, 1000, 30, 2.0, 0.05)]
elif len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (n, d, k, nb, spread, sigma) in shapes:
    X, _ = cb.device_blobs(n, d, nb, spread, sigma, 4)
    cen = X[torch.randperm(n, device=cb.DEV)[:k]].clone()
    labs = {}
    for mode in (0, 2, 3):
        ctx.set_option("screen_cluster", mode)
        lab = torch.empty(n, dtype=torch.int32, device=cb.DEV)
        _lib.check(ctx.lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, d, C.c_void_p(cen.data_ptr()), k, 0, C.c_void_p(lab.data_ptr()), None))
        torch.cuda.synchronize()
        labs[mode] = lab
        torch.manual_seed(0)
        r = cb.lloyd_and_assign(ctx, X, k, 3, "probe")
        print("n=%d d=%d k=%d cluster=%d: lloyd %.2f ms, screen kernel %.2f ms" % (n, d, k, mode, r["lloyd_ms_per_iter"], r["screen_gemm_ms_per_iter"]), flush=True)
    print("   labels identical: multicast", bool((labs[0] == labs[2]).all()), "pair", bool((labs[0] == labs[3]).all()), flush=True)
    del X
ctx.set_option("screen_cluster", 0)
