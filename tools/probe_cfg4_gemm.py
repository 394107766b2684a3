import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import config_bench as cb
from pyemma_b200 import _lib
ctx = _lib.context(0); ctx.set_stream(torch.cuda.current_stream(cb.DEV).cuda_stream)
for (n, d, k, nb, spread, sigma) in ((2_000_000, 256, 5000, 200, 5.0, 1.0), (4_000_000, 64, 2000, 50, 1.0, 0.3)):
    X, _ = cb.device_blobs(n, d, nb, spread, sigma, 4)
    cen = X[torch.randperm(n, device=cb.DEV)[:k]].clone(); lab = torch.empty(n, dtype=torch.int32, device=cb.DEV)
    for _ in range(2):
        _lib.check(ctx.lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, d, C.c_void_p(cen.data_ptr()), k, 0, C.c_void_p(lab.data_ptr()), None))
    torch.cuda.synchronize()
    del X
