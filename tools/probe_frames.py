import sys, ctypes as C, torch
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import config_bench as cb
from pyemma_b200 import _lib
ctx = _lib.context(0); ctx.set_stream(torch.cuda.current_stream(cb.DEV).cuda_stream)
for (n,d,k) in ((4_000_000, 64, 2000), (10_000_000, 10, 1000)):
    X,_ = cb.device_blobs(n, d, 20, 1.0, 0.3, 3)
    cen = X[:k].clone(); lab = torch.empty(n, dtype=torch.int32, device=cb.DEV)
    for _ in range(2):
        _lib.check(ctx.lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, d, C.c_void_p(cen.data_ptr()), k, 0, C.c_void_p(lab.data_ptr()), None))
    torch.cuda.synchronize()
