import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import config_bench as cb
from pyemma_b200 import _lib
ctx = _lib.context(0); ctx.set_stream(torch.cuda.current_stream(cb.DEV).cuda_stream)
n, k = 200_000, 1000
X = cb.conformations_device(n, 300, 30, 5)
cen = X[:k].clone(); lab = torch.empty(n, dtype=torch.int32, device=cb.DEV)
for _ in range(2):
    _lib.check(ctx.lib.b2k_dev_assign(ctx.handle, C.c_void_p(X.data_ptr()), n, 900, C.c_void_p(cen.data_ptr()), k, 1, C.c_void_p(lab.data_ptr()), None))
torch.cuda.synchronize()
