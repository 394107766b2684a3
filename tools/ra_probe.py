import sys, torch
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import config_bench as cb
from pyemma_b200 import _lib
ctx=_lib.context(0); ctx.set_stream(torch.cuda.current_stream(cb.DEV).cuda_stream)
X,_=cb.device_blobs(12_500_000,64,50,1.0,0.3,3,positive=True)
for ra in (0,1):
    ctx.set_option("screen_resident_a", ra)
    torch.manual_seed(0)
    r=cb.lloyd_and_assign(ctx,X,2000,4,"cfg3")
    print("resident_a=%d lloyd %.2f ms gemm %.2f ms groups/frame %.3f fallback %d"%(ra,r["lloyd_ms_per_iter"],r["screen_gemm_ms_per_iter"],r["cand_groups_per_frame"],r["fallback_frames"]),flush=True)
