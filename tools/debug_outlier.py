import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemma_b200 import _lib as b2k
from oracle import oracle as O
def blobs(rng, n, d, nb, spread=5.0, sigma=1.0):
    cen = rng.uniform(-spread, spread, size=(nb, d))
    lab = rng.randint(0, nb, size=n)
    return (cen[lab] + sigma * rng.randn(n, d)).astype(np.float32)
ctx = b2k.context()
ctx.set_option("assign_engine", b2k.ENGINE_SCREEN)
rng = np.random.RandomState(4)
X = blobs(rng, 6000, 8, 5)
X[100] *= 1e4
X[200] = 0
C = X[rng.choice(6000, 300, replace=False)].copy()
ref = O.assign(X, C, n_threads=8)
for terms in (1, 3):
    ctx.set_option("screen_terms", terms)
    got = b2k.assign(X, C)
    bad = np.nonzero(got != ref)[0]
    print("terms", terms, "mismatch", len(bad), "cand/frame", ctx.get_stat("screen_cand_chunks") / 6000, "fallback", ctx.get_stat("screen_fallback_frames"))
    for i in bad[:8]:
        dg = np.sqrt(((X[i].astype(np.float64) - C[got[i]]) ** 2).sum()); dr = np.sqrt(((X[i].astype(np.float64) - C[ref[i]]) ** 2).sum())
        print("  frame", i, "got", got[i], "ref", ref[i], "d_got", dg, "d_ref", dr, "chunk got/ref", got[i] // 32, ref[i] // 32)
print("outlier is a center:", bool((np.abs(C).max(axis=1) > 1e3).any()), "absmax X", np.abs(X).max())
# milder outliers
for f in (1e1, 1e2, 1e3):
    rng = np.random.RandomState(4)
    X = blobs(rng, 6000, 8, 5); X[100] *= f
    C = X[rng.choice(6000, 300, replace=False)].copy()
    ref = O.assign(X, C, n_threads=8)
    ctx.set_option("screen_terms", 3)
    got = b2k.assign(X, C)
    print("outlier x%g" % f, "mismatch", int((got != ref).sum()), "cand/frame", ctx.get_stat("screen_cand_chunks") / 6000, "fallback", ctx.get_stat("screen_fallback_frames"))
