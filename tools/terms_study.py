"""CPU study (numpy): how many centers per frame survive a screen margin of eps*|x~|*max|c~| on blob data shaped like the
configs -- eps = 2^-21 (3-term fp16 split), 1.01*2^-10 (hi-only operands, the margin Margin::init uses for terms=1) and a
hypothetical per-element bound.  Decides whether `screen_terms=1` (3x fewer MMA flops) is worth measuring on wide rows."""
import numpy as np
rng=np.random.RandomState(4)
def study(n,d,k,nb,spread,sigma,label):
    cen_b = rng.uniform(-spread, spread, size=(nb,d))
    lab = rng.randint(0,nb,size=n+k)
    X = (cen_b[lab] + sigma*rng.randn(n+k,d)).astype(np.float32)
    C = X[n:]; X = X[:n]
    mu = C.mean(0); Xc = X-mu; Cc = C-mu
    d2 = (Xc**2).sum(1)[:,None] - 2*Xc@Cc.T + (Cc**2).sum(1)[None,:]
    xn = np.sqrt((Xc**2).sum(1)); cmax = np.sqrt((Cc**2).sum(1)).max()
    dmin = d2.min(1)
    out=[]
    for name,eps in (("3-term (2^-21)",2.0**-21),("1-term (2^-10, the library margin)",1.01*2.0**-10),("1-term, tight per-element bound /sqrt(d)",2.0**-11/np.sqrt(d)*3)):
        delta = eps*xn*cmax           # score error bound
        cand = (d2 <= (dmin+4*delta)[:,None]).sum(1)
        # groups of 2 centers (default for d>16): count distinct pairs
        grp = np.zeros((n,(k+1)//2),bool)
        m = d2 <= (dmin+4*delta)[:,None]
        grp = m[:, :k - k%2].reshape(n,-1,2).any(2)
        out.append((name, cand.mean(), np.percentile(cand,99), grp.sum(1).mean()))
    print(label, "n=%d d=%d k=%d"%(n,d,k))
    for o in out: print("   %-45s mean candidates %.1f  p99 %.0f  groups-of-2 %.1f"%o)
study(1500,256,5000,200,10.0,1.0,"cfg4-like")
study(1500,64,2000,50,1.0,0.3,"cfg3-like")
study(1500,128,2000,100,2.0,0.5,"d=128")
study(1500,10,1000,20,1.0,0.6,"cfg2-like (overlapping blobs)")
