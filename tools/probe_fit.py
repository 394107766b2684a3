"""cluster_kmeans on a pageable host array at cfg2 size: where the wall time goes (gather / Lloyd / dtrajs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import pyemma_b200 as coor
X = bench.synth_host(10_000_000, 3)
C0 = X[:1000].copy()
coor.cluster_kmeans(X[:200_000], k=1000, max_iter=2, clustercenters=C0)   # warm-up
for rep in range(2):
    t0 = time.perf_counter()
    km = coor.cluster_kmeans(X, k=1000, max_iter=10, tolerance=0.0, clustercenters=C0)
    t1 = time.perf_counter()
    dt = km.dtrajs
    t2 = time.perf_counter()
    print("fit (gather + 10 Lloyd iterations) %.1f ms, dtrajs %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
