"""pyemma_b200 -- B200-native k-means / Voronoi-assignment engine behind PyEMMA's clustering API.

Only the hot path named in BASELINE.json is implemented: k-means++ init,
Lloyd iterations, assign / dtrajs, cluster_regspace and metric='minRMSD'.  The arithmetic lives in
hand-written sm_100a CUDA kernels behind the C ABI of include/b2k.h (libb2k.so, loaded with
ctypes); there is no CPU fallback.
"""
__version__ = "0.1.0"

from .api import (assign_to_centers, cluster_kmeans, cluster_mini_batch_kmeans, cluster_regspace,  # noqa: E402,F401
                  cluster_uniform_time)
from .clustering import (AssignCenters, KmeansClustering, MiniBatchKmeansClustering,  # noqa: E402,F401
                         RegularSpaceClustering, UniformTimeClustering)
