from .assign import AssignCenters
from .interface import AbstractClustering, NotConvergedWarning
from .kmeans import KmeansClustering, MiniBatchKmeansClustering
from .regspace import RegularSpaceClustering
from .uniform_time import UniformTimeClustering

__all__ = ["AbstractClustering", "AssignCenters", "KmeansClustering", "MiniBatchKmeansClustering",
           "RegularSpaceClustering", "UniformTimeClustering", "NotConvergedWarning"]
