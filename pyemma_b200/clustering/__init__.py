from .assign import AssignCenters
from .interface import AbstractClustering, NotConvergedWarning
from .kmeans import KmeansClustering, MiniBatchKmeansClustering
from .regspace import RegularSpaceClustering

__all__ = ["AbstractClustering", "AssignCenters", "KmeansClustering", "MiniBatchKmeansClustering",
           "RegularSpaceClustering", "NotConvergedWarning"]
