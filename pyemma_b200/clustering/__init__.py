from .assign import AssignCenters
from .interface import AbstractClustering, NotConvergedWarning
from .kmeans import KmeansClustering
from .regspace import RegularSpaceClustering

__all__ = ["AbstractClustering", "AssignCenters", "KmeansClustering", "RegularSpaceClustering",
           "NotConvergedWarning"]
