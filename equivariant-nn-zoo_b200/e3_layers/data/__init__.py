from .data import Data
from .batch import Batch
from .compute_edge import computeEdgeIndex, computeEdgeVector
