from .data import Data
from .batch import Batch
from .compute_edge import computeEdgeIndex, computeEdgeVector
from .dataset import CondensedDataset
from .dataloader import Collater, DataLoader, DevicePipeline, getDataIters
