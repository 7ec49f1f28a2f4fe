from .base import TestDataset, TestDatasetRelation
from .bundled import Fb15k237, Wn18rr, Yago310
from .dataset import Dataset, from_directory

__all__ = ["Dataset", "from_directory", "TestDataset", "TestDatasetRelation", "Wn18rr", "Fb15k237", "Yago310"]
