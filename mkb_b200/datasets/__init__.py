from .base import TestDataset, TestDatasetRelation
from .dataset import Dataset, from_directory

__all__ = ["Dataset", "from_directory", "TestDataset", "TestDatasetRelation"]
