"""Link-prediction evaluation on the ranking kernels, triplet classification on the positives kernel."""
from .classif import accuracy, find_threshold
from .evaluation import Evaluation

__all__ = ["Evaluation", "accuracy", "find_threshold"]
