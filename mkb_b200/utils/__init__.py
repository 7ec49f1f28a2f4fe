from .bar import Bar
from .filters import build_filter_csr, triples_to_array
from .predict import FetchToPredict, make_prediction
from .top_k import TopK

__all__ = ["Bar", "build_filter_csr", "triples_to_array", "FetchToPredict", "make_prediction", "TopK"]
