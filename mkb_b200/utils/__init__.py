from .bar import Bar, BarRange
from .filters import build_filter_csr, triples_to_array
from .io import read_csv, read_csv_classification, read_json
from .predict import FetchToPredict, make_prediction
from .top_k import TopK

__all__ = ["Bar", "BarRange", "build_filter_csr", "triples_to_array", "FetchToPredict", "make_prediction", "TopK",
           "read_csv", "read_csv_classification", "read_json"]
