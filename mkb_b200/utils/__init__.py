from .bar import Bar
from .filters import build_filter_csr, triples_to_array

__all__ = ["Bar", "build_filter_csr", "triples_to_array"]
