"""Negative sampling on the device (kge_sample_negatives / kge_filter_pool); the CSR form of the
reference's true-head / true-tail dictionaries lives in mkb_b200.utils.filters."""
from .negative_sampling import NegativeSampling, positive_triples

__all__ = ["NegativeSampling", "positive_triples"]
