"""True-triple sets as CSR arrays (the device-side form of positive_triples,
mkb/sampling/negative_sampling.py:7-28, and of TestDataset's true_triples set,
mkb/datasets/base.py:185-190).  Vectorised numpy: a 1 M-triple graph builds in well under a second."""
import numpy as np

__all__ = ["build_filter_csr", "triples_to_array"]


def triples_to_array(triples):
    arr = np.asarray(triples, dtype=np.int64)
    if arr.size == 0:
        return arr.reshape(0, 3)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError("triples must be a list of (head, relation, tail) ids")
    return arr


def build_filter_csr(triples, n_entity, side):
    """``side='head'``: key (r,t) -> sorted unique heads; ``side='tail'``: key (h,r) -> sorted unique
    tails.  Key code = relation * n_entity + fixed entity.  Returns int64 (keys, offsets, members)."""
    arr = triples_to_array(triples)
    if side == "head":
        code, member = arr[:, 1] * n_entity + arr[:, 2], arr[:, 0]
    elif side == "tail":
        code, member = arr[:, 1] * n_entity + arr[:, 0], arr[:, 2]
    else:
        raise ValueError(side)
    pairs = np.unique(np.stack([code, member], axis=1), axis=0)  # sorted by (code, member), deduplicated
    keys, start = np.unique(pairs[:, 0], return_index=True)
    offsets = np.concatenate([start, [pairs.shape[0]]]).astype(np.int64)
    return keys.astype(np.int64), offsets, np.ascontiguousarray(pairs[:, 1]).astype(np.int64)
