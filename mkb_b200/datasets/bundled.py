"""``Wn18rr``, ``Fb15k237``, ``Yago310`` — the three bundled datasets BASELINE.json's configs name
(mkb/datasets/wn18rr.py:62-82, fb15k237.py, yago310.py).

The reference ships each as a directory of ``train/valid/test.csv`` (integer ``h,r,t`` rows),
``entities.json``, ``relations.json`` and ``classification_{valid,test}.csv`` next to its Python
wrapper.  This package does not redistribute those files: the wrappers below read them from where they
already are — ``path=`` if given, else ``$MKB_DATASETS/<name>``, else the ``datasets/<name>`` directory
of an installed ``mkb`` — and build the tensorised ``Dataset`` of this package with the reference's
constructor arguments.
"""
from __future__ import annotations

import importlib.util
import json
import os

from ..utils.io import read_csv_classification
from .dataset import Dataset, _read_csv

__all__ = ["Wn18rr", "Fb15k237", "Yago310"]


def _locate(name, path=None):
    tried = []
    if path is not None:
        cands = [path]
    else:
        cands = []
        if os.environ.get("MKB_DATASETS"):
            cands.append(os.path.join(os.environ["MKB_DATASETS"], name))
        try:
            spec = importlib.util.find_spec("mkb")
        except (ImportError, ValueError):
            spec = None
        if spec is not None and spec.submodule_search_locations:
            cands += [os.path.join(p, "datasets", name) for p in spec.submodule_search_locations]
    for c in cands:
        tried.append(c)
        if os.path.exists(os.path.join(c, "train.csv")):
            return c
    raise FileNotFoundError(
        f"{name}: train.csv not found (looked in {tried or 'nowhere'}); pass path=, set MKB_DATASETS to the "
        f"directory holding '{name}/', or install mkb, whose package data contains it")


def _read_classification(path):
    return read_csv_classification(path) if os.path.exists(path) else None


class _Bundled(Dataset):
    filename = ""

    def __init__(self, batch_size, classification=False, shuffle=True, pre_compute=True, num_workers=1, seed=None,
                 path=None, device=None, pin_memory=False):
        root = _locate(self.filename, path)
        with open(os.path.join(root, "entities.json")) as f:
            entities = json.load(f)
        with open(os.path.join(root, "relations.json")) as f:
            relations = json.load(f)
        super().__init__(
            train=_read_csv(os.path.join(root, "train.csv")), valid=_read_csv(os.path.join(root, "valid.csv")),
            test=_read_csv(os.path.join(root, "test.csv")), entities=entities, relations=relations,
            batch_size=batch_size, shuffle=shuffle, classification=classification, pre_compute=pre_compute,
            num_workers=num_workers, seed=seed,
            classification_valid=_read_classification(os.path.join(root, "classification_valid.csv")),
            classification_test=_read_classification(os.path.join(root, "classification_test.csv")),
            device=device, pin_memory=pin_memory)


class Wn18rr(_Bundled):
    """40 943 entities, 11 relations, 86 835 / 3 034 / 3 134 triples (mkb/datasets/wn18rr.py:44-49)."""
    filename = "wn18rr"


class Fb15k237(_Bundled):
    """14 541 entities, 237 relations, 272 115 / 17 535 / 20 466 triples (mkb/datasets/fb15k237.py:44-49)."""
    filename = "fb15k237"


class Yago310(_Bundled):
    """123 182 entities, 37 relations, 1 079 040 / 5 000 / 5 000 triples (mkb/datasets/yago310.py:44-49)."""
    filename = "yago310"
