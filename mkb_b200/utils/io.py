"""File readers with the reference's names (mkb/utils/read_csv.py:8-33, read_json.py:6-8): integer ``h,r,t`` rows,
``h,r,t,label`` rows of the triplet-classification files, and the entity / relation label maps."""
import csv
import json

__all__ = ["read_csv", "read_csv_classification", "read_json"]


def read_csv(file_path):
    with open(file_path, newline="") as f:
        return [(int(h), int(r), int(t)) for h, r, t in csv.reader(f)]


def read_csv_classification(path):
    X, y = [], []
    with open(path, newline="") as f:
        for row in csv.reader(f):
            h, r, t, label = (int(v) for v in row[:4])
            X.append([h, r, t])
            y.append(label)
    return {"X": X, "y": y}


def read_json(file_path):
    with open(file_path) as f:
        return json.load(f)
