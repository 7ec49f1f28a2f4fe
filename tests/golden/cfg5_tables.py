"""Deterministic, structured embedding tables for the config-5-at-size ranking fixture (test infrastructure).

The fixture pins the REFERENCE's ranks at N = 40 943, D = 1000 (BASELINE.json configs[4]); a trained
table of that size (164-327 MB) cannot be committed, so both the generator (tests/golden/make_golden.py
cfg5, run where /root/reference exists) and the GPU test rebuild the same tables from a seed:

  1. U(-range, range) init, the reference's distribution (mkb/models/base.py:86-100), numpy RandomState;
  2. two sweeps of neighbour averaging over the TRAIN triples (every entity moves half-way to the mean of
     the entities it is linked to), which is what a few epochs of training do to first order: linked
     entities end up close, so true triples score well above random ones, the filtered candidates
     (other true heads / tails) are exactly the ones crowding the top of the list, and ranks spread
     from 1 to N instead of being uniform noise;
  3. relation rows are pulled towards each model's "links score high" direction — a small translation
     (TransE), a rotation by a small phase (RotatE), a positive diagonal (DistMult), a positive real part
     with a small imaginary part (ComplEx) — which is where training takes the mostly symmetric Wn18rr
     relations.

Only IEEE-exact fp32 operations (add, multiply, divide; ``np.add.at`` accumulates in index order) are
used, so the tables are bit-identical on every machine.
"""
import numpy as np

EMBEDDING_EPS = 2.0  # mkb/models/base.py:74


def table_dims(model, D):
    ent = 2 * D if model in ("RotatE", "ComplEx") else D
    rel = 2 * D if model == "ComplEx" else D
    return ent, rel


def make_tables(model, train, n_entity, n_relation, D, gamma, seed=2024, sweeps=2):
    rng = np.random.RandomState(seed)
    r = np.float32((gamma + EMBEDDING_EPS) / D)
    de, dr = table_dims(model, D)
    ent = rng.uniform(-r, r, size=(n_entity, de)).astype(np.float32)
    rel = rng.uniform(-r, r, size=(n_relation, dr)).astype(np.float32)
    train = np.asarray(train, dtype=np.int64)
    h, t = train[:, 0], train[:, 2]
    deg = np.zeros(n_entity, dtype=np.float32)
    np.add.at(deg, h, np.float32(1))
    np.add.at(deg, t, np.float32(1))
    deg = np.maximum(deg, np.float32(1))[:, None]
    for _ in range(sweeps):
        acc = np.zeros_like(ent)
        np.add.at(acc, t, ent[h])
        np.add.at(acc, h, ent[t])
        ent = (np.float32(0.5) * ent + np.float32(0.5) * (acc / deg)).astype(np.float32)
    # keep the init's scale (averaging shrinks the rows) so that scores stay in the range gamma expects
    ent = (ent * np.float32(2.0)).astype(np.float32)
    small = np.float32(0.05)
    if model in ("TransE", "RotatE"):
        rel = rel * small
    elif model == "DistMult":
        rel = np.abs(rel)
    elif model == "ComplEx":
        rel = np.concatenate([np.abs(rel[:, :D]), rel[:, D:] * small], axis=1)
    return np.ascontiguousarray(ent), np.ascontiguousarray(rel.astype(np.float32))
