"""TEST INFRASTRUCTURE: numpy-level calls into libkge_emu.so (the kernels of mkb_b200/csrc compiled for
the host against tests/emu/include/cuda_runtime.h).  Same C ABI, same ctypes prototypes as the product
binding (mkb_b200/_native.py); "device pointers" are numpy buffers."""
import ctypes as C
import os

import numpy as np

from mkb_b200 import _native as N

from . import build_emu

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(os.environ.get("KGE_EMU_LIB") or build_emu.build())  # KGE_EMU_LIB: e.g. an ASAN build
        for name, (res, args) in N.PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        l.kge_emu_launch_count.restype = C.c_long
        _lib = l
    return _lib


def P(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "emulated device buffers must be contiguous"
    return a.ctypes.data


def ok(rc, what=""):
    assert rc == 0, f"{what} returned {rc}: {lib().kge_strerror(rc).decode()}"


NC = {"TransE": 1, "DistMult": 1, "ComplEx": 2, "RotatE": 2, "pRotatE": 1}
RC = {"TransE": 1, "DistMult": 1, "ComplEx": 2, "RotatE": 1, "pRotatE": 1}


def mode_id(mode):
    return N.HEAD_BATCH if mode == "head-batch" else N.TAIL_BATCH


def tables(model, ent, rel, gamma, n_entity=None, modulus=None):
    D = rel.shape[1] // RC[model]
    rng = np.float32((np.float32(gamma) + np.float32(2)) / np.float32(D))  # (gamma + 2) / D as the model stores it
    tb = N.KgeTables(P(ent), P(rel), ent.shape[0] if n_entity is None else n_entity, rel.shape[0], D,
                     N.MODEL_IDS[model], float(gamma), float(rng), None)
    if modulus is not None:  # pRotatE: a "device" scalar
        tb._mod = np.array([modulus], dtype=np.float32)
        tb.modulus = P(tb._mod)
    return tb


def csr_struct(csr):
    k, o, m = (np.ascontiguousarray(x, dtype=np.int64) for x in csr)
    st = N.KgeFilterCsr(P(k), P(o), P(m), k.shape[0])
    st._keep = (k, o, m)
    return st


def shards_struct(ent_shards, grad_shards=None, scalar_red=False):
    st = N.KgeShards()
    for s, t in enumerate(ent_shards):
        st.entity[s] = P(t)
        st.grad_entity[s] = P(grad_shards[s]) if grad_shards is not None else None
    st.n_shards = len(ent_shards)
    st.scalar_red = int(scalar_red)
    st._keep = (ent_shards, grad_shards)
    return st


def split_rows(table, G):
    rows = -(-table.shape[0] // G)
    out = []
    for s in range(G):
        sh = np.zeros((rows, table.shape[1]), dtype=table.dtype)
        part = table[s::G]
        sh[: part.shape[0]] = part
        out.append(sh)
    return out


def merge_rows(shards, n):
    G = len(shards)
    out = np.empty((n, shards[0].shape[1]), dtype=shards[0].dtype)
    for s, sh in enumerate(shards):
        out[s::G] = sh[: (n - s + G - 1) // G]
    return out


def score(model, ent, rel, gamma, sample, neg=None, mode=None, shards=None, modulus=None):
    l = lib()
    B = sample.shape[0]
    K = 1 if neg is None else neg.shape[1]
    out = np.full((B, K), np.nan, dtype=np.float32)
    tb = tables(model, ent, rel, gamma, modulus=modulus)
    if shards is None:
        ok(l.kge_score_fwd(C.byref(tb), mode_id(mode), P(sample), B, P(neg), 0 if neg is None else K, P(out), None))
    else:
        tb.entity = None
        ok(l.kge_score_fwd_sharded(C.byref(tb), C.byref(shards), mode_id(mode), P(sample), B, P(neg),
                                   0 if neg is None else K, P(out), None))
    return out


def fused_fwd(model, ent, rel, gamma, sample, neg, w, mode, alpha=0.5, shards=None, modulus=None):
    l = lib()
    B, K = neg.shape
    r = dict(pos=np.full((B, 1), np.nan, np.float32), neg=np.full((B, K), np.nan, np.float32),
             cpos=np.full(B, np.nan, np.float32), cneg=np.full((B, K), np.nan, np.float32),
             stats=np.zeros(4, np.float32))
    ws = np.zeros(l.kge_loss_workspace_bytes(B) + 64, dtype=np.uint8)
    tb = tables(model, ent, rel, gamma, modulus=modulus)
    if shards is None:
        ok(l.kge_fused_fwd(C.byref(tb), mode_id(mode), P(sample), B, P(neg), K, P(w), alpha, P(r["pos"]),
                           P(r["neg"]), P(r["cpos"]), P(r["cneg"]), P(r["stats"]), P(ws), None), "kge_fused_fwd")
    else:
        tb.entity = None
        ok(l.kge_fused_fwd_sharded(C.byref(tb), C.byref(shards), mode_id(mode), P(sample), B, P(neg), K, P(w), alpha,
                                   P(r["pos"]), P(r["neg"]), P(r["cpos"]), P(r["cneg"]), P(r["stats"]), P(ws), None),
           "kge_fused_fwd_sharded")
    assert not ws[:4].any(), "ticket not reset"
    return r


def modulus_grad(f, gamma, modulus, grad_loss=None):
    """dL/dmodulus of a fused step from its saved scores and coefficients (two kge_modulus_grad calls)."""
    l = lib()
    out = np.zeros(1, np.float32)
    mod = np.array([modulus], np.float32)
    gl = None if grad_loss is None else np.array([grad_loss], np.float32)
    for sc, co in ((f["pos"], f["cpos"]), (f["neg"], f["cneg"])):
        ok(l.kge_modulus_grad(P(sc), P(co), sc.size, P(f["stats"]), P(gl), float(gamma), P(mod), P(out), None))
    return float(out[0])


def fused_bwd(model, ent, rel, gamma, sample, neg, mode, f, shards=None, grad_loss=None, modulus=None):
    l = lib()
    B, K = neg.shape
    g_rel = np.zeros_like(rel)
    tb = tables(model, ent, rel, gamma, modulus=modulus)
    gl = None if grad_loss is None else np.array([grad_loss], np.float32)
    if shards is None:
        g_ent = np.zeros_like(ent)
        ok(l.kge_fused_bwd(C.byref(tb), mode_id(mode), P(sample), B, P(neg), K, P(f["cpos"]), P(f["cneg"]),
                           P(f["stats"]), P(gl), P(g_ent), P(g_rel), None), "kge_fused_bwd")
        return g_ent, g_rel
    tb.entity = None
    ok(l.kge_fused_bwd_sharded(C.byref(tb), C.byref(shards), mode_id(mode), P(sample), B, P(neg), K, P(f["cpos"]),
                               P(f["cneg"]), P(f["stats"]), P(gl), P(g_rel), None), "kge_fused_bwd_sharded")
    return None, g_rel
