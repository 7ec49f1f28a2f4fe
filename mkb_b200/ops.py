"""Thin tensor-level wrappers over the C ABI and the autograd glue around them.

Nothing here computes: every function validates devices/dtypes, allocates outputs with torch and
launches one kernel of libkge_b200.so on the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

__all__ = [
    "score", "adversarial_loss", "fused_adversarial_step", "sample_negatives", "filter_pool",
    "rank_all", "adam_step", "FilterCSR", "TableSpec", "kl_divergence", "topk_rows", "ShardSet", "shard_rows",
    "split_rows", "merge_rows", "score_sharded",
]


class TableSpec:
    """The constants of a model that the kernels need, detached from nn.Module."""

    def __init__(self, model_name, hidden_dim, gamma, embedding_range):
        self.model_name = model_name
        self.model_id = N.MODEL_IDS[model_name]
        self.hidden_dim = int(hidden_dim)
        self.gamma = float(gamma)
        self.embedding_range = float(embedding_range)

    def _modulus_ptr(self, modulus):
        """pRotatE reads its trainable modulus on the device (kge_tables_t.modulus)."""
        if self.model_name != "pRotatE":
            return None
        if modulus is None:
            raise ValueError("pRotatE needs its modulus tensor")
        N.require_cuda(modulus)
        if modulus.dtype != torch.float32 or modulus.numel() != 1:
            raise TypeError("modulus must be a float32 tensor with one element")
        return modulus.data_ptr()

    def struct(self, ent, rel, modulus=None):
        nc = 2 if self.model_name in ("ComplEx", "RotatE") else 1
        rc = 2 if self.model_name == "ComplEx" else 1
        if ent.shape[1] != nc * self.hidden_dim or rel.shape[1] != rc * self.hidden_dim:
            raise ValueError(
                f"{self.model_name}: table shapes {tuple(ent.shape)}/{tuple(rel.shape)} do not match "
                f"hidden_dim={self.hidden_dim}")
        return N.KgeTables(ent.data_ptr(), rel.data_ptr(), ent.shape[0], rel.shape[0], self.hidden_dim,
                           self.model_id, self.gamma, self.embedding_range, self._modulus_ptr(modulus))

    def struct_sharded(self, n_entity, rel, modulus=None):
        """Tables struct of a row-sharded model: no local entity table, the GLOBAL entity count."""
        rc = 2 if self.model_name == "ComplEx" else 1
        if rel.shape[1] != rc * self.hidden_dim:
            raise ValueError(f"{self.model_name}: relation table {tuple(rel.shape)} does not match "
                             f"hidden_dim={self.hidden_dim}")
        return N.KgeTables(None, rel.data_ptr(), int(n_entity), rel.shape[0], self.hidden_dim, self.model_id,
                           self.gamma, self.embedding_range, self._modulus_ptr(modulus))

    @property
    def entity_dim(self):
        return self.hidden_dim * (2 if self.model_name in ("ComplEx", "RotatE") else 1)


def _mode_id(mode):
    if mode == "head-batch":
        return N.HEAD_BATCH
    if mode in ("tail-batch", None):
        return N.TAIL_BATCH
    raise ValueError(f"unknown mode {mode!r}")


def _prep_tables(ent, rel):
    N.require_cuda(ent, rel)
    if ent.dtype != torch.float32 or rel.dtype != torch.float32:
        raise TypeError("embedding tables must be float32")
    return ent.contiguous(), rel.contiguous()


def _prep_ids(t, device):
    if t is None:
        return None
    N.require_cuda(t)
    return t.to(device=device, dtype=torch.int64).contiguous()


# ---------------------------------------------------------------------------------------------
# K1 score (+ autograd)
# ---------------------------------------------------------------------------------------------
def _modulus_grad(spec, scores, grad_scores, modulus, out, stats=None, grad_loss=None):
    """out[0] += d/dmodulus of sum(grad_scores * scores) (pRotatE; kge_modulus_grad)."""
    lib = N.load()
    N.check(lib.kge_modulus_grad(N.ptr(scores), N.ptr(grad_scores), scores.numel(), N.ptr(stats), N.ptr(grad_loss),
                                 spec.gamma, N.ptr(modulus), N.ptr(out), N.stream_ptr(scores.device)),
            "kge_modulus_grad")
    N.count_launch()


def _score_fwd(spec, ent, rel, sample, neg, mode, modulus=None):
    lib = N.load()
    B = sample.shape[0]
    K = 1 if neg is None else neg.shape[1]
    out = torch.empty((B, K), dtype=torch.float32, device=ent.device)
    tb = spec.struct(ent, rel, modulus)
    N.check(lib.kge_score_fwd(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(neg),
                              0 if neg is None else K, N.ptr(out), N.stream_ptr(ent.device)),
            "kge_score_fwd")
    N.count_launch()
    return out


class _ScoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ent, rel, sample, neg, mode, spec, modulus):
        ent_c, rel_c = _prep_tables(ent, rel)
        mod_c = modulus.detach() if modulus is not None else None
        with torch.cuda.device(ent.device):
            out = _score_fwd(spec, ent_c, rel_c, sample, neg, mode, mod_c)
        # pRotatE keeps its scores: d score / d modulus = (score - gamma) / modulus
        ctx.save_for_backward(ent_c, rel_c, sample, neg, mod_c, out if mod_c is not None else None)
        ctx.mode, ctx.spec = mode, spec
        return out

    @staticmethod
    def backward(ctx, grad_scores):
        ent, rel, sample, neg, modulus, out = ctx.saved_tensors
        lib = N.load()
        g = grad_scores.contiguous().float()
        # index_select's backward yields dense gradients: same contract here
        g_ent = torch.zeros_like(ent)
        g_rel = torch.zeros_like(rel)
        tb = ctx.spec.struct(ent, rel, modulus)
        B = sample.shape[0]
        g_mod = None
        with torch.cuda.device(ent.device):
            N.check(lib.kge_score_bwd(C.byref(tb), _mode_id(ctx.mode), N.ptr(sample), B, N.ptr(neg),
                                      0 if neg is None else neg.shape[1], N.ptr(g), N.ptr(g_ent),
                                      N.ptr(g_rel), N.stream_ptr(ent.device)), "kge_score_bwd")
            if modulus is not None and ctx.needs_input_grad[6]:
                g_mod = torch.zeros_like(modulus)
                _modulus_grad(ctx.spec, out, g, modulus, g_mod)
        N.count_launch()
        return g_ent, g_rel, None, None, None, None, g_mod


def score(spec, ent, rel, sample, neg=None, mode=None, modulus=None):
    """``model(sample[, negative_sample, mode])`` -> float32 ``[B,1]`` / ``[B,K]``, differentiable
    w.r.t. both tables — and pRotatE's ``modulus`` — (mkb/models/base.py:153-207 + the model's forward)."""
    N.require_cuda(ent, rel, sample, neg)
    sample = _prep_ids(sample, ent.device)
    neg = _prep_ids(neg, ent.device)
    if neg is None:
        mode = None
    if sample.dim() != 2 or sample.shape[1] != 3:
        raise ValueError("sample must be [B,3]")
    if neg is not None and (neg.dim() != 2 or neg.shape[0] != sample.shape[0]):
        raise ValueError("negative_sample must be [B,K]")
    if spec.model_name != "pRotatE":
        modulus = None
    return _ScoreFn.apply(ent, rel, sample, neg, mode, spec, modulus)


# ---------------------------------------------------------------------------------------------
# stand-alone adversarial loss (+ autograd)
# ---------------------------------------------------------------------------------------------
_workspaces = {}


def _loss_workspace(B, device):
    """Zero-initialised ticket + partials buffer, cached per (device, stream, B-bucket)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    need = N.load().kge_loss_workspace_bytes(B)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 16), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


class _AdvLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, neg, weight, alpha):
        lib = N.load()
        pos_c = pos.contiguous().float().view(-1)
        neg_c = neg.contiguous().float()
        w_c = weight.contiguous().float().view(-1)
        B, K = neg_c.shape
        stats = torch.empty(4, dtype=torch.float32, device=neg_c.device)
        with torch.cuda.device(neg_c.device):
            ws = _loss_workspace(B, neg_c.device)
            N.check(lib.kge_adv_loss_fwd(N.ptr(pos_c), N.ptr(neg_c), N.ptr(w_c), B, K, alpha, N.ptr(stats),
                                         N.ptr(ws), N.stream_ptr(neg_c.device)), "kge_adv_loss_fwd")
        N.count_launch()
        ctx.save_for_backward(pos_c, neg_c, w_c, stats)
        ctx.alpha, ctx.pos_shape = alpha, pos.shape
        return stats[3].clone()

    @staticmethod
    def backward(ctx, grad_loss):
        pos, neg, w, stats = ctx.saved_tensors
        lib = N.load()
        B, K = neg.shape
        gpos = torch.empty_like(pos)
        gneg = torch.empty_like(neg)
        gl = grad_loss.contiguous().float()
        with torch.cuda.device(neg.device):
            N.check(lib.kge_adv_loss_bwd(N.ptr(pos), N.ptr(neg), N.ptr(w), B, K, ctx.alpha, N.ptr(stats),
                                         N.ptr(gl), N.ptr(gpos), N.ptr(gneg), N.stream_ptr(neg.device)),
                    "kge_adv_loss_bwd")
        N.count_launch()
        return gpos.view(ctx.pos_shape), gneg, None, None


def adversarial_loss(pos, neg, weight, alpha=0.5):
    """losses.Adversarial.__call__ (mkb/losses/adversarial.py:21-30) as one kernel."""
    N.require_cuda(pos, neg, weight)
    return _AdvLossFn.apply(pos, neg, weight, float(alpha))


# ---------------------------------------------------------------------------------------------
# KL divergence between row softmaxes (distillation loss) + autograd
# ---------------------------------------------------------------------------------------------
class _KlDivFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, T):
        lib = N.load()
        s = student.contiguous().float()
        t = teacher.contiguous().float()
        if s.dim() != 2 or s.shape != t.shape:
            raise ValueError("student_score and teacher_score must both be [n, k]")
        B, K = s.shape
        loss = torch.empty(1, dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            ws = _loss_workspace(B, s.device)
            N.check(lib.kge_kl_div_fwd(N.ptr(s), N.ptr(t), B, K, T, N.ptr(loss), N.ptr(ws), N.stream_ptr(s.device)),
                    "kge_kl_div_fwd")
        N.count_launch()
        ctx.save_for_backward(s, t)
        ctx.T = T
        ctx.shapes = (student.shape, teacher.shape)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, grad_loss):
        s, t = ctx.saved_tensors
        lib = N.load()
        B, K = s.shape
        gs = torch.empty_like(s)
        gt = torch.empty_like(t) if ctx.needs_input_grad[1] else None
        gl = grad_loss.contiguous().float()
        with torch.cuda.device(s.device):
            N.check(lib.kge_kl_div_bwd(N.ptr(s), N.ptr(t), B, K, ctx.T, N.ptr(gl), N.ptr(gs), N.ptr(gt),
                                       N.stream_ptr(s.device)), "kge_kl_div_bwd")
        N.count_launch()
        return gs.view(ctx.shapes[0]), (gt.view(ctx.shapes[1]) if gt is not None else None), None


def kl_divergence(student_score, teacher_score, T=1):
    """losses.KlDivergence.__call__ (mkb/losses/kl_divergence.py:22-29) as one kernel each way."""
    N.require_cuda(student_score, teacher_score)
    return _KlDivFn.apply(student_score, teacher_score, float(T))


# ---------------------------------------------------------------------------------------------
# K8: exact row-wise top-k
# ---------------------------------------------------------------------------------------------
def topk_rows(scores, k, return_values=False):
    """Columns of the ``k`` largest entries of every row of ``scores[rows, cols]`` in descending score
    order, ties by ascending column (= ``argsort(descending, stable)[:, :k]``) -> int64 ``[rows, k]``."""
    lib = N.load()
    N.require_cuda(scores)
    if scores.dim() == 1:
        scores = scores.view(1, -1)
    if scores.dim() != 2:
        raise ValueError("scores must be [rows, cols]")
    x = scores.detach()
    if x.dtype != torch.float32 or x.stride(1) != 1:
        x = x.float().contiguous()
    rows, cols = x.shape
    idx = torch.empty((rows, k), dtype=torch.int64, device=x.device)
    val = torch.empty((rows, k), dtype=torch.float32, device=x.device) if return_values else None
    with torch.cuda.device(x.device):
        N.check(lib.kge_topk_rows(N.ptr(x), rows, cols, x.stride(0) if rows > 1 else cols, int(k), N.ptr(idx),
                                  N.ptr(val), N.stream_ptr(x.device)), "kge_topk_rows")
    N.count_launch()
    return (idx, val) if return_values else idx


# ---------------------------------------------------------------------------------------------
# K2 + K3: the fused training step
# ---------------------------------------------------------------------------------------------
class _FusedStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ent, rel, sample, neg, weight, mode, alpha, spec, want_scores, modulus):
        lib = N.load()
        ent_c, rel_c = _prep_tables(ent, rel)
        mod_c = modulus.detach() if modulus is not None else None
        B, K = neg.shape
        dev = ent.device
        coef_pos = torch.empty(B, dtype=torch.float32, device=dev)
        coef_neg = torch.empty((B, K), dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        keep_scores = want_scores or mod_c is not None  # pRotatE's modulus gradient needs them
        pos_s = torch.empty((B, 1), dtype=torch.float32, device=dev) if keep_scores else None
        neg_s = torch.empty((B, K), dtype=torch.float32, device=dev) if keep_scores else None
        tb = spec.struct(ent_c, rel_c, mod_c)
        with torch.cuda.device(dev):
            ws = _loss_workspace(B, dev)
            N.check(lib.kge_fused_fwd(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(neg), K,
                                      N.ptr(weight), alpha, N.ptr(pos_s), N.ptr(neg_s), N.ptr(coef_pos),
                                      N.ptr(coef_neg), N.ptr(stats), N.ptr(ws), N.stream_ptr(dev)),
                    "kge_fused_fwd")
        N.count_launch()
        ctx.save_for_backward(ent_c, rel_c, sample, neg, coef_pos, coef_neg, stats, mod_c,
                              pos_s if mod_c is not None else None, neg_s if mod_c is not None else None)
        ctx.mode, ctx.spec = mode, spec
        loss = stats[3].clone()
        if want_scores:
            ctx.mark_non_differentiable(pos_s, neg_s)
            return loss, pos_s, neg_s
        return loss

    @staticmethod
    def backward(ctx, grad_loss, *_):
        ent, rel, sample, neg, coef_pos, coef_neg, stats, modulus, pos_s, neg_s = ctx.saved_tensors
        lib = N.load()
        g_ent = torch.zeros_like(ent)
        g_rel = torch.zeros_like(rel)
        gl = grad_loss.contiguous().float()
        tb = ctx.spec.struct(ent, rel, modulus)
        B, K = neg.shape
        g_mod = None
        with torch.cuda.device(ent.device):
            N.check(lib.kge_fused_bwd(C.byref(tb), _mode_id(ctx.mode), N.ptr(sample), B, N.ptr(neg), K,
                                      N.ptr(coef_pos), N.ptr(coef_neg), N.ptr(stats), N.ptr(gl),
                                      N.ptr(g_ent), N.ptr(g_rel), N.stream_ptr(ent.device)),
                    "kge_fused_bwd")
            if modulus is not None and ctx.needs_input_grad[9]:
                g_mod = torch.zeros_like(modulus)
                _modulus_grad(ctx.spec, pos_s, coef_pos, modulus, g_mod, stats, gl)
                _modulus_grad(ctx.spec, neg_s, coef_neg, modulus, g_mod, stats, gl)
        N.count_launch()
        return g_ent, g_rel, None, None, None, None, None, None, None, g_mod


def fused_adversarial_step(spec, ent, rel, sample, neg, weight, mode, alpha=0.5, return_scores=False,
                           modulus=None):
    """``loss(model(sample), model(sample, neg, mode), weight)`` (mkb/compose/pipeline.py:211-234) as
    ONE forward kernel; ``.backward()`` on the result launches ONE backward kernel."""
    N.require_cuda(ent, rel, sample, neg, weight)
    if mode not in ("head-batch", "tail-batch"):
        raise ValueError(f"unknown mode {mode!r}")
    sample = _prep_ids(sample, ent.device)
    neg = _prep_ids(neg, ent.device)
    weight = weight.to(device=ent.device, dtype=torch.float32).contiguous().view(-1)
    if neg.dim() != 2 or neg.shape[0] != sample.shape[0] or weight.shape[0] != sample.shape[0]:
        raise ValueError("shape mismatch between sample, negative_sample and weight")
    if spec.model_name != "pRotatE":
        modulus = None
    return _FusedStepFn.apply(ent, rel, sample, neg, weight, mode, float(alpha), spec, bool(return_scores), modulus)


def fused_forward_raw(spec, ent, rel, sample, neg, weight, mode, alpha, coef_pos, coef_neg, stats, ws,
                      pos_score=None, neg_score=None, modulus=None):
    """Autograd-free K2 launch into caller-owned buffers (the device-resident training loop)."""
    lib = N.load()
    tb = spec.struct(ent, rel, modulus)
    B, K = neg.shape
    N.check(lib.kge_fused_fwd(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(neg), K, N.ptr(weight),
                              alpha, N.ptr(pos_score), N.ptr(neg_score), N.ptr(coef_pos), N.ptr(coef_neg),
                              N.ptr(stats), N.ptr(ws), N.stream_ptr(ent.device)), "kge_fused_fwd")
    N.count_launch()


def fused_backward_raw(spec, ent, rel, sample, neg, mode, coef_pos, coef_neg, stats, g_ent, g_rel,
                       grad_loss=None, modulus=None):
    """Autograd-free K3 launch: ADDS into g_ent / g_rel."""
    lib = N.load()
    tb = spec.struct(ent, rel, modulus)
    B, K = neg.shape
    N.check(lib.kge_fused_bwd(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(neg), K, N.ptr(coef_pos),
                              N.ptr(coef_neg), N.ptr(stats), N.ptr(grad_loss), N.ptr(g_ent), N.ptr(g_rel),
                              N.stream_ptr(ent.device)), "kge_fused_bwd")
    N.count_launch()


def byent_workspace(spec, ent, rel, B, K, modulus=None):
    """Scratch buffer for ``bwd_by_entity_adam_raw`` (kge_byent_workspace_bytes)."""
    lib = N.load()
    tb = spec.struct(ent, rel, modulus)
    return torch.empty(max(lib.kge_byent_workspace_bytes(C.byref(tb), B, K), 256), dtype=torch.uint8, device=ent.device)


def bwd_by_entity_adam_raw(spec, ent, rel, sample, neg, mode, coef_pos, coef_neg, stats, exp_avg, exp_avg_sq,
                           rel_exp_avg, rel_exp_avg_sq, step, lr, beta1, beta2, eps, workspace, grad_loss=None,
                           modulus=None):
    """Atomics-free backward with Adam fused in (kge_bwd_by_entity_adam): updates ``ent``, ``rel`` and
    their moments IN PLACE; bit-reproducible."""
    lib = N.load()
    tb = spec.struct(ent, rel, modulus)
    B, K = neg.shape
    N.check(lib.kge_bwd_by_entity_adam(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(neg), K, N.ptr(coef_pos),
                                       N.ptr(coef_neg), N.ptr(stats), N.ptr(grad_loss), N.ptr(ent), N.ptr(exp_avg),
                                       N.ptr(exp_avg_sq), N.ptr(rel), N.ptr(rel_exp_avg), N.ptr(rel_exp_avg_sq),
                                       int(step), lr, beta1, beta2, eps, N.ptr(workspace),
                                       N.stream_ptr(ent.device)), "kge_bwd_by_entity_adam")
    N.count_launch(7)


def fused_backward_chunk_raw(spec, ent, rel, sample, neg, mode, coef_pos, coef_neg, stats, col0, ncols,
                             g_ent_chunk, g_rel_chunk, grad_loss=None, n_records=1, record_stride=0):
    """K3 restricted to hidden-dim columns [col0, col0+ncols): ADDS into the dense chunk buffers.
    ``n_records`` > 1: the tensors are record 0 of a packed multi-record batch (see the header)."""
    lib = N.load()
    tb = spec.struct(ent, rel)
    B, K = neg.shape
    N.check(lib.kge_fused_bwd_chunk(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(neg), K,
                                    N.ptr(coef_pos), N.ptr(coef_neg), N.ptr(stats), N.ptr(grad_loss), col0, ncols,
                                    n_records, record_stride, N.ptr(g_ent_chunk), N.ptr(g_rel_chunk),
                                    N.stream_ptr(ent.device)),
            "kge_fused_bwd_chunk")
    N.count_launch()


def adam_step_chunk(param, grad_chunk, exp_avg, exp_avg_sq, comps, ncols, col0, im_off, step, lr, beta1=0.9,
                    beta2=0.999, eps=1e-8, zero_grad=True):
    lib = N.load()
    N.check(lib.kge_adam_step_chunk(N.ptr(param), N.ptr(grad_chunk), N.ptr(exp_avg), N.ptr(exp_avg_sq),
                                    param.shape[0], comps, ncols, col0, param.shape[1], im_off, int(step), lr,
                                    beta1, beta2, eps, int(bool(zero_grad)), N.stream_ptr(param.device)),
            "kge_adam_step_chunk")
    N.count_launch()


def adam_slice_bcast(replica_ptrs, self_index, grad_slice, m_slice, v_slice, rows, comps, ncols, col0, row_stride,
                     im_off, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, zero_grad=True, device=None):
    """Adam on this rank's column slice, new values stored into every GPU's replica (peer pointers)."""
    lib = N.load()
    arr = (C.c_void_p * len(replica_ptrs))(*[int(p) for p in replica_ptrs])
    N.check(lib.kge_adam_slice_bcast(arr, len(replica_ptrs), self_index, N.ptr(grad_slice), N.ptr(m_slice),
                                     N.ptr(v_slice), rows, comps, ncols, col0, row_stride, im_off, int(step), lr,
                                     beta1, beta2, eps, int(bool(zero_grad)), N.stream_ptr(device)),
            "kge_adam_slice_bcast")
    N.count_launch()


def score_forward_raw(spec, ent, rel, sample, neg, mode, out):
    """K1 into a caller-owned buffer: out[B] = model(sample) when ``neg`` is None, else out[B,K]."""
    lib = N.load()
    tb = spec.struct(ent, rel)
    N.check(lib.kge_score_fwd(C.byref(tb), _mode_id(mode), N.ptr(sample), sample.shape[0], N.ptr(neg),
                              0 if neg is None else neg.shape[1], N.ptr(out), N.stream_ptr(ent.device)), "kge_score_fwd")
    N.count_launch()


def adversarial_loss_raw(pos, neg, weight, alpha, stats, ws, grad_pos, grad_neg):
    """Stand-alone self-adversarial loss, forward + backward, into caller-owned buffers: stats[4] = (S_p, S_n, W,
    loss), grad_pos[B] / grad_neg[B,K] = dL/dscore (already carrying 1/(2W))."""
    lib = N.load()
    B, K = neg.shape
    st = N.stream_ptr(neg.device)
    N.check(lib.kge_adv_loss_fwd(N.ptr(pos), N.ptr(neg), N.ptr(weight), B, K, alpha, N.ptr(stats), N.ptr(ws), st),
            "kge_adv_loss_fwd")
    N.check(lib.kge_adv_loss_bwd(N.ptr(pos), N.ptr(neg), N.ptr(weight), B, K, alpha, N.ptr(stats), None,
                                 N.ptr(grad_pos), N.ptr(grad_neg), st), "kge_adv_loss_bwd")
    N.count_launch(2)


def peer_copy(src, dst_ptrs, self_index, dst_offset_bytes, nbytes=None):
    """``src`` (contiguous device tensor) -> every peer's buffer at ``dst_offset_bytes`` (NVLink stores)."""
    lib = N.load()
    nbytes = src.numel() * src.element_size() if nbytes is None else int(nbytes)
    arr = (C.c_void_p * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
    N.check(lib.kge_peer_copy(N.ptr(src), arr, len(dst_ptrs), self_index, int(dst_offset_bytes), nbytes,
                              N.stream_ptr(src.device)), "kge_peer_copy")
    if len(dst_ptrs) > 1 and nbytes:
        N.count_launch()


def peer_signal(flag_ptrs, slot, value, device):
    """flags[r][slot] = value on every peer r, released at system scope after all prior work of the stream."""
    lib = N.load()
    arr = (C.c_void_p * len(flag_ptrs))(*[int(p) for p in flag_ptrs])
    N.check(lib.kge_peer_signal(arr, len(flag_ptrs), int(slot), int(value) & 0xFFFFFFFF, N.stream_ptr(device)),
            "kge_peer_signal")
    N.count_launch()


def peer_wait(flags, n, value, status, timeout_s=20.0):
    """Blocks the STREAM (not the host) until flags[0:n] >= value; a time-out raises ``status`` instead of hanging."""
    lib = N.load()
    N.check(lib.kge_peer_wait(N.ptr(flags), int(n), int(value) & 0xFFFFFFFF, int(timeout_s * 1e9), N.ptr(status),
                              N.stream_ptr(flags.device)), "kge_peer_wait")
    N.count_launch()


# ---------------------------------------------------------------------------------------------
# K7: row-sharded entity table (block-cyclic: entity e -> shard e % G, local row e / G)
# ---------------------------------------------------------------------------------------------
def shard_rows(n_entity, n_shards, shard=None):
    """Rows of shard ``shard`` (or, with ``shard=None``, of the largest shard = the common allocation)."""
    if shard is None:
        return -(-n_entity // n_shards)
    return (n_entity - shard + n_shards - 1) // n_shards if shard < n_entity else 0


def split_rows(table, n_shards):
    """Full [N, dim] table -> list of G [ceil(N/G), dim] shards (zero-padded), block-cyclic."""
    rows = shard_rows(table.shape[0], n_shards)
    out = []
    for s in range(n_shards):
        sh = torch.zeros((rows, table.shape[1]), dtype=table.dtype, device=table.device)
        part = table[s::n_shards]
        sh[: part.shape[0]] = part
        out.append(sh)
    return out


def merge_rows(shards, n_entity, out=None):
    """Inverse of split_rows."""
    G = len(shards)
    if out is None:
        out = torch.empty((n_entity, shards[0].shape[1]), dtype=shards[0].dtype, device=shards[0].device)
    for s, sh in enumerate(shards):
        out[s::G] = sh[: shard_rows(n_entity, G, s)]
    return out


class ShardSet:
    """Base pointers of every shard of the entity table (and of its gradient) as seen from THIS GPU:
    local tensors' data_ptr() for shards held here, NVLink peer mappings for the others."""

    def __init__(self, entity_ptrs, grad_ptrs=None, scalar_red=False):
        G = len(entity_ptrs)
        if not 1 <= G <= N.MAX_SHARDS:
            raise ValueError(f"1..{N.MAX_SHARDS} shards supported, got {G}")
        if grad_ptrs is not None and len(grad_ptrs) != G:
            raise ValueError("entity_ptrs and grad_ptrs differ in length")
        self.n_shards = G
        st = N.KgeShards()
        for s in range(G):
            st.entity[s] = int(entity_ptrs[s])
            st.grad_entity[s] = int(grad_ptrs[s]) if grad_ptrs is not None else None
        st.n_shards = G
        st.scalar_red = int(bool(scalar_red))
        self._struct = st

    @classmethod
    def of_tensors(cls, shards, grads=None, **kw):
        """All shards local (one GPU holding every shard: tests, single-GPU runs of the sharded path)."""
        for t in list(shards) + list(grads or []):
            N.require_cuda(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("shards must be contiguous float32 CUDA tensors")
        obj = cls([t.data_ptr() for t in shards], [t.data_ptr() for t in grads] if grads is not None else None, **kw)
        obj._keep = (list(shards), list(grads or []))
        return obj

    def struct(self):
        return self._struct


def fused_forward_sharded_raw(spec, shards, n_entity, rel, sample, neg, weight, mode, alpha, coef_pos, coef_neg,
                              stats, ws, pos_score=None, neg_score=None):
    """K2 through the shard table (kge_fused_fwd_sharded) into caller-owned buffers."""
    lib = N.load()
    tb = spec.struct_sharded(n_entity, rel)
    B, K = neg.shape
    N.check(lib.kge_fused_fwd_sharded(C.byref(tb), C.byref(shards.struct()), _mode_id(mode), N.ptr(sample), B,
                                      N.ptr(neg), K, N.ptr(weight), alpha, N.ptr(pos_score), N.ptr(neg_score),
                                      N.ptr(coef_pos), N.ptr(coef_neg), N.ptr(stats), N.ptr(ws),
                                      N.stream_ptr(rel.device)), "kge_fused_fwd_sharded")
    N.count_launch()


def fused_backward_sharded_raw(spec, shards, n_entity, rel, sample, neg, mode, coef_pos, coef_neg, stats, g_rel,
                               grad_loss=None):
    """K3 through the shard table: entity-row gradients are ADDED into the owners' gradient shards
    (system-scope REDs, possibly over NVLink), the relation gradient into the local ``g_rel``."""
    lib = N.load()
    tb = spec.struct_sharded(n_entity, rel)
    B, K = neg.shape
    N.check(lib.kge_fused_bwd_sharded(C.byref(tb), C.byref(shards.struct()), _mode_id(mode), N.ptr(sample), B,
                                      N.ptr(neg), K, N.ptr(coef_pos), N.ptr(coef_neg), N.ptr(stats),
                                      N.ptr(grad_loss), N.ptr(g_rel), N.stream_ptr(rel.device)),
            "kge_fused_bwd_sharded")
    N.count_launch()


def score_sharded(spec, shards, n_entity, rel, sample, neg=None, mode=None):
    """``model(sample[, negative_sample, mode])`` of a row-sharded model (no autograd): float32
    ``[B,1]`` / ``[B,K]``."""
    lib = N.load()
    N.require_cuda(rel, sample, neg)
    rel = rel.detach().contiguous()
    sample = _prep_ids(sample, rel.device)
    neg = _prep_ids(neg, rel.device)
    B = sample.shape[0]
    K = 1 if neg is None else neg.shape[1]
    out = torch.empty((B, K), dtype=torch.float32, device=rel.device)
    tb = spec.struct_sharded(n_entity, rel)
    with torch.cuda.device(rel.device):
        N.check(lib.kge_score_fwd_sharded(C.byref(tb), C.byref(shards.struct()), _mode_id(mode if neg is not None else None),
                                          N.ptr(sample), B, N.ptr(neg), 0 if neg is None else K, N.ptr(out),
                                          N.stream_ptr(rel.device)), "kge_score_fwd_sharded")
    N.count_launch()
    return out


def rank_counts_sharded(spec, shards, shard_index, n_entity, rel, queries, mode, csr=None, modulus=None):
    """int64 ``[Q]``: how many unfiltered entities of shard ``shard_index`` outrank each query's positive
    (kge_rank_counts_sharded).  ``rank = 1 + sum over shards``."""
    lib = N.load()
    N.require_cuda(rel, queries)
    rel = rel.detach().contiguous()
    queries = _prep_ids(queries, rel.device)
    Q = queries.shape[0]
    dev = rel.device
    counts = torch.zeros(Q, dtype=torch.int64, device=dev)
    tb = spec.struct_sharded(n_entity, rel, modulus.detach() if modulus is not None else None)
    full = N.KgeTables(None, rel.data_ptr(), int(n_entity), rel.shape[0], spec.hidden_dim, spec.model_id, spec.gamma,
                       spec.embedding_range, None)
    ws = torch.empty(max(lib.kge_rank_workspace_bytes(C.byref(full), Q), 8), dtype=torch.uint8, device=dev)
    fs = csr.to(dev).struct() if csr is not None else None
    with torch.cuda.device(dev):
        N.check(lib.kge_rank_counts_sharded(C.byref(tb), C.byref(shards.struct()), int(shard_index), _mode_id(mode),
                                            N.ptr(queries), Q, C.byref(fs) if fs is not None else None,
                                            N.ptr(counts), None, N.ptr(ws), N.stream_ptr(dev)),
                "kge_rank_counts_sharded")
    N.count_launch(2)
    return counts


# ---------------------------------------------------------------------------------------------
# K4 sampler
# ---------------------------------------------------------------------------------------------
class FilterCSR:
    """Device-resident true-entity sets (mkb/sampling/negative_sampling.py:7-28 as a CSR)."""

    def __init__(self, keys, offsets, members):
        self.keys, self.offsets, self.members = keys, offsets, members

    @property
    def device(self):
        return self.keys.device

    def to(self, device):
        if self.keys.device == torch.device(device):
            return self
        return FilterCSR(self.keys.to(device), self.offsets.to(device), self.members.to(device))

    def struct(self):
        return N.KgeFilterCsr(self.keys.data_ptr(), self.offsets.data_ptr(), self.members.data_ptr(),
                              self.keys.shape[0])


def sample_negatives(csr, sample, mode, size, n_entity, seed, offset, status=None, out=None, sort_rows=True):
    lib = N.load()
    N.require_cuda(sample, csr.keys)
    B = sample.shape[0]
    dev = sample.device
    if out is None:
        out = torch.empty((B, size), dtype=torch.int64, device=dev)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    fs = csr.struct()
    with torch.cuda.device(dev):
        N.check(lib.kge_sample_negatives(C.byref(fs), _mode_id(mode), N.ptr(sample), B, size, n_entity,
                                         C.c_uint64(seed & (2**64 - 1)), C.c_uint64(offset), int(bool(sort_rows)), N.ptr(out),
                                         N.ptr(status), N.stream_ptr(dev)), "kge_sample_negatives")
    N.count_launch()
    return out, status


def filter_pool(csr, sample, mode, size, n_entity, pool, status=None, out=None, positions=None):
    """Reference-pool negatives; ``positions`` (int32 ``[B,size]``, optional) additionally receives the index
    into ``pool`` of every chosen negative (what the pooled tensor-core step gathers its scores with)."""
    lib = N.load()
    N.require_cuda(sample, csr.keys, pool)
    B = sample.shape[0]
    dev = sample.device
    if out is None:
        out = torch.empty((B, size), dtype=torch.int64, device=dev)
    if status is None:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
    fs = csr.struct()
    with torch.cuda.device(dev):
        if positions is not None and (positions.dtype != torch.int32 or not positions.is_contiguous()):
            raise TypeError("positions must be a contiguous int32 tensor")
        N.check(lib.kge_filter_pool_positions(C.byref(fs), _mode_id(mode), N.ptr(sample), B, size, n_entity,
                                              N.ptr(pool), pool.shape[0], N.ptr(out), N.ptr(positions), N.ptr(status),
                                              N.stream_ptr(dev)), "kge_filter_pool_positions")
    N.count_launch()
    return out, status


# ---------------------------------------------------------------------------------------------
# pooled negatives on the tensor cores (DistMult / ComplEx with the reference's shared pool)
# ---------------------------------------------------------------------------------------------
def pooled_workspace(spec, ent, rel, B, K, P):
    lib = N.load()
    tb = spec.struct(ent, rel)
    return torch.empty(max(lib.kge_pooled_workspace_bytes(C.byref(tb), B, K, P), 256), dtype=torch.uint8,
                       device=ent.device)


def pooled_dot_forward_raw(spec, ent, rel, sample, pool, positions, weight, mode, alpha, coef_pos, stats, workspace,
                           loss_ws, pos_score=None, neg_score=None):
    """S = Q·Pool^T + self-adversarial terms (kge_pooled_dot_fwd); keeps Q, Pool and dS in ``workspace``."""
    lib = N.load()
    tb = spec.struct(ent, rel)
    B, K = positions.shape
    N.check(lib.kge_pooled_dot_fwd(C.byref(tb), _mode_id(mode), N.ptr(sample), B, N.ptr(pool), pool.shape[0],
                                   N.ptr(positions), K, N.ptr(weight), alpha, N.ptr(pos_score), N.ptr(neg_score),
                                   N.ptr(coef_pos), N.ptr(stats), N.ptr(workspace), N.ptr(loss_ws),
                                   N.stream_ptr(ent.device)), "kge_pooled_dot_fwd")
    N.count_launch(5)


def pooled_dot_backward_raw(spec, ent, rel, sample, pool, K, mode, coef_pos, stats, g_ent, g_rel, workspace,
                            grad_loss=None):
    """dQ = dS·Pool, dPool = dS^T·Q, chain rule and row scatter (kge_pooled_dot_bwd): ADDS into g_ent / g_rel."""
    lib = N.load()
    tb = spec.struct(ent, rel)
    N.check(lib.kge_pooled_dot_bwd(C.byref(tb), _mode_id(mode), N.ptr(sample), sample.shape[0], N.ptr(pool),
                                   pool.shape[0], K, N.ptr(coef_pos), N.ptr(stats), N.ptr(grad_loss), N.ptr(g_ent),
                                   N.ptr(g_rel), N.ptr(workspace), N.stream_ptr(ent.device)), "kge_pooled_dot_bwd")
    N.count_launch(8)


# ---------------------------------------------------------------------------------------------
# K5 ranking
# ---------------------------------------------------------------------------------------------
def rank_all(spec, ent, rel, queries, mode, csr=None, return_scores=False, modulus=None):
    """Filtered rank of the true entity for every query (int64 ``[Q]``)."""
    lib = N.load()
    ent, rel = _prep_tables(ent.detach(), rel.detach())
    queries = _prep_ids(queries, ent.device)
    Q = queries.shape[0]
    dev = ent.device
    ranks = torch.empty(Q, dtype=torch.int64, device=dev)
    scores = torch.empty((Q, ent.shape[0]), dtype=torch.float32, device=dev) if return_scores else None
    tb = spec.struct(ent, rel, modulus.detach() if modulus is not None else None)
    ws = torch.empty(max(lib.kge_rank_workspace_bytes(C.byref(tb), Q), 8), dtype=torch.uint8, device=dev)
    fs = csr.to(dev).struct() if csr is not None else None
    with torch.cuda.device(dev):
        N.check(lib.kge_rank_all(C.byref(tb), _mode_id(mode), N.ptr(queries), Q,
                                 C.byref(fs) if fs is not None else None, N.ptr(ranks), N.ptr(scores),
                                 N.ptr(ws), N.stream_ptr(dev)), "kge_rank_all")
    N.count_launch(2)
    return (ranks, scores) if return_scores else ranks


# ---------------------------------------------------------------------------------------------
# dense Adam
# ---------------------------------------------------------------------------------------------
def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, zero_grad=False):
    lib = N.load()
    N.require_cuda(param, grad, exp_avg, exp_avg_sq)
    with torch.cuda.device(param.device):
        N.check(lib.kge_adam_step(N.ptr(param), N.ptr(grad), N.ptr(exp_avg), N.ptr(exp_avg_sq), param.numel(),
                                  int(step), lr, beta1, beta2, eps, int(bool(zero_grad)),
                                  N.stream_ptr(param.device)), "kge_adam_step")
    N.count_launch()
