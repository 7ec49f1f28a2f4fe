"""Device-resident training step: sampler -> fused forward -> fused backward -> dense Adam, five
kernel launches on one stream, no host synchronisation and no allocation per step.

This is the loop body of mkb/compose/pipeline.py:206-242 with every tensor the step touches kept in
HBM (tables, gradients, Adam moments, the batch's negatives, per-score coefficients).  It is what
``Pipeline.learn`` reduces to when nothing on the step needs the host, and what ``bench.py`` times.

Multi-GPU (one process per GPU, torch.distributed/NCCL): tables are replicated, each rank scores
its own positives; the three loss sums are all-reduced between forward and backward so every rank
normalises by the global sum of weights (losses/adversarial.py:28-30 over the global batch), and the
dense gradients are all-reduced before the (replicated) Adam step.
"""
from __future__ import annotations

import torch

from .. import ops
from . import parallel

__all__ = ["DeviceTrainer"]


class DeviceTrainer:
    @classmethod
    def from_optimizer(cls, model, sampling, optimizer, alpha=0.5, max_batch=1024, **kw):
        """Adopt the hyper-parameters (and, if any, the moments) of a ``mkb_b200.optim.DenseAdam`` built
        over ``model.parameters()`` and keep that optimizer's state pointing at the trainer's buffers,
        so ``optimizer.state_dict()`` stays meaningful after training."""
        group = optimizer.param_groups[0]
        t = cls(model, sampling, lr=group["lr"], betas=tuple(group["betas"]), eps=group["eps"], alpha=alpha,
                max_batch=max_batch, **kw)
        for p, m, v in ((model.entity_embedding, t.m_ent, t.v_ent), (model.relation_embedding, t.m_rel, t.v_rel)):
            st = optimizer.state[p]
            if st:
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                t.t = max(t.t, int(st["step"]))
            st["exp_avg"], st["exp_avg_sq"], st["step"] = m, v, t.t
        t._optimizer = optimizer
        return t

    def sync_optimizer_state(self):
        opt = getattr(self, "_optimizer", None)
        if opt is not None:
            for p in (self.model.entity_embedding, self.model.relation_embedding):
                opt.state[p]["step"] = self.t

    def __init__(self, model, sampling, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, alpha=0.5, max_batch=1024,
                 process_group=None, distributed=False, chunks=None):
        ent, rel = model.entity_embedding, model.relation_embedding
        if not ent.is_cuda:
            raise ops.N.KgeError("DeviceTrainer needs the model on a CUDA device")
        self.model, self.sampling = model, sampling
        self.spec = model.spec
        self.ent, self.rel = ent.data, rel.data
        self.dev = ent.device
        self.lr, self.betas, self.eps, self.alpha = lr, betas, eps, alpha
        self.distributed = distributed
        self.group = process_group
        K = sampling.size
        f32 = dict(dtype=torch.float32, device=self.dev)
        # Gradient storage.  chunks == 1: one flat buffer in the tables' layout.  chunks > 1: the hidden
        # dim is cut into column chunks and the buffer is chunk-major ([entity chunk c | relation chunk c]
        # contiguous), so the backward of chunk c, its all-reduce and its Adam update form a pipeline:
        # while the main stream computes chunk c+1, a side stream reduces and applies chunk c.
        D = model.hidden_dim
        self.nc = self.ent.shape[1] // D
        self.rc = self.rel.shape[1] // D
        if chunks is None:
            chunks = 4 if distributed else 1
        if D % 4 != 0 or D < 128 * chunks:
            chunks = 1
        self.chunks = []
        self._gflat = torch.zeros(self.ent.numel() + self.rel.numel(), **f32)
        if chunks == 1:
            self.g_ent = self._gflat[: self.ent.numel()].view_as(self.ent)
            self.g_rel = self._gflat[self.ent.numel():].view_as(self.rel)
        else:
            width = -(-D // chunks // 32) * 32
            off, col = 0, 0
            while col < D:
                w = min(width, D - col)
                ne, nr = self.ent.shape[0] * self.nc * w, self.rel.shape[0] * self.rc * w
                flat = self._gflat[off: off + ne + nr]
                self.chunks.append((col, w, flat, flat[:ne].view(self.ent.shape[0], self.nc * w),
                                    flat[ne:].view(self.rel.shape[0], self.rc * w)))
                off += ne + nr
                col += w
            self.side = torch.cuda.Stream(device=self.dev)
            self._chunk_done = [torch.cuda.Event() for _ in self.chunks]
        self.m_ent, self.v_ent = torch.zeros_like(self.ent), torch.zeros_like(self.ent)
        self.m_rel, self.v_rel = torch.zeros_like(self.rel), torch.zeros_like(self.rel)
        self.neg = torch.empty((max_batch, K), dtype=torch.int64, device=self.dev)
        self.coef_pos = torch.empty(max_batch, **f32)
        self.coef_neg = torch.empty((max_batch, K), **f32)
        self.stats = torch.zeros(4, **f32)
        self.ws = torch.zeros(max(ops.N.load().kge_loss_workspace_bytes(max_batch), 64), dtype=torch.uint8,
                              device=self.dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.max_batch, self.K = max_batch, K
        self.t = 0
        self._csr = {m: sampling._csr("head" if m == "head-batch" else "tail", self.dev)
                     for m in ("head-batch", "tail-batch")}
        self.hooks = None  # optional (pre_fwd, post_fwd, pre_bwd, post_bwd) event recorders for bench.py


    def step(self, sample, weight, mode):
        """One optimisation step on a device-resident batch; returns the device loss scalar (a view
        of the stats buffer, valid until the next step)."""
        B = sample.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds max_batch {self.max_batch}")
        neg = self.neg[:B]
        coef_pos, coef_neg = self.coef_pos[:B], self.coef_neg[:B]
        s = self.sampling
        if s.pool == "reference":  # the reference's host-drawn shared pool: 2K ids cross PCIe
            pool = torch.from_numpy(s._rng.randint(s.n_entity, size=self.K * 2).astype("int64")).to(
                self.dev, non_blocking=True)
            ops.filter_pool(self._csr[mode], sample, mode, self.K, s.n_entity, pool, self.status, neg)
        else:
            ops.sample_negatives(self._csr[mode], sample, mode, self.K, s.n_entity, s.seed, s._calls, self.status, neg,
                                 sort_rows=s.sort_rows)
        s._calls += 1
        h = self.hooks
        if h:
            h[0].record()
        ops.fused_forward_raw(self.spec, self.ent, self.rel, sample, neg, weight, mode, self.alpha, coef_pos,
                              coef_neg, self.stats, self.ws)
        if h:
            h[1].record()
        if self.distributed:
            parallel.allreduce_loss_sums(self.stats, self.group)
        if h:
            h[2].record()
        self.t += 1
        b1, b2 = self.betas
        if not self.chunks:
            ops.fused_backward_raw(self.spec, self.ent, self.rel, sample, neg, mode, coef_pos, coef_neg,
                                   self.stats, self.g_ent, self.g_rel)
            if h:
                h[3].record()
            if self.distributed:
                parallel.allreduce_gradients(self._gflat, self.group)
            ops.adam_step(self.ent, self.g_ent, self.m_ent, self.v_ent, self.t, self.lr, b1, b2, self.eps,
                          zero_grad=True)
            ops.adam_step(self.rel, self.g_rel, self.m_rel, self.v_rel, self.t, self.lr, b1, b2, self.eps,
                          zero_grad=True)
            return self.stats
        main = torch.cuda.current_stream(self.dev)
        D = self.model.hidden_dim
        for c, (col, w, flat, ge, gr) in enumerate(self.chunks):
            ops.fused_backward_chunk_raw(self.spec, self.ent, self.rel, sample, neg, mode, coef_pos, coef_neg,
                                         self.stats, col, w, ge, gr)
            self._chunk_done[c].record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self._chunk_done[c])
                if self.distributed:
                    parallel.allreduce_gradients(flat, self.group)
                ops.adam_step_chunk(self.ent, ge, self.m_ent, self.v_ent, self.nc, w, col, D, self.t, self.lr,
                                    b1, b2, self.eps)
                ops.adam_step_chunk(self.rel, gr, self.m_rel, self.v_rel, self.rc, w, col, D, self.t, self.lr,
                                    b1, b2, self.eps)
        if h:
            h[3].record()
        main.wait_stream(self.side)
        return self.stats

    def loss(self):
        """Global loss of the last step (host float; synchronises)."""
        return float(parallel.loss_from_sums(self.stats).item())
