"""Device-resident training step: sampler -> fused forward -> fused backward -> dense Adam on one
stream, no host synchronisation and no allocation per step.

This is the loop body of mkb/compose/pipeline.py:206-242 with every tensor the step touches kept in
HBM (tables, gradients, Adam moments, the batch's negatives, per-score coefficients).  It is what
``Pipeline.learn`` reduces to when nothing on the step needs the host, and what ``bench.py`` times.

Multi-GPU (one process per GPU, torch.distributed/NCCL, tables replicated, rank r scores its own
positives).  Two schemes:

``colpar`` (default) — batch-parallel forward, COLUMN-parallel backward.
    The backward is element-wise in the hidden dim, so instead of every rank producing a full dense
    gradient that must be all-reduced (2 x table bytes over NVLink), rank r computes the gradient of
    hidden-dim columns slice r for the GLOBAL batch: the per-score coefficients, negatives and triples
    of all ranks are all-gathered (~3 MB per rank), each rank runs the chunked backward over them for
    its columns only (same bytes as a single-GPU backward), applies Adam to its slice (1/G of the
    optimizer work and state) and stores the updated slice into every replica through NVLink peer
    pointers (`kge_adam_slice_bcast`: update + all-gather fused, 1 x table bytes over NVLink).
    Needs the tables in torch symmetric memory; falls back to ``allreduce`` when that is unavailable.

``allreduce`` — every rank computes the full gradient of its positives; one all-reduce of the flat
    gradient buffer; replicated Adam.

``rowshard`` (opt-in; SURVEY §8(e), BASELINE config 4) — the entity table is ROW-sharded block-cyclically
    (entity e on rank e % G, local row e / G), so each GPU holds 1/G of the table, of its gradient and of
    the Adam moments: the scheme for tables that do not fit one GPU's HBM.  The forward gathers remote
    rows with plain loads through NVLink peer pointers (`kge_fused_fwd_sharded`); the backward adds row
    gradients into the OWNER's gradient shard with system-scope vector reductions over NVLink
    (`kge_fused_bwd_sharded`); the relation table (<= 237 rows) stays replicated and its gradient is
    all-reduced, which doubles as the "every rank's backward has landed" point; each rank then runs Adam
    on its shard only, and a 4-byte all-reduce orders the next forward behind every owner's update.
    ``virtual_shards=G`` runs the same kernels in ONE process with all G shards on this GPU (how the
    single-GPU tests and benches exercise the sharded addressing).  At the configs of BASELINE.json the
    tables fit one GPU many times over and (G-1)/G of all gathers would cross NVLink (~8x slower than
    HBM), so this is not the default; see DESIGN.md §6.

``colshard`` (opt-in; the scheme for tables that do not fit one GPU — or whose per-step all-gather would dominate)
    — the tables are sharded by COLUMNS of the hidden dim (tensor parallel): rank r holds columns slice r of EVERY
    row (1/G of table, gradient and Adam state) and never sees the rest.  Every score of the five models is a sum
    over the hidden dim (L1 / complex-modulus distances, dot products), so each rank computes the PARTIAL scores of
    the global batch over its columns (K1 on its sub-table, full-speed local HBM reads), one all-reduce sums the
    [G*B, 1+K] partial scores (4-8 MB: the only bytes that cross NVLink besides the ~2 MB batch records), every rank
    forms loss and per-score gradients of the global batch redundantly, runs the fused backward on its sub-table and
    Adam on its slice.  No table row, gradient row or updated column ever moves: where ``colpar`` ships the whole
    updated table over NVLink every step (493 MB at config 4: 0.5 ms) and ``rowshard`` pulls every remote row,
    this scheme moves a few MB.  See DESIGN.md §6 for the measurements and which scheme suits which table size.

In all of them, every rank normalises by the GLOBAL sum of weights (losses/adversarial.py:28-30 over the
global batch): the three loss sums are all-reduced between forward and backward.

``handshake="peer"`` (colpar default since round 2) — NO NCCL call on the step.  The step records live in
symmetric memory; after its forward a rank pushes its record into every peer (`kge_peer_copy`, NVLink
stores) and raises flag[phase 0][rank] = step there (`kge_peer_signal`); a one-warp wait kernel holds the
stream until all G flags show the step (`kge_peer_wait`); ONE multi-record backward launch covers the global
batch with the loss sums read from the records (no all-reduce); after the fused Adam + all-gather
(`kge_adam_slice_bcast`) flag[phase 1][rank] = step is raised everywhere and the NEXT forward waits for all G
of those.  Round 1 spent ~0.1 ms per step (more at 8 GPUs) in three latency-bound NCCL collectives instead.
``handshake="nccl"`` keeps that flow (and its ``packed_records`` / ``merged_backward`` variants) for A/B runs.
"""
from __future__ import annotations

import torch

from .. import ops
from . import parallel

__all__ = ["DeviceTrainer"]


def _column_slices(D, parts, align=32):
    width = -(-D // parts // align) * align
    out, col = [], 0
    for _ in range(parts):
        w = max(0, min(width, D - col))
        out.append((col, w))
        col += w
    return out


class DeviceTrainer:
    COLSHARD_ABOVE_BYTES = 256 << 20  # entity-table size above which the default multi-GPU scheme is colshard

    @staticmethod
    def modulus_free(model):
        return getattr(model, "kernel_modulus", None) is None  # pRotatE's trainable modulus: single-GPU flow only

    @classmethod
    def from_optimizer(cls, model, sampling, optimizer, alpha=0.5, max_batch=1024, **kw):
        """Adopt the hyper-parameters (and, if any, the moments) of a ``mkb_b200.optim.DenseAdam`` built
        over ``model.parameters()`` and keep that optimizer's state pointing at the trainer's buffers,
        so ``optimizer.state_dict()`` stays meaningful after training (single-GPU / allreduce modes)."""
        group = optimizer.param_groups[0]
        t = cls(model, sampling, lr=float(group["lr"]), betas=tuple(group["betas"]), eps=group["eps"], alpha=alpha,
                max_batch=max_batch, **kw)
        # a stock torch.optim.Adam (the optimizer of the reference's quick-start, README.md:123-126) keeps its
        # step count as a float32 CPU tensor; optim.DenseAdam as an int
        t._step_as_tensor = type(optimizer) is torch.optim.Adam
        if t.mode not in ("colpar", "rowshard", "colshard"):
            for p, m, v in ((model.entity_embedding, t.m_ent, t.v_ent), (model.relation_embedding, t.m_rel, t.v_rel)):
                st = optimizer.state[p]
                if st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    t.t = max(t.t, int(st["step"]))
                st["exp_avg"], st["exp_avg_sq"] = m, v
            t._optimizer = optimizer
            t.sync_optimizer_state()
        return t

    def adopt_hyper_parameters(self, optimizer):
        """Re-read lr / betas / eps from the optimizer's param group (LR schedulers, manual edits)."""
        g = optimizer.param_groups[0]
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(float(b) for b in g["betas"]), float(g["eps"])

    def sync_optimizer_state(self):
        opt = getattr(self, "_optimizer", None)
        if opt is not None:
            for p in (self.model.entity_embedding, self.model.relation_embedding):
                opt.state[p]["step"] = (torch.tensor(float(self.t), dtype=torch.float32)
                                        if getattr(self, "_step_as_tensor", False) else self.t)

    def __init__(self, model, sampling, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, alpha=0.5, max_batch=1024,
                 process_group=None, distributed=False, mode=None, packed_records=False, virtual_shards=None,
                 scalar_red=False, backward="scatter", pooled_gemm=False, merged_backward=False, handshake=None):
        ent, rel = model.entity_embedding, model.relation_embedding
        if not ent.is_cuda:
            raise ops.N.KgeError("DeviceTrainer needs the model on a CUDA device")
        self.model, self.sampling = model, sampling
        self.spec = model.spec
        self.dev = ent.device
        self.lr, self.betas, self.eps, self.alpha = lr, betas, eps, alpha
        self.group = process_group
        self.world = torch.distributed.get_world_size(process_group) if distributed else 1
        self.rank = torch.distributed.get_rank(process_group) if distributed else 0
        self.distributed = distributed and self.world > 1
        D = model.hidden_dim
        self.D = D
        self.nc = ent.shape[1] // D
        self.rc = rel.shape[1] // D
        K = sampling.size
        self.max_batch, self.K = max_batch, K
        f32 = dict(dtype=torch.float32, device=self.dev)

        self.virtual_shards = int(virtual_shards) if virtual_shards else 0
        if self.virtual_shards and self.distributed:
            raise ValueError("virtual_shards is a single-process mode")
        if mode is None:
            mode = "rowshard" if self.virtual_shards else ("colpar" if self.distributed else "single")
            # replicated tables cost one table all-gather over NVLink per step (colpar); above ~256 MB that
            # dominates the step and the column-sharded scheme wins at every GPU count measured (DESIGN.md §6:
            # config 4, 493 MB: 0.80 vs 1.30 ms/step on 4 GPUs; config 2, 116 MB: colpar 0.90 vs 0.94 on 8)
            if (mode == "colpar" and ent.numel() * 4 > self.COLSHARD_ABOVE_BYTES and D % 4 == 0
                    and D >= 32 * self.world and self.modulus_free(model)):
                mode = "colshard"
        if mode == "colshard" and (not self.distributed or D % 4 != 0 or D < 32 * self.world):
            raise ValueError("colshard needs torch.distributed, hidden_dim % 4 == 0 and >= 32 columns per rank")
        if not self.distributed and mode != "rowshard":
            mode = "single"
        if mode == "colpar" and (D % 4 != 0 or D < 32 * self.world or self.world > 16):
            mode = "allreduce"
        self.mode_note = ""
        if mode == "colpar":
            try:
                self._setup_symmetric_tables(model)
            except Exception as e:  # no symmetric memory on this box / build
                self.mode_note = f"colpar unavailable ({type(e).__name__}: {e}); using allreduce"
                mode = "allreduce"
        self.mode = mode
        if handshake not in (None, "peer", "nccl"):
            raise ValueError("handshake must be 'peer' or 'nccl'")
        # peer-memory flags instead of NCCL collectives (colpar only); the nccl variants stay for A/B runs
        self.handshake = (handshake or ("nccl" if (packed_records or merged_backward) else "peer")) if mode == "colpar" else None
        if self.handshake == "peer":
            packed_records = True  # the loss sums ride in the records
        # colpar, opt-in: ONE backward launch over all G gathered records (global stats from the all-reduce)
        # instead of one launch per source rank; keeps the collectives of the measured default flow
        self.merged_backward = bool(merged_backward) and mode == "colpar"
        self.pooled_gemm = False
        self.packed_records = bool(packed_records) and mode == "colpar"
        self.ent, self.rel = model.entity_embedding.data, model.relation_embedding.data
        # pRotatE: the trainable scalar modulus joins the step (single-GPU flow only)
        self.modulus = model.kernel_modulus.data if getattr(model, "kernel_modulus", None) is not None else None
        if self.modulus is not None:
            if mode != "single":
                raise NotImplementedError("pRotatE runs on the single-GPU DeviceTrainer flow only")
            self.pos_score = torch.empty((max_batch, 1), **f32)
            self.neg_score = torch.empty((max_batch, K), **f32)
            self.g_mod = torch.zeros_like(self.modulus)
            self.m_mod, self.v_mod = torch.zeros_like(self.modulus), torch.zeros_like(self.modulus)

        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.stats = torch.zeros(4, **f32)
        self.ws = torch.zeros(max(ops.N.load().kge_loss_workspace_bytes(max_batch), 64), dtype=torch.uint8,
                              device=self.dev)
        self._csr = {m: sampling._csr("head" if m == "head-batch" else "tail", self.dev)
                     for m in ("head-batch", "tail-batch")}
        # every rank draws its own negatives: the same user seed, a rank-specific Philox key
        self._seed = (int(sampling.seed) + 0x9E3779B97F4A7C15 * self.rank) & (2 ** 64 - 1)
        self.t = 0
        self.hooks = None  # optional [pre_fwd, post_fwd, pre_bwd, post_bwd(, pre_adam, post_adam)] CUDA events (bench.py)

        # single-GPU flow: "scatter" = K3 (vector REDs into a dense gradient) + dense Adam;
        # "by_entity" = csrc/byent.cu (per-step CSR by entity, no atomics, Adam fused, deterministic)
        if backward not in ("scatter", "by_entity"):
            raise ValueError("backward must be 'scatter' or 'by_entity'")
        self.backward = backward if mode == "single" else "scatter"
        if self.backward == "by_entity" and D % 4 != 0:
            raise ValueError("backward='by_entity' needs hidden_dim % 4 == 0")
        if mode == "colshard":
            if self.modulus is not None:
                raise NotImplementedError("pRotatE runs on the single-GPU DeviceTrainer flow only")
            self._setup_colshard(model, f32)
        elif mode == "rowshard":
            self._setup_rowshard(model, f32, scalar_red)
        elif mode == "colpar":
            self._setup_colpar(f32)
        else:
            # one flat buffer for both gradients => a single all-reduce in the allreduce mode
            # (the relation part starts on a 16-byte boundary: kge_adam_step / the vector REDs need aligned
            # pointers, and N*dim is not a multiple of 4 for e.g. TransE dim 50 x 14 541 entities)
            rel_off = (self.ent.numel() + 3) // 4 * 4
            self._gflat = torch.zeros(rel_off + self.rel.numel(), **f32)
            self.g_ent = self._gflat[: self.ent.numel()].view_as(self.ent)
            self.g_rel = self._gflat[rel_off: rel_off + self.rel.numel()].view_as(self.rel)
            self.m_ent, self.v_ent = torch.zeros_like(self.ent), torch.zeros_like(self.ent)
            self.m_rel, self.v_rel = torch.zeros_like(self.rel), torch.zeros_like(self.rel)
            self.neg = torch.empty((max_batch, K), dtype=torch.int64, device=self.dev)
            self.coef_pos = torch.empty(max_batch, **f32)
            self.coef_neg = torch.empty((max_batch, K), **f32)
            if self.backward == "by_entity":
                self._byent_ws = ops.byent_workspace(self.spec, self.ent, self.rel, max_batch, K, self.modulus)
            # opt-in: DistMult / ComplEx with the reference's shared pool score the whole batch as ONE GEMM
            # S = Q·Pool^T on the tensor cores (csrc/pooled.cu) instead of B·K row gathers
            self.pooled_gemm = (bool(pooled_gemm) and mode == "single" and sampling.pool == "reference"
                                and self.spec.model_name in ("DistMult", "ComplEx"))
            if self.pooled_gemm:
                self.neg_pos = torch.empty((max_batch, K), dtype=torch.int32, device=self.dev)
                self._pool_ws = ops.pooled_workspace(self.spec, self.ent, self.rel, max_batch, K, 2 * K)

    # ------------------------------------------------------------------------------------------
    # colpar set-up
    # ------------------------------------------------------------------------------------------
    def _setup_symmetric_tables(self, model):
        """Move both tables into torch symmetric memory and exchange peer pointers (collective)."""
        import torch.distributed._symmetric_memory as symm

        group = self.group if self.group is not None else torch.distributed.group.WORLD
        self._symm = []
        self._replicas = []
        for name in ("entity_embedding", "relation_embedding"):
            p = getattr(model, name)
            buf = symm.empty(*p.shape, dtype=torch.float32, device=self.dev)
            buf.copy_(p.data)
            hdl = symm.rendezvous(buf, group)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            if len(ptrs) != self.world or ptrs[self.rank] != buf.data_ptr():
                raise RuntimeError("unexpected symmetric-memory pointer table")
            p.data = buf
            self._symm.append(hdl)
            self._replicas.append(ptrs)
        torch.cuda.synchronize(self.dev)
        torch.distributed.barrier(group=self.group)

    def _setup_colpar(self, f32):
        B, K, G = self.max_batch, self.K, self.world
        self.col0, self.ncols = _column_slices(self.D, G)[self.rank]
        w = self.ncols
        N_, R_ = self.ent.shape[0], self.rel.shape[0]
        self.g_ent = torch.zeros(N_, self.nc * w, **f32)
        self.g_rel = torch.zeros(R_, self.rc * w, **f32)
        self.m_ent, self.v_ent = torch.zeros_like(self.g_ent), torch.zeros_like(self.g_ent)
        self.m_rel, self.v_rel = torch.zeros_like(self.g_rel), torch.zeros_like(self.g_rel)
        # per-rank step record, packed so ONE all-gather moves everything the backward needs
        # (it also carries the rank's three loss sums, so no separate all-reduce is needed)
        o_sample, o_neg = 0, B * 24
        o_cpos = o_neg + B * K * 8
        o_cneg = o_cpos + B * 4
        o_stats = (o_cneg + B * K * 4 + 15) // 16 * 16
        rec = o_stats + (16 if self.packed_records else 0)
        self._rec_stride = rec
        if self.handshake == "peer":
            # records + two flag arrays (phase 0: "record of step t is here", phase 1: "table slice of step t is
            # here"; 16 uint32 each) in ONE symmetric allocation, so peers can store into both
            import torch.distributed._symmetric_memory as symm

            group = self.group if self.group is not None else torch.distributed.group.WORLD
            buf = symm.empty(G * rec + 128, dtype=torch.uint8, device=self.dev)
            buf.zero_()
            torch.cuda.synchronize(self.dev)
            hdl = symm.rendezvous(buf, group)
            self._symm.append(hdl)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            if len(ptrs) != G or ptrs[self.rank] != buf.data_ptr():
                raise RuntimeError("unexpected symmetric-memory pointer table")
            self._rec_ptrs = ptrs
            self._flag_ptrs = [[p + G * rec + 64 * ph for p in ptrs] for ph in (0, 1)]
            self._flags = [buf[G * rec + 64 * ph: G * rec + 64 * ph + 64].view(torch.int32) for ph in (0, 1)]
            self._peer_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self._rec_all = buf[: G * rec]
            torch.cuda.synchronize(self.dev)
            torch.distributed.barrier(group=self.group)  # every rank's flags are zero before anyone signals
        else:
            self._rec_all = torch.zeros(G * rec, dtype=torch.uint8, device=self.dev)
        self._recs = []
        for r in range(G):
            base = self._rec_all[r * rec:(r + 1) * rec]
            self._recs.append((base[o_sample:o_neg].view(torch.int64).view(B, 3),
                               base[o_neg:o_cpos].view(torch.int64).view(B, K),
                               base[o_cpos:o_cneg].view(torch.float32),
                               base[o_cneg:o_cneg + B * K * 4].view(torch.float32).view(B, K),
                               base[o_stats:o_stats + 16].view(torch.float32) if self.packed_records else None))
        self._rec_local = self._rec_all[self.rank * rec:(self.rank + 1) * rec]
        _, self.neg, self.coef_pos, self.coef_neg, self._stats_local = self._recs[self.rank]
        if self.packed_records:  # [G,4] strided view of every record's loss sums
            self._stats_all = self._rec_all.view(G, rec)[:, o_stats:o_stats + 16].view(torch.float32)
        self._tiny = torch.zeros(1, **f32)

    # ------------------------------------------------------------------------------------------
    # colshard: tables sharded by columns of the hidden dim
    # ------------------------------------------------------------------------------------------
    def _setup_colshard(self, model, f32):
        import torch.distributed._symmetric_memory as symm

        # blocks of the global batch travel in 16-byte units: the per-rank batch is rounded up to a multiple of 4
        # (short batches are padded with weight-0 rows anyway)
        self.max_batch = (self.max_batch + 3) // 4 * 4
        G, r, B, K = self.world, self.rank, self.max_batch, self.K
        self.col0, self.ncols = _column_slices(self.D, G)[r]
        w = self.ncols
        if w <= 0 or w % 4:
            raise ValueError(f"colshard: rank {r} would own {w} columns of {self.D}")
        full_e, full_r = model.entity_embedding.data, model.relation_embedding.data
        N_, R_ = full_e.shape[0], full_r.shape[0]
        cut = lambda t, comps: t.view(t.shape[0], comps, self.D)[:, :, self.col0:self.col0 + w].reshape(t.shape[0], comps * w).contiguous()
        self.ent_loc, self.rel_loc = cut(full_e, self.nc), cut(full_r, self.rc)
        # a model of hidden_dim = w over this rank's columns; embedding_range (RotatE's phase divisor) stays global
        self.spec_loc = ops.TableSpec(self.spec.model_name, w, self.spec.gamma, self.spec.embedding_range)
        self.g_ent, self.g_rel = torch.zeros_like(self.ent_loc), torch.zeros_like(self.rel_loc)
        self.m_ent, self.v_ent = torch.zeros_like(self.ent_loc), torch.zeros_like(self.ent_loc)
        self.m_rel, self.v_rel = torch.zeros_like(self.rel_loc), torch.zeros_like(self.rel_loc)
        # the global batch (G blocks of B positives: triples, negatives, weights) + two flag arrays, symmetric
        al = lambda x: (x + 15) // 16 * 16
        o_neg = al(G * B * 24)
        o_w = o_neg + al(G * B * K * 8)
        o_flags = o_w + al(G * B * 4)
        buf = symm.empty(o_flags + 128, dtype=torch.uint8, device=self.dev)
        buf.zero_()
        torch.cuda.synchronize(self.dev)
        group = self.group if self.group is not None else torch.distributed.group.WORLD
        hdl = symm.rendezvous(buf, group)
        self._symm = [hdl]
        ptrs = [int(x) for x in hdl.buffer_ptrs]
        if len(ptrs) != G or ptrs[r] != buf.data_ptr():
            raise RuntimeError("unexpected symmetric-memory pointer table")
        self._gb_ptrs = ptrs
        self._gb_off = (0, o_neg, o_w)
        self.sample_all = buf[: G * B * 24].view(torch.int64).view(G * B, 3)
        self.neg_all = buf[o_neg: o_neg + G * B * K * 8].view(torch.int64).view(G * B, K)
        self.weight_all = buf[o_w: o_w + G * B * 4].view(torch.float32)
        self._flag_ptrs = [[p + o_flags + 64 * ph for p in ptrs] for ph in (0, 1)]
        self._flags = [buf[o_flags + 64 * ph: o_flags + 64 * ph + 64].view(torch.int32) for ph in (0, 1)]
        self._peer_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.neg = self.neg_all[r * B:(r + 1) * B]  # the sampler writes this rank's block in place
        # partial scores of the global batch, negatives first then positives: ONE all-reduce
        self._sc = torch.zeros(G * B * (K + 1), **f32)
        self._sc_neg, self._sc_pos = self._sc[: G * B * K].view(G * B, K), self._sc[G * B * K:]
        self._g_neg, self._g_pos = torch.zeros(G * B, K, **f32), torch.zeros(G * B, **f32)
        self._loss_ws = torch.zeros(max(ops.N.load().kge_loss_workspace_bytes(G * B), 64), dtype=torch.uint8,
                                    device=self.dev)
        self._unit_stats = torch.tensor([0.0, 0.0, 0.5, 0.0], **f32)  # scale 1 / (2 * 0.5): gradients arrive final
        self.coef_pos = self.coef_neg = None
        torch.cuda.synchronize(self.dev)
        torch.distributed.barrier(group=self.group)

    def _step_colshard(self, sample, weight, mode, h):
        G, r, B, K = self.world, self.rank, self.max_batch, self.K
        b = sample.shape[0]
        if b < B:  # short last batch: weight-0 copies of row 0 add nothing to the loss sums or to any gradient
            sample = torch.cat([sample, sample[:1].expand(B - b, -1)])
            weight = torch.cat([weight, torch.zeros(B - b, dtype=weight.dtype, device=weight.device)])
        self.sample_all[r * B:(r + 1) * B].copy_(sample)
        self.weight_all[r * B:(r + 1) * B].copy_(weight)
        self._sample(self.sample_all[r * B:(r + 1) * B], mode, self.neg)
        # peers are done reading the previous global batch before anybody overwrites its blocks
        ops.peer_wait(self._flags[1], G, self.t, self._peer_status)
        for off, blk in zip(self._gb_off, (self.sample_all, self.neg_all, self.weight_all)):
            mine = blk[r * B:(r + 1) * B]
            nbytes = mine.numel() * mine.element_size()
            ops.peer_copy(mine, self._gb_ptrs, r, off + r * nbytes, nbytes)
        self.t += 1
        ops.peer_signal(self._flag_ptrs[0], r, self.t, self.dev)
        ops.peer_wait(self._flags[0], G, self.t, self._peer_status)
        if h:
            h[0].record()
        # partial scores of the GLOBAL batch over this rank's columns (K1 on the sub-table)
        ops.score_forward_raw(self.spec_loc, self.ent_loc, self.rel_loc, self.sample_all, None, mode, self._sc_pos)
        ops.score_forward_raw(self.spec_loc, self.ent_loc, self.rel_loc, self.sample_all, self.neg_all, mode, self._sc_neg)
        if h:
            h[1].record()
        torch.distributed.all_reduce(self._sc, group=self.group)  # the path's one real exchange: sum over column slices
        if self.spec.model_name in ("TransE", "RotatE", "pRotatE"):
            self._sc.sub_((G - 1) * self.spec.gamma)  # every partial carried its own "gamma -"
        ops.adversarial_loss_raw(self._sc_pos, self._sc_neg, self.weight_all, self.alpha, self.stats, self._loss_ws,
                                 self._g_pos, self._g_neg)
        if h:
            h[2].record()
        ops.fused_backward_raw(self.spec_loc, self.ent_loc, self.rel_loc, self.sample_all, self.neg_all, mode,
                               self._g_pos, self._g_neg, self._unit_stats, self.g_ent, self.g_rel)
        if h:
            h[3].record()
        ops.peer_signal(self._flag_ptrs[1], r, self.t, self.dev)  # "I am done with this global batch"
        b1, b2 = self.betas
        if h and len(h) > 5:
            h[4].record()
        ops.adam_step(self.ent_loc, self.g_ent, self.m_ent, self.v_ent, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        ops.adam_step(self.rel_loc, self.g_rel, self.m_rel, self.v_rel, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        if h and len(h) > 5:
            h[5].record()
        return self.stats

    def _gather_columns(self, local, comps, out):
        """All-gather the ranks' column slices of a table into the full [rows, comps * D] layout."""
        G = self.world
        widths = [w for _, w in _column_slices(self.D, G)]
        wmax = max(widths)
        pad = torch.zeros(local.shape[0], comps, wmax, dtype=local.dtype, device=self.dev)
        pad[:, :, : self.ncols] = local.view(local.shape[0], comps, self.ncols)
        parts = torch.empty((G,) + tuple(pad.shape), dtype=local.dtype, device=self.dev)
        torch.distributed.all_gather_into_tensor(parts, pad, group=self.group)
        full = out.view(out.shape[0], comps, self.D)
        for rr, (c0, w) in enumerate(_column_slices(self.D, G)):
            full[:, :, c0:c0 + w] = parts[rr][:, :, :w]

    # ------------------------------------------------------------------------------------------
    # rowshard set-up
    # ------------------------------------------------------------------------------------------
    def _setup_rowshard(self, model, f32, scalar_red):
        """Split the (replicated, identically initialised) entity table into block-cyclic row shards.
        Distributed: this rank keeps shard ``rank`` — table and gradient in torch symmetric memory so the
        peers can read rows / add gradients through NVLink.  virtual_shards: all G shards local."""
        full = model.entity_embedding.data
        self.n_entity = full.shape[0]
        if self.D % 4 != 0:
            raise ValueError("rowshard needs hidden_dim % 4 == 0 (16-byte row segments)")
        B, K = self.max_batch, self.K
        if self.distributed:
            import torch.distributed._symmetric_memory as symm

            G, r = self.world, self.rank
            group = self.group if self.group is not None else torch.distributed.group.WORLD
            rows = ops.shard_rows(self.n_entity, G)
            ptrs, self._symm = [], []
            bufs = []
            for _ in range(2):  # table shard, gradient shard
                buf = symm.empty(rows, full.shape[1], dtype=torch.float32, device=self.dev)
                buf.zero_()
                hdl = symm.rendezvous(buf, group)
                p = [int(x) for x in hdl.buffer_ptrs]
                if len(p) != G or p[r] != buf.data_ptr():
                    raise RuntimeError("unexpected symmetric-memory pointer table")
                bufs.append(buf)
                ptrs.append(p)
                self._symm.append(hdl)
            mine = full[r::G]
            bufs[0][: mine.shape[0]].copy_(mine)
            self.ent_shards, self.g_shards = [bufs[0]], [bufs[1]]  # what this rank owns
            self.shards = ops.ShardSet(ptrs[0], ptrs[1], scalar_red=scalar_red)
            torch.cuda.synchronize(self.dev)
            torch.distributed.barrier(group=self.group)
        else:
            G = self.virtual_shards or 1
            self.ent_shards = ops.split_rows(full, G)
            self.g_shards = [torch.zeros_like(t) for t in self.ent_shards]
            self.shards = ops.ShardSet.of_tensors(self.ent_shards, self.g_shards, scalar_red=scalar_red)
        self.n_shards = G
        self.m_shards = [torch.zeros_like(t) for t in self.ent_shards]
        self.v_shards = [torch.zeros_like(t) for t in self.ent_shards]
        self.g_rel = torch.zeros_like(self.rel)
        self.m_rel, self.v_rel = torch.zeros_like(self.rel), torch.zeros_like(self.rel)
        self.neg = torch.empty((B, K), dtype=torch.int64, device=self.dev)
        self.coef_pos = torch.empty(B, **f32)
        self.coef_neg = torch.empty((B, K), **f32)
        self._tiny = torch.zeros(1, **f32)

    def sync_model(self):
        """Write the trained shards back into ``model.entity_embedding`` (all-gather in the distributed
        case) so evaluation / save / the user's own code see the current table."""
        self.flush()
        if self.mode == "colshard":
            self._gather_columns(self.ent_loc, self.nc, self.model.entity_embedding.data)
            self._gather_columns(self.rel_loc, self.rc, self.model.relation_embedding.data)
            return
        if self.mode != "rowshard":
            return
        full = self.model.entity_embedding.data
        if self.distributed:
            mine = self.ent_shards[0]
            gathered = torch.empty((self.world,) + tuple(mine.shape), dtype=mine.dtype, device=self.dev)
            torch.distributed.all_gather_into_tensor(gathered, mine.contiguous(), group=self.group)
            ops.merge_rows(list(gathered), self.n_entity, out=full)
        else:
            ops.merge_rows(self.ent_shards, self.n_entity, out=full)

    def sharded_ranks(self, evaluation, dataset, mode):
        """Filtered ranks of ``dataset``'s queries straight from the row shards — no gathered table
        (SURVEY §8(e)): every rank counts the entities of ITS shard that outrank each positive
        (`kge_rank_counts_sharded`), one all-reduce sums the counts.  Same ranks on every rank."""
        import numpy as np

        if self.mode != "rowshard":
            raise RuntimeError("sharded_ranks needs mode='rowshard'")
        queries = torch.as_tensor(np.asarray(dataset, dtype=np.int64).reshape(-1, 3)).to(self.dev)
        csr = evaluation._filter("head" if mode == "head-batch" else "tail", self.dev)
        mine = [self.rank] if self.distributed else range(self.n_shards)
        counts = torch.zeros(queries.shape[0], dtype=torch.int64, device=self.dev)
        chunk = max(int(evaluation.batch_size), 1) * 64
        for lo in range(0, queries.shape[0], chunk):
            for s_idx in mine:
                counts[lo:lo + chunk] += ops.rank_counts_sharded(self.spec, self.shards, s_idx, self.n_entity,
                                                                 self.rel, queries[lo:lo + chunk], mode, csr)
        if self.distributed:
            torch.distributed.all_reduce(counts, group=self.group)
        return counts + 1

    def _step_rowshard(self, sample, weight, B, mode, h):
        neg, coef_pos, coef_neg = self.neg[:B], self.coef_pos[:B], self.coef_neg[:B]
        if h:
            h[0].record()
        ops.fused_forward_sharded_raw(self.spec, self.shards, self.n_entity, self.rel, sample, neg, weight, mode,
                                      self.alpha, coef_pos, coef_neg, self.stats, self.ws)
        if h:
            h[1].record()
        self.t += 1
        if self.distributed:
            parallel.allreduce_loss_sums(self.stats, self.group)
        if h:
            h[2].record()
        ops.fused_backward_sharded_raw(self.spec, self.shards, self.n_entity, self.rel, sample, neg, mode, coef_pos,
                                       coef_neg, self.stats, self.g_rel)
        if h:
            h[3].record()
        if self.distributed:
            # sums the replicated relation gradient AND marks "every rank's backward kernel has completed":
            # all remote row-gradient REDs into my shard have landed before my Adam reads it
            torch.distributed.all_reduce(self.g_rel, group=self.group)
        b1, b2 = self.betas
        for p, g, m, v in zip(self.ent_shards, self.g_shards, self.m_shards, self.v_shards):
            ops.adam_step(p, g, m, v, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        ops.adam_step(self.rel, self.g_rel, self.m_rel, self.v_rel, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        if self.distributed:
            # every owner has updated (and re-zeroed the gradient of) its shard before anyone's next step
            torch.distributed.all_reduce(self._tiny, group=self.group)
        return self.stats

    # ------------------------------------------------------------------------------------------
    def _sample(self, sample, mode, neg):
        s = self.sampling
        if s.pool == "reference":  # the reference's host-drawn shared pool: 2K ids cross PCIe
            pool = torch.from_numpy(s._rng.randint(s.n_entity, size=self.K * 2).astype("int64")).to(
                self.dev, non_blocking=True)
            self._pool = pool
            ops.filter_pool(self._csr[mode], sample, mode, self.K, s.n_entity, pool, self.status, neg,
                            positions=self.neg_pos[: sample.shape[0]] if self.pooled_gemm else None)
        else:
            ops.sample_negatives(self._csr[mode], sample, mode, self.K, s.n_entity, self._seed, s._calls, self.status,
                                 neg, sort_rows=s.sort_rows)
        s._calls += 1

    def step(self, sample, weight, mode):
        """One optimisation step on a device-resident batch; returns the device stats buffer
        (S_p, S_n, W, local loss), valid until the next step."""
        B = sample.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} exceeds max_batch {self.max_batch}")
        if self.mode == "colshard":
            return self._step_colshard(sample, weight, mode, self.hooks)
        neg = self.neg[:B]
        coef_pos, coef_neg = self.coef_pos[:B], self.coef_neg[:B]
        self._sample(sample, mode, neg)
        h = self.hooks
        if self.mode == "rowshard":
            return self._step_rowshard(sample, weight, B, mode, h)
        if self.handshake == "peer":
            # every peer's column slice of the previous step has landed in this replica
            ops.peer_wait(self._flags[1], self.world, self.t, self._peer_status)
        if h:
            h[0].record()
        packed = self.packed_records
        mod = self.modulus
        if self.pooled_gemm:
            return self._step_pooled(sample, weight, B, mode, h)
        ops.fused_forward_raw(self.spec, self.ent, self.rel, sample, neg, weight, mode, self.alpha, coef_pos,
                              coef_neg, self._stats_local if packed else self.stats, self.ws,
                              self.pos_score[:B] if mod is not None else None,
                              self.neg_score[:B] if mod is not None else None, modulus=mod)
        if h:
            h[1].record()
        self.t += 1
        b1, b2 = self.betas
        if self.distributed and not packed:
            parallel.allreduce_loss_sums(self.stats, self.group)
        if self.mode == "colpar":
            return self._step_colpar(sample, B, mode, h)
        if h:
            h[2].record()
        if self.backward == "by_entity":
            ops.bwd_by_entity_adam_raw(self.spec, self.ent, self.rel, sample, neg, mode, coef_pos, coef_neg, self.stats,
                                       self.m_ent, self.v_ent, self.m_rel, self.v_rel, self.t, self.lr, b1, b2,
                                       self.eps, self._byent_ws, modulus=mod)
        else:
            ops.fused_backward_raw(self.spec, self.ent, self.rel, sample, neg, mode, coef_pos, coef_neg, self.stats,
                                   self.g_ent, self.g_rel, modulus=mod)
        if h:
            h[3].record()
        if mod is not None:
            ops._modulus_grad(self.spec, self.pos_score[:B], coef_pos, mod, self.g_mod, self.stats)
            ops._modulus_grad(self.spec, self.neg_score[:B], coef_neg, mod, self.g_mod, self.stats)
            ops.adam_step(mod, self.g_mod, self.m_mod, self.v_mod, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        if self.distributed:
            parallel.allreduce_gradients(self._gflat, self.group)
        if h and len(h) > 5:
            h[4].record()
        if self.backward != "by_entity":  # by_entity already applied Adam to both tables
            ops.adam_step(self.ent, self.g_ent, self.m_ent, self.v_ent, self.t, self.lr, b1, b2, self.eps,
                          zero_grad=True)
            ops.adam_step(self.rel, self.g_rel, self.m_rel, self.v_rel, self.t, self.lr, b1, b2, self.eps,
                          zero_grad=True)
        if h and len(h) > 5:
            h[5].record()
        return self.stats

    def _step_pooled(self, sample, weight, B, mode, h):
        """DistMult / ComplEx, reference pool: three GEMMs instead of B*K row gathers (csrc/pooled.cu)."""
        coef_pos = self.coef_pos[:B]
        ops.pooled_dot_forward_raw(self.spec, self.ent, self.rel, sample, self._pool, self.neg_pos[:B], weight, mode,
                                   self.alpha, coef_pos, self.stats, self._pool_ws, self.ws)
        if h:
            h[1].record()
            h[2].record()
        self.t += 1
        b1, b2 = self.betas
        ops.pooled_dot_backward_raw(self.spec, self.ent, self.rel, sample, self._pool, self.K, mode, coef_pos,
                                    self.stats, self.g_ent, self.g_rel, self._pool_ws)
        if h:
            h[3].record()
        ops.adam_step(self.ent, self.g_ent, self.m_ent, self.v_ent, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        ops.adam_step(self.rel, self.g_rel, self.m_rel, self.v_rel, self.t, self.lr, b1, b2, self.eps, zero_grad=True)
        return self.stats

    def _step_colpar(self, sample, B, mode, h):
        # One all-gather moves every rank's step record (triples, negatives, coefficients[, loss sums]).
        # Together with the loss-sum all-reduce before it, it is the "every rank finished its forward"
        # point after which table columns may be overwritten by their owners.
        self._recs[self.rank][0][:B].copy_(sample)
        peer = self.handshake == "peer"
        if peer:
            ops.peer_copy(self._rec_local, self._rec_ptrs, self.rank, self.rank * self._rec_stride)
            ops.peer_signal(self._flag_ptrs[0], self.rank, self.t, self.dev)  # "my forward is done, my record is there"
            ops.peer_wait(self._flags[0], self.world, self.t, self._peer_status)
        else:
            torch.distributed.all_gather_into_tensor(self._rec_all, self._rec_local, group=self.group)
        if h:
            h[2].record()
        if self.ncols > 0 and self.packed_records:
            # the global batch (G records x B positives) in ONE launch, my columns only
            s0, n0, cp0, cn0, st0 = self._recs[0]
            ops.fused_backward_chunk_raw(self.spec, self.ent, self.rel, s0[:B], n0[:B], mode, cp0[:B], cn0[:B], st0,
                                         self.col0, self.ncols, self.g_ent, self.g_rel, n_records=self.world,
                                         record_stride=self._rec_stride)
        elif self.ncols > 0 and self.merged_backward and B == self.max_batch:
            s0, n0, cp0, cn0, _ = self._recs[0]
            ops.fused_backward_chunk_raw(self.spec, self.ent, self.rel, s0[:B], n0[:B], mode, cp0[:B], cn0[:B],
                                         self.stats, self.col0, self.ncols, self.g_ent, self.g_rel,
                                         n_records=-self.world, record_stride=self._rec_stride)
        elif self.ncols > 0:
            for s_r, n_r, cp_r, cn_r, _ in self._recs:  # the global batch, one source rank at a time
                ops.fused_backward_chunk_raw(self.spec, self.ent, self.rel, s_r[:B], n_r[:B], mode, cp_r[:B],
                                             cn_r[:B], self.stats, self.col0, self.ncols, self.g_ent, self.g_rel)
        if h:
            h[3].record()
        b1, b2 = self.betas
        if h and len(h) > 5:
            h[4].record()
        if self.ncols > 0:
            for reps, g, m, v, comps, tbl in ((self._replicas[0], self.g_ent, self.m_ent, self.v_ent, self.nc, self.ent),
                                              (self._replicas[1], self.g_rel, self.m_rel, self.v_rel, self.rc, self.rel)):
                ops.adam_slice_bcast(reps, self.rank, g, m, v, tbl.shape[0], comps, self.ncols, self.col0,
                                     tbl.shape[1], self.D, self.t, self.lr, b1, b2, self.eps, device=self.dev)
        if h and len(h) > 5:
            h[5].record()
        # every replica must hold every slice before anyone's next forward reads the tables
        if peer:
            ops.peer_signal(self._flag_ptrs[1], self.rank, self.t, self.dev)  # waited for at the next step's start
        else:
            torch.distributed.all_reduce(self._tiny, group=self.group)
        if self.packed_records:
            torch.sum(self._stats_all, dim=0, out=self.stats)  # global (S_p, S_n, W, -) for loss()/logging
        return self.stats

    def flush(self):
        """colpar/peer: hold the stream until every peer's slice of the LAST step is in this replica (the next
        step would do it; callers that read the tables — evaluation, save — need it now), then check that no
        handshake timed out (synchronises)."""
        if self.handshake == "peer" or self.mode == "colshard":
            ops.peer_wait(self._flags[1], self.world, self.t, self._peer_status)
            bad = int(self._peer_status.item())
            if bad:
                raise RuntimeError(f"multi-GPU handshake timed out waiting for rank bit mask {bad:#x}")

    def loss(self):
        """Global loss of the last step (host float; synchronises)."""
        self.flush()
        return float(parallel.loss_from_sums(self.stats).item())
