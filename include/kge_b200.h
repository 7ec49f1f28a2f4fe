/*
 * kge_b200.h — C ABI of libkge_b200.so: the B200 (sm_100a) kernels behind the mkb hot path.
 *
 * raphaelsty/mkb is pure Python/PyTorch and has NO foreign-function interface of its own: the
 * "interface" these entry points replace is the sequence of ATen calls its per-batch loop issues.
 * Each entry point below names the reference code (file:line under /root/reference) whose work it
 * takes over.  The Python package `mkb_b200` binds this header with ctypes (see INTEGRATION.md for
 * the stub a reference maintainer would add to mkb itself).
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless the comment says host;
 *   - the library never allocates, frees or keeps device memory, and holds no mutable global
 *     state => safe to call from any host thread; the caller owns every buffer incl. workspaces;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), no hidden syncs;
 *   - return value: 0 = ok, <0 = argument error (KGE_E_*), >0 = a cudaError_t from the launch;
 *   - tables are fp32 row-major; entity rows of ComplEx/RotatE are [re(D) | im(D)] exactly as the
 *     reference's torch.chunk(2, dim=2) expects (mkb/models/rotate.py:76-77, complex.py:70-72);
 *   - ids are int64 (torch.LongTensor), sample is int64[B,3] = (head, relation, tail).
 */
#ifndef KGE_B200_H
#define KGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGE_ABI_VERSION 2

typedef void* kge_stream_t; /* cudaStream_t */

/* mkb/models/{transe,distmult,complex,rotate,protate}.py */
enum kge_model { KGE_TRANSE = 0, KGE_DISTMULT = 1, KGE_COMPLEX = 2, KGE_ROTATE = 3, KGE_PROTATE = 4 };
/* `mode` of BaseModel.batch (mkb/models/base.py:153-164); mode=None uses the tail-batch formula */
enum kge_mode { KGE_TAIL_BATCH = 0, KGE_HEAD_BATCH = 1 };

enum kge_error {
  KGE_OK = 0,
  KGE_E_NULL = -1,      /* a required pointer is NULL */
  KGE_E_SIZE = -2,      /* a size is negative / zero where it must not be / too large for the kernel */
  KGE_E_MODEL = -3,     /* unknown model enum */
  KGE_E_MODE = -4,      /* unknown mode enum */
  KGE_E_ALIGN = -5,     /* a pointer is not 4-byte (float) / 8-byte (int64) aligned */
  KGE_E_UNSUPPORTED = -6 /* shape outside what the fused kernel supports: use the unfused calls */
};

/* The two embedding tables and the model constants: BaseModel.__init__ (mkb/models/base.py:66-100).
 * entity_dim = hidden_dim * (2 for ComplEx/RotatE, else 1); relation_dim = hidden_dim * (2 for
 * ComplEx, else 1).  embedding_range = (gamma + 2) / hidden_dim as the reference stores it (fp32). */
typedef struct kge_tables {
  const float* entity;   /* [n_entity, entity_dim] */
  const float* relation; /* [n_relation, relation_dim] */
  int64_t n_entity;
  int64_t n_relation;
  int32_t hidden_dim;
  int32_t model; /* enum kge_model */
  float gamma;
  float embedding_range;
  /* pRotatE only (mkb/models/protate.py:72,91): DEVICE pointer to the trainable scalar `modulus`
   * (score = gamma - modulus * sum |sin(phase)|); read on the device so a training step never waits
   * for the host.  NULL for the other models. */
  const float* modulus;
} kge_tables_t;

int kge_abi_version(void);
/* Human-readable text for a return code of this library (negative) or of CUDA (positive). */
const char* kge_strerror(int code);
/* Device the calling thread is bound to: SM count and compute capability (host pointers). */
int kge_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * K1  gather -> score.   Replaces BaseModel.batch + Model.forward:
 *     mkb/models/base.py:132-207, transe.py:65-76, distmult.py:63-75, complex.py:65-85,
 *     rotate.py:69-99.
 * neg == NULL : scores[B]    = model(sample)                       (tail-batch formula)
 * neg != NULL : scores[B,K]  = model(sample, negative_sample, mode)
 * ------------------------------------------------------------------------------------------- */
int kge_score_fwd(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                  const int64_t* neg, int64_t K, float* scores, kge_stream_t stream);

/* Backward of K1 (what autograd does at mkb/compose/pipeline.py:236 for one model call):
 * ADDS d(sum grad_scores*scores)/d(table) into the dense grad_entity [n_entity, entity_dim] and
 * grad_relation [n_relation, relation_dim] with atomics (index_select's backward is a dense
 * index_add_).  The caller zero-fills the buffers when it wants '=' instead of '+='. */
int kge_score_bwd(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                  const int64_t* neg, int64_t K, const float* grad_scores, float* grad_entity,
                  float* grad_relation, kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Self-adversarial loss on its own.  Replaces losses.Adversarial.__call__
 * (mkb/losses/adversarial.py:21-30).  stats[4] receives
 *   {S_p = sum_i w_i logsig(p_i), S_n = sum_i w_i sum_j a_ij logsig(-n_ij), W = sum_i w_i,
 *    loss = -(S_p + S_n) / (2 W)}.
 * workspace: kge_loss_workspace_bytes(B) bytes, zero-filled once by the caller before first use.
 * ------------------------------------------------------------------------------------------- */
size_t kge_loss_workspace_bytes(int64_t B);
int kge_adv_loss_fwd(const float* pos_score, const float* neg_score, const float* weight, int64_t B,
                     int64_t K, float alpha, float* stats, void* workspace, kge_stream_t stream);
/* grad_pos[B], grad_neg[B,K] = dL/dpos, dL/dneg (softmax weights detached, adversarial.py:25).
 * grad_loss: device scalar upstream gradient or NULL (= 1).  stats: from the forward (uses W). */
int kge_adv_loss_bwd(const float* pos_score, const float* neg_score, const float* weight, int64_t B,
                     int64_t K, float alpha, const float* stats, const float* grad_loss,
                     float* grad_pos, float* grad_neg, kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * KL divergence between row softmaxes, the distillation loss.  Replaces losses.KlDivergence.__call__
 * (mkb/losses/kl_divergence.py:22-29):
 *   loss[0] = mean_{i,j} q_ij (log q_ij - log p_ij),  p = softmax(student / T, dim=1),
 *                                                      q = softmax(teacher / T, dim=1)
 * student, teacher: [B,K] row-major.  workspace: kge_loss_workspace_bytes(B) bytes, zero-filled once.
 * Backward: grad_student[B,K] = dL/dstudent; grad_teacher (may be NULL) = dL/dteacher; grad_loss is
 * a device scalar or NULL (= 1).
 * ------------------------------------------------------------------------------------------- */
int kge_kl_div_fwd(const float* student, const float* teacher, int64_t B, int64_t K, float T, float* loss,
                   void* workspace, kge_stream_t stream);
int kge_kl_div_bwd(const float* student, const float* teacher, int64_t B, int64_t K, float T,
                   const float* grad_loss, float* grad_student, float* grad_teacher, kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K2  fused gather -> score(1+K) -> adversarial loss, ONE kernel.  Replaces the three calls at
 *     mkb/compose/pipeline.py:211, :230-232, :234 (model(sample); model(sample, neg, mode); loss).
 * Outputs: pos_score[B] and neg_score[B,K] (either may be NULL = not wanted), stats[4] as above,
 *          coef_pos[B], coef_neg[B,K] = the per-score loss gradients up to the common factor
 *          1/(2W):  coef_pos_i = -w_i sig(-p_i),  coef_neg_ij = w_i a_ij sig(n_ij)
 *          (saved for K3; this is what autograd would have recomputed from the saved scores).
 * Returns KGE_E_UNSUPPORTED when K is too large for one CTA's shared memory (use K1 + loss).
 * ------------------------------------------------------------------------------------------- */
int kge_fused_fwd(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                  const int64_t* neg, int64_t K, const float* weight, float alpha, float* pos_score,
                  float* neg_score, float* coef_pos, float* coef_neg, float* stats, void* workspace,
                  kge_stream_t stream);

/* K3  fused backward: recompute the residuals, scatter row gradients with atomics.
 *     Replaces error.backward() (mkb/compose/pipeline.py:236) for the fused forward above.
 * ADDS into grad_entity / grad_relation (caller-zeroed).  The common factor is taken on the device:
 * scale = (grad_loss ? *grad_loss : 1) / (2 * stats[2]); in a multi-GPU run the caller all-reduces
 * stats[0..2] between K2 and K3 so every rank divides by the global sum of weights. */
int kge_fused_bwd(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                  const int64_t* neg, int64_t K, const float* coef_pos, const float* coef_neg,
                  const float* stats, const float* grad_loss, float* grad_entity,
                  float* grad_relation, kge_stream_t stream);

/* K3, one column chunk: the same backward restricted to hidden-dim columns [col0, col0 + ncols)
 * (of both components).  The backward is element-wise in the hidden dim, so the chunks of one step
 * are independent launches; grad_entity_chunk / grad_relation_chunk are dense
 * [n_entity, NC*ncols] / [n_relation, RC*ncols] buffers for that chunk only, which makes each chunk's
 * gradient one contiguous region: the multi-GPU step all-reduces chunk c (and runs its Adam update,
 * kge_adam_step_chunk) while chunk c+1 is still being computed.  col0 and ncols must be multiples
 * of 4 for the vector path.
 * Multi-record batches: with n_records > 1 the sample / neg / coef_pos / coef_neg / stats pointers
 * describe record 0 and record r lives record_stride_bytes further on (the packed, all-gathered
 * step records of the G ranks); each record holds B positives and the global normaliser is the sum
 * of the records' stats[2].  n_records < -1: |n_records| records whose normaliser is already global —
 * `stats` is then ONE buffer (e.g. the all-reduced loss sums), not one per record. */
int kge_fused_bwd_chunk(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                        const int64_t* neg, int64_t K, const float* coef_pos, const float* coef_neg,
                        const float* stats, const float* grad_loss, int32_t col0, int32_t ncols,
                        int32_t n_records, int64_t record_stride_bytes, float* grad_entity_chunk,
                        float* grad_relation_chunk, kge_stream_t stream);

/* pRotatE only: gradient of the trainable scalar `modulus` (mkb/models/protate.py:72,91;
 * score = gamma - modulus * sum_d |sin(phase_d)|), which autograd produces at
 * mkb/compose/pipeline.py:236 next to the table gradients:
 *   grad_modulus[0] += scale * sum_k grad_scores[k] * (scores[k] - gamma) / modulus
 * over n scores saved from the forward.  stats != NULL: scale = (grad_loss ? *grad_loss : 1) /
 * (2 * stats[2]) and grad_scores are K2's coef_pos / coef_neg; stats == NULL: scale = 1 and grad_scores
 * are plain upstream gradients (K1's backward).  One CTA, fixed order. */
int kge_modulus_grad(const float* scores, const float* grad_scores, int64_t n, const float* stats,
                     const float* grad_loss, float gamma, const float* modulus, float* grad_modulus,
                     kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K7  row-sharded entity table: K2 / K3 with every entity row resolved through a table of shard base
 *     pointers (BASELINE config 4 "entity table row-sharded across 4 x B200 with P2P
 *     remote-row gather").  Replaces the same reference code as K2 / K3 — index_select gathers of
 *     mkb/models/base.py:167-205 and their dense index_add_ backward — for a table that is spread over
 *     the HBM of several GPUs.
 * Layout: block-cyclic.  Entity e lives on shard e % n_shards at local row e / n_shards, so shard s is
 *     the fp32 row-major matrix [ceil((n_entity - s) / n_shards), entity_dim].  entity[s] / grad_entity[s]
 *     are DEVICE pointers valid on the calling GPU: local memory for its own shard, peer mappings
 *     (cudaIpc / torch symmetric memory over NVLink) for the others.  Remote rows are read with plain
 *     16-byte loads; row gradients are added into the OWNER's gradient shard with system-scope vector
 *     reductions (red.relaxed.sys.global.add.v4.f32) performed by the owner's L2.
 * tables->entity is ignored (may be NULL); tables->n_entity is the GLOBAL entity count (< 2^31);
 *     the relation table and its gradient stay replicated / local (the caller all-reduces
 *     grad_relation).  hidden_dim must be a multiple of 4 and every shard pointer 16-byte aligned
 *     (KGE_E_UNSUPPORTED / KGE_E_ALIGN otherwise).  scalar_red != 0 issues 4-byte instead of 16-byte
 *     reductions (debugging aid).
 * Synchronisation is the caller's: all ranks' kge_fused_bwd_sharded must have completed before an
 *     owner consumes its gradient shard, and owners must have finished updating their table shard
 *     before anyone's next forward reads it.
 * ------------------------------------------------------------------------------------------- */
#define KGE_MAX_SHARDS 16
typedef struct kge_shards {
  const float* entity[KGE_MAX_SHARDS];
  float* grad_entity[KGE_MAX_SHARDS]; /* backward only */
  int32_t n_shards;
  int32_t scalar_red;
} kge_shards_t;

int kge_fused_fwd_sharded(const kge_tables_t* tables, const kge_shards_t* shards, int mode,
                          const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                          const float* weight, float alpha, float* pos_score, float* neg_score,
                          float* coef_pos, float* coef_neg, float* stats, void* workspace,
                          kge_stream_t stream);
int kge_fused_bwd_sharded(const kge_tables_t* tables, const kge_shards_t* shards, int mode,
                          const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                          const float* coef_pos, const float* coef_neg, const float* stats,
                          const float* grad_loss, float* grad_relation, kge_stream_t stream);
/* Unfused scores through the shard table (model(sample) / model(sample, negative_sample, mode) for a
 * row-sharded model; also what a sharded evaluation uses): same contract as kge_score_fwd. */
int kge_score_fwd_sharded(const kge_tables_t* tables, const kge_shards_t* shards, int mode,
                          const int64_t* sample, int64_t B, const int64_t* neg, int64_t K, float* scores,
                          kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K4  negative sampling on the device.  Replaces NegativeSampling.generate
 *     (mkb/sampling/negative_sampling.py:158-201) and the dictionaries of positive_triples (:7-28),
 *     which become two CSR filters keyed by  relation * n_entity + fixed_entity:
 *        head-batch: key (r,t) -> sorted true heads;   tail-batch: key (h,r) -> sorted true tails.
 * kge_sample_negatives: independent Philox4x32-10 stream per output slot (seed, offset given per
 *     call; the caller advances offset by 1 per call).  Rejected candidates (members of the true
 *     set) are redrawn.  sort_rows != 0 (and K <= 2048) returns every row sorted ascending by id:
 *     the same multiset of negatives (the loss does not depend on their order) in the order that
 *     keeps the scoring kernels' gathers L2-friendly.  status (device int32, OR-ed): bit0 = key not
 *     found (the reference raises KeyError), bit1 = a true set covers every entity.
 * kge_filter_pool: the reference's exact semantics given the batch's host-drawn pool
 *     (RandomState.randint(n_entity, 2*size), :166): each positive takes the first K survivors of
 *     the shared pool, repeating cyclically when fewer survive (:176-195).  bit2 of status = no
 *     pool entry survived (the reference would spin forever).
 * ------------------------------------------------------------------------------------------- */
typedef struct kge_filter_csr {
  const int64_t* keys;    /* [n_keys] sorted ascending */
  const int64_t* offsets; /* [n_keys + 1] */
  const int64_t* members; /* [offsets[n_keys]] sorted inside each segment */
  int64_t n_keys;
} kge_filter_csr_t;

int kge_sample_negatives(const kge_filter_csr_t* filter, int mode, const int64_t* sample, int64_t B,
                         int64_t K, int64_t n_entity, uint64_t seed, uint64_t offset, int sort_rows,
                         int64_t* negatives, int32_t* status, kge_stream_t stream);
int kge_filter_pool(const kge_filter_csr_t* filter, int mode, const int64_t* sample, int64_t B,
                    int64_t K, int64_t n_entity, const int64_t* pool, int64_t pool_size,
                    int64_t* negatives, int32_t* status, kge_stream_t stream);
/* Same, additionally returning positions[B,K] (int32, may be NULL): the index INTO THE POOL of every
 * chosen negative — what the pooled tensor-core path (kge_pooled_dot_*) gathers its scores with. */
int kge_filter_pool_positions(const kge_filter_csr_t* filter, int mode, const int64_t* sample, int64_t B,
                              int64_t K, int64_t n_entity, const int64_t* pool, int64_t pool_size,
                              int64_t* negatives, int32_t* positions, int32_t* status, kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Pooled negatives on the tensor cores (csrc/pooled.cu) — DistMult / ComplEx training when the whole
 *     batch shares ONE candidate pool, i.e. the reference's own sampler
 *     (mkb/sampling/negative_sampling.py:166: one randint(n_entity, 2*size) pool per generate()).
 *     All (1+K)·B scores are then entries of S[B,P] = Q·Pool^T (Q = per-positive query vectors h∘r /
 *     conj(r)∘t, Pool = the P gathered rows), so K2 / K3's B·K row gathers become three small GEMMs on the
 *     tcgen05 3xTF32 kernel (fp32 tiles when the shape does not allow): S = Q·Pool^T, dQ = dS·Pool,
 *     dPool = dS^T·Q.  Replaces the same reference code as K2 / K3 (mkb/compose/pipeline.py:211-236,
 *     distmult.py:63-75, complex.py:65-85, losses/adversarial.py:21-30) for that sampler.
 * pool[P]: the batch's candidate ids; positions[B,K] (int32): index into the pool of each positive's K
 *     negatives, from kge_filter_pool_positions.  Outputs as K2 (pos_score / neg_score optional).
 * The backward ADDS into grad_entity / grad_relation like K3.  workspace: kge_pooled_workspace_bytes()
 *     bytes, 16-byte aligned, uninitialised; it carries Q, Pool and dS from the forward to the backward,
 *     so the same buffer must be passed to both, untouched in between.  loss_workspace: as K2's.
 * KGE_E_UNSUPPORTED for the distance models (TransE / RotatE / pRotatE: their scores are not dot
 *     products) — use K2 / K3.
 * ------------------------------------------------------------------------------------------- */
size_t kge_pooled_workspace_bytes(const kge_tables_t* tables, int64_t B, int64_t K, int64_t P);
int kge_pooled_dot_fwd(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                       const int64_t* pool, int64_t P, const int32_t* positions, int64_t K,
                       const float* weight, float alpha, float* pos_score, float* neg_score, float* coef_pos,
                       float* stats, void* workspace, void* loss_workspace, kge_stream_t stream);
int kge_pooled_dot_bwd(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                       const int64_t* pool, int64_t P, int64_t K, const float* coef_pos, const float* stats,
                       const float* grad_loss, float* grad_entity, float* grad_relation, void* workspace,
                       kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K5  filtered all-entity ranking.  Replaces TestDataset.__getitem__ (mkb/datasets/base.py:196-241)
 *     + Evaluation.compute_score (mkb/evaluation/evaluation.py:237-263) for head-/tail-batch:
 *     ranks[q] = 1 + #{e unfiltered: s_e > s_pos} + #{e < pos unfiltered: s_e == s_pos}.
 * filter may be NULL (raw ranking).  scores_out (optional, [Q, n_entity]) receives the biased
 * scores the reference would have sorted (filtered slots = s_pos - 1e5).
 * workspace: kge_rank_workspace_bytes(tables, Q) bytes of scratch (query vectors, positive scores,
 * filter segments and, for ComplEx / DistMult, a [Q, entity_dim] copy of the positives' rows: the
 * tensor-core path scores them through the same GEMM as the candidates); contents need not be
 * initialised.  Any number of entities (entity tiles fold over grid.y and grid.z).
 * ------------------------------------------------------------------------------------------- */
size_t kge_rank_workspace_bytes(const kge_tables_t* tables, int64_t Q);
int kge_rank_all(const kge_tables_t* tables, int mode, const int64_t* queries, int64_t Q,
                 const kge_filter_csr_t* filter, int64_t* ranks, float* scores_out, void* workspace,
                 kge_stream_t stream);

/* K5 over ONE shard of a row-sharded table ("shard candidates by owner, each GPU counts
 * #{e in shard: s_e > s_pos}, all-reduce of the counts"): counts[q] = number of unfiltered entities of
 * shard `shard_index` that outrank query q's positive (ties by entity id as in kge_rank_all), so that
 * rank[q] = 1 + sum over shards of counts[q].  The query's own two rows (fixed side, positive) are
 * read through the shard table wherever they live, and every rank computes the positive's score with
 * the same instruction sequence, so no broadcast of it is needed.  fp32 tile kernel for all models.
 * scores_out (optional) is the FULL [Q, n_entity] matrix; only this shard's columns are written. */
int kge_rank_counts_sharded(const kge_tables_t* tables, const kge_shards_t* shards, int32_t shard_index,
                            int mode, const int64_t* queries, int64_t Q, const kge_filter_csr_t* filter,
                            int64_t* counts, float* scores_out, void* workspace, kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K8  exact top-k of every row of a score matrix.  Replaces the argsort-then-slice of
 *     utils.TopK._get_rank (mkb/utils/top_k.py:226-234) and of the distillation loop's top-k negative
 *     sampling (mkb/distillation/top_k_sampling.py:664-677).
 * scores: [rows, cols] with row stride row_stride (floats).  indices[rows,k] receives the columns of
 * the k largest scores of each row in DESCENDING score order, equal scores by ascending column (the
 * order of a stable descending argsort); values[rows,k] (optional) the scores themselves.
 * 1 <= k <= min(cols, 1024) (KGE_E_SIZE / KGE_E_UNSUPPORTED otherwise).  No workspace.
 * ------------------------------------------------------------------------------------------- */
int kge_topk_rows(const float* scores, int64_t rows, int64_t cols, int64_t row_stride, int32_t k,
                  int64_t* indices, float* values, kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Dense Adam step over one table (the user-owned torch.optim.Adam of README.md:123-126 called at
 * mkb/compose/pipeline.py:238-240), fused with zeroing the gradient for the next step.
 * step is 1-based.  zero_grad != 0 clears grad after use (optimizer.zero_grad()).
 * ------------------------------------------------------------------------------------------- */
int kge_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  int64_t step, float lr, float beta1, float beta2, float eps, int zero_grad,
                  kge_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * By-entity backward with the optimizer fused in (csrc/byent.cu): an atomics-free alternative to
 * kge_fused_bwd + kge_adam_step for the ENTITY table of the device-resident step — replaces
 * error.backward(); optimizer.step(); optimizer.zero_grad() (mkb/compose/pipeline.py:236-240).
 * Same inputs as kge_fused_bwd (K2's coefficients and stats).  The candidate-row gradients are formed
 * entity by entity from a per-step CSR of (positive, candidate) pairs, summed in a fixed order in
 * registers, and consumed by torch.optim.Adam's update of that row in place: the entity gradient never
 * exists in memory.  The relation table is handled the same way (per-positive relation-gradient rows,
 * summed per relation in index order, Adam in place), so no atomic touches a float anywhere and the step
 * is bit-reproducible run to run.  entity_table / relation_table must be tables->entity / ->relation
 * (updated IN PLACE); exp_avg* are their dense moments.
 * hidden_dim % 4 == 0; step is 1-based; workspace: kge_byent_workspace_bytes(tables, B, K) bytes of
 * scratch, 16-byte aligned, contents need not be initialised.
 * ------------------------------------------------------------------------------------------- */
size_t kge_byent_workspace_bytes(const kge_tables_t* tables, int64_t B, int64_t K);
int kge_bwd_by_entity_adam(const kge_tables_t* tables, int mode, const int64_t* sample, int64_t B,
                           const int64_t* neg, int64_t K, const float* coef_pos, const float* coef_neg,
                           const float* stats, const float* grad_loss, float* entity_table, float* exp_avg,
                           float* exp_avg_sq, float* relation_table, float* rel_exp_avg, float* rel_exp_avg_sq,
                           int64_t step, float lr, float beta1, float beta2, float eps, void* workspace,
                           kge_stream_t stream);

/* Adam on one column chunk: grad_chunk is dense [rows, comps*ncols]; param and the moments keep the
 * table layout (row stride row_stride floats, second component at +im_off, chunk at column col0). */
int kge_adam_step_chunk(float* param, float* grad_chunk, float* exp_avg, float* exp_avg_sq, int64_t rows,
                        int32_t comps, int32_t ncols, int32_t col0, int32_t row_stride, int32_t im_off,
                        int64_t step, float lr, float beta1, float beta2, float eps, int zero_grad,
                        kge_stream_t stream);

/* Debug aid for the TMA-staged variants of K2 / K3 (csrc/score_tma.cuh, selected with KGE_FWD_TMA=1 /
 * KGE_BWD_TMA=1): their mbarrier waits are bounded so that a copy that never completes cannot hang the device;
 * this returns (and clears) a non-zero value if any wait gave up since the last call.  Synchronises. */
int kge_tma_fail_flag(void);

/* ---------------------------------------------------------------------------------------------
 * NVLink peer-memory handshakes of the multi-GPU step (csrc/peer.cu).  The reference has no multi-GPU
 * path (single device, mkb/compose/pipeline.py:183-187); these replace the NCCL collectives a
 * data-parallel wrapper around its loop would issue per step.  All pointer arrays are HOST arrays of
 * n_peers (<= 16) DEVICE pointers into peer-mapped (symmetric) memory, index = rank.
 *   kge_peer_copy   : src[0:bytes] -> dst_peers[r] + dst_offset_bytes for every r != self (16-byte units).
 *   kge_peer_signal : flag_peers[r][slot] = value on every peer r (self included), released at system
 *                     scope: everything launched before it on `stream` is visible to whoever sees the flag.
 *   kge_peer_wait   : returns (on the stream) once flags[r] >= value for all r < n; gives up after
 *                     timeout_ns and ORs bit (r & 15) into *status instead of hanging the device.
 * ------------------------------------------------------------------------------------------- */
int kge_peer_copy(const void* src, void* const* dst_peers, int32_t n_peers, int32_t self,
                  int64_t dst_offset_bytes, int64_t bytes, kge_stream_t stream);
int kge_peer_signal(void* const* flag_peers, int32_t n_peers, int32_t slot, uint32_t value, kge_stream_t stream);
int kge_peer_wait(const uint32_t* flags, int32_t n, uint32_t value, int64_t timeout_ns, int32_t* status,
                  kge_stream_t stream);

/* Column-parallel multi-GPU step (fused update + all-gather over NVLink peer memory): this rank owns
 * columns [col0, col0+ncols) of every row.  grad_slice / exp_avg_slice / exp_avg_sq_slice are dense
 * [rows, comps*ncols] slice buffers; param_replicas is a HOST array of n_replicas (<= 16) DEVICE
 * pointers to the table on every GPU (peer-mapped, e.g. from torch symmetric memory), self_index the
 * local one.  Reads the local replica, applies Adam, stores the new values into all replicas. */
int kge_adam_slice_bcast(float* const* param_replicas, int32_t n_replicas, int32_t self_index,
                         float* grad_slice, float* exp_avg_slice, float* exp_avg_sq_slice, int64_t rows,
                         int32_t comps, int32_t ncols, int32_t col0, int32_t row_stride, int32_t im_off,
                         int64_t step, float lr, float beta1, float beta2, float eps, int zero_grad,
                         kge_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KGE_B200_H */
