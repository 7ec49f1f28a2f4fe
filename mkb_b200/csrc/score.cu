// K1/K2/K3: gather -> score (-> self-adversarial loss) forward and the atomic-scatter backward.
//
// Work decomposition
//   forward : one CTA per positive triple.  The CTA builds the positive's *query* vector
//             (h∘r, h+r, conj(r)∘t, ... — everything that does not depend on the candidate) once in
//             shared memory, then its 8 warps stream the K candidate rows: one warp per candidate,
//             16-byte coalesced loads of the row, fp32 accumulate, warp-shuffle reduction over the
//             hidden dim.  In the fused variant the K scores never leave shared memory before the
//             softmax-weighted loss terms and the backward coefficients are formed; a ticket
//             counter lets the last CTA fold the per-positive partial sums in a fixed order.
//   backward: one CTA per positive (x K-slices).  Threads own hidden-dim elements, so the
//             candidate-row gradient needs no reduction at all: each thread recomputes its residual,
//             fires a vector RED into the dense gradient row of the candidate, and keeps the
//             query-side gradient in registers; after the K loop one RED per element goes to the
//             head / relation / tail rows of the positive.
//
// Reference being replaced: mkb/models/base.py:132-207 (gathers), transe.py:65-76,
// distmult.py:63-75, complex.py:65-85, rotate.py:69-99 (scores), losses/adversarial.py:21-30,
// and autograd's backward of all of them (compose/pipeline.py:236).
#include <cstdlib>

#include "kge_common.cuh"

#ifndef KGE_FWD_R
#define KGE_FWD_R 2  // candidate rows per warp iteration (shares each shared-memory read of q)
#endif
#ifndef KGE_FWD_MINB
#define KGE_FWD_MINB 4  // resident CTAs per SM the forward kernel is compiled for (caps registers at 64)
#endif
#ifndef KGE_FWD_U
#define KGE_FWD_U 2  // 16-byte chunks in flight per row per lane
#endif
#ifndef KGE_BWD_BULK_DEFAULT
#define KGE_BWD_BULK_DEFAULT 0  // 1: K3b (bulk-reduction row gradients) unless KGE_BWD_BULK=0
#endif

namespace kge {

struct FwdParams {
  const float* ent;
  const float* rel;
  const int64_t* sample;
  const int64_t* neg;
  const float* weight;
  float* pos_score;
  float* neg_score;
  float* coef_pos;
  float* coef_neg;
  float* partials;       // [3, B]
  unsigned int* ticket;  // zero on entry, reset by the last CTA
  float* stats;          // [4]
  int B, K, D, Dp;
  int ent_stride, rel_stride;
  int k_per_cta;
  float gamma, phase_div, alpha;
  const float* modulus;  // pRotatE: device scalar (else null)
  // K7 (row-sharded entity table): base pointer of every shard (local HBM or an NVLink peer mapping).
  // Appended last so the unsharded kernels read their parameters at unchanged offsets.
  const float* shard[KGE_MAX_SHARDS];
  unsigned n_shards;
};
static_assert(sizeof(FwdParams) <= 4096, "kernel parameters are passed by value: 4 KB limit");

// Row of an entity from a split id (kge_common.cuh: shard_split): unsharded = id * stride.
template <bool SHARD, typename P>
__device__ __forceinline__ const float* ent_row(const P& p, int64_t sid) {
  if constexpr (SHARD) {
    return p.shard[(unsigned)(sid >> 32)] + (sid & 0xffffffffll) * (int64_t)p.ent_stride;
  } else {
    return p.ent + sid * (int64_t)p.ent_stride;
  }
}

// pRotatE's trainable modulus (1 for every other model: folded away at compile time)
template <int M, typename P>
__device__ __forceinline__ float load_modulus(const P& p) {
  if constexpr (Traits<M>::kPhase) return __ldg(p.modulus);
  else return 1.f;
}

// ------------------------------------------------------------------------------------------------
// positives only: one warp per triple (model(sample) and the 3-D sample path)
// ------------------------------------------------------------------------------------------------
template <int M, int VEC, bool SHARD = false>
__global__ void __launch_bounds__(kThreads) score_pos_kernel(FwdParams p) {
  using T = Traits<M>;
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (i >= p.B) return;
  const float* h = ent_row<SHARD>(p, shard_split<SHARD>(p.sample[3 * i + 0], p.n_shards));
  const float* r = p.rel + p.sample[3 * i + 1] * (int64_t)p.rel_stride;
  const float* t = ent_row<SHARD>(p, shard_split<SHARD>(p.sample[3 * i + 2], p.n_shards));
  float acc = 0.f;
  for (int d = lane * VEC; d < p.D; d += 32 * VEC) {
    float h0[VEC], h1[VEC] = {}, rr0[VEC], rr1[VEC] = {}, t0[VEC], t1[VEC] = {};
    ld_global<VEC>(h + d, h0);
    ld_global<VEC>(r + d, rr0);
    ld_global<VEC>(t + d, t0);
    if constexpr (T::NC == 2) {
      ld_global<VEC>(h + p.D + d, h1);
      ld_global<VEC>(t + p.D + d, t1);
    }
    if constexpr (T::RC == 2) ld_global<VEC>(r + p.D + d, rr1);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float r0, r1, q0, q1;
      rel_effective<M>(rr0[v], rr1[v], p.phase_div, r0, r1);
      make_query<M, false>(h0[v], h1[v], r0, r1, q0, q1, p.phase_div);
      acc += cand_term<M>(q0, q1, t0[v], t1[v], p.phase_div);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) p.pos_score[i] = finish_score<M>(acc, p.gamma, load_modulus<M>(p));
}

// ------------------------------------------------------------------------------------------------
// R candidate rows at a time against the query in shared memory (one warp).  Every query chunk read
// from shared memory is used for R rows: the L1/TEX data path carries the global rows AND the
// shared-memory reads of q, and ncu showed it (not L2 or HBM) saturating first with R = 1.
// R * U row chunks (x NC components) of 16 bytes are in flight per lane.
// ------------------------------------------------------------------------------------------------
template <int M, int VEC, int R, int U>
__device__ __forceinline__ void rows_reduce(const float* const (&row)[R], const float* __restrict__ q, int D,
                                            int Dp, int lane, float (&out)[R], float pd) {
  using T = Traits<M>;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  for (int d0 = lane * VEC; d0 < D; d0 += 32 * VEC * U) {
    float e0[R][U][VEC], e1[R][U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int d = d0 + u * 32 * VEC;
      if (d < D) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          ld_global<VEC>(row[r] + d, e0[r][u]);
          if constexpr (T::NC == 2) ld_global<VEC>(row[r] + D + d, e1[r][u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int d = d0 + u * 32 * VEC;
      if (d < D) {
        float q0[VEC], q1[VEC] = {};
        ld_shared<VEC>(q + d, q0);
        if constexpr (T::NC == 2) ld_shared<VEC>(q + Dp + d, q1);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            acc[r] += cand_term<M>(q0[v], q1[v], e0[r][u][v], T::NC == 2 ? e1[r][u][v] : 0.f, pd);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = warp_sum(acc[r]);
}

// ------------------------------------------------------------------------------------------------
// forward with candidates: grid = (B, K-slices); FUSED => K-slices == 1 and the loss is folded in
// ------------------------------------------------------------------------------------------------
template <int M, bool HEAD, int VEC, bool FUSED, bool SHARD = false>
__global__ void __launch_bounds__(kThreads, KGE_FWD_MINB) score_neg_kernel(FwdParams p) {
  using T = Traits<M>;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[33];
  float* q = smem;                  // [NC][Dp]
  float* sc = smem + T::NC * p.Dp;  // FUSED: [K]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i = blockIdx.x;
  const int64_t hid = shard_split<SHARD>(p.sample[3 * i + 0], p.n_shards), rid = p.sample[3 * i + 1],
                tidx = shard_split<SHARD>(p.sample[3 * i + 2], p.n_shards);
  const float* fixed = ent_row<SHARD>(p, HEAD ? tidx : hid);
  const float* relrow = p.rel + rid * (int64_t)p.rel_stride;
  const bool want_pos = (FUSED || p.pos_score != nullptr) && blockIdx.y == 0;

  // 1. query -> shared memory (and, in head-batch, the tail-form positive on the fly)
  float pacc = 0.f;
  for (int d = tid * VEC; d < p.D; d += kThreads * VEC) {
    float a0[VEC], a1[VEC] = {}, rr0[VEC], rr1[VEC] = {}, q0[VEC], q1[VEC];
    ld_global<VEC>(fixed + d, a0);
    if constexpr (T::NC == 2) ld_global<VEC>(fixed + p.D + d, a1);
    ld_global<VEC>(relrow + d, rr0);
    if constexpr (T::RC == 2) ld_global<VEC>(relrow + p.D + d, rr1);
    float r0[VEC], r1[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      rel_effective<M>(rr0[v], rr1[v], p.phase_div, r0[v], r1[v]);
      make_query<M, HEAD>(a0[v], a1[v], r0[v], r1[v], q0[v], q1[v], p.phase_div);
    }
    st_shared<VEC>(q + d, q0);
    if constexpr (T::NC == 2) st_shared<VEC>(q + p.Dp + d, q1);
    if (want_pos) {
      // positive = tail-batch formula on (h, r, t)   (compose/pipeline.py:211, mode=None)
      const float* hrow = ent_row<SHARD>(p, hid);
      const float* trow = ent_row<SHARD>(p, tidx);
      float t0[VEC], t1[VEC] = {};
      ld_global<VEC>(trow + d, t0);
      if constexpr (T::NC == 2) ld_global<VEC>(trow + p.D + d, t1);
      if constexpr (HEAD) {
        float h0[VEC], h1[VEC] = {};
        ld_global<VEC>(hrow + d, h0);
        if constexpr (T::NC == 2) ld_global<VEC>(hrow + p.D + d, h1);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float qp0, qp1;
          make_query<M, false>(h0[v], h1[v], r0[v], r1[v], qp0, qp1, p.phase_div);
          pacc += cand_term<M>(qp0, qp1, t0[v], t1[v], p.phase_div);
        }
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) pacc += cand_term<M>(q0[v], q1[v], t0[v], t1[v], p.phase_div);
      }
    }
  }
  float pos = 0.f;
  if (want_pos) {  // uniform across the CTA
    pos = finish_score<M>(block_sum(pacc, red), p.gamma, load_modulus<M>(p));
    if (tid == 0 && p.pos_score) p.pos_score[i] = pos;
  }
  __syncthreads();

  // 2. candidates: warp w takes j = j0 + w, j0 + w + 8, ...; indices are prefetched 32 at a time
  const int j0 = blockIdx.y * p.k_per_cta;
  const int j1 = min(p.K, j0 + p.k_per_cta);
  const int64_t* negrow = p.neg + i * (int64_t)p.K;
  for (int jb = j0 + warp; jb < j1; jb += kWarps * 32) {
    const int jmine = jb + lane * kWarps;
    const int64_t my_id = jmine < j1 ? shard_split<SHARD>(negrow[jmine], p.n_shards) : 0;
    const int cnt = min(32, (j1 - jb + kWarps - 1) / kWarps);
    constexpr int R = KGE_FWD_R, U = KGE_FWD_U;
    for (int m = 0; m < cnt; m += R) {
      const float* rows[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {  // rows past the end re-read the last valid one (result discarded)
        const int64_t id = __shfl_sync(kFull, my_id, min(m + r, cnt - 1));
        rows[r] = ent_row<SHARD>(p, id);
      }
      float acc[R];
      rows_reduce<M, VEC, R, U>(rows, q, p.D, p.Dp, lane, acc, p.phase_div);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (m + r < cnt) {
            const int j = jb + (m + r) * kWarps;
            const float s = finish_score<M>(acc[r], p.gamma, load_modulus<M>(p));
            if (p.neg_score) p.neg_score[i * (int64_t)p.K + j] = s;
            if constexpr (FUSED) sc[j] = s;
          }
        }
      }
    }
  }

  if constexpr (FUSED) {
    // 3. self-adversarial terms for this positive  (losses/adversarial.py:22-30)
    __syncthreads();
    const float w = p.weight[i];
    const float nt = adv_row_terms(sc, p.K, p.alpha, w, p.coef_neg + i * (int64_t)p.K, red);
    if (tid == 0) {
      p.coef_pos[i] = -w * sigmoid(-pos);
      p.partials[i] = w * log_sigmoid(pos);
      p.partials[p.B + i] = w * nt;
      p.partials[2 * p.B + i] = w;
    }
    // 4. last CTA folds the partials in a fixed order (deterministic loss)
    fold_partials(p.partials, p.B, p.ticket, gridDim.x, p.stats, red);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
struct BwdParams {
  const float* ent;
  const float* rel;
  const int64_t* sample;
  const int64_t* neg;
  const float* gpos;   // [B] or null
  const float* gneg;   // [B,K] or null
  const float* stats;  // null => scale 1
  const float* grad_loss;
  float* grad_ent;
  float* grad_rel;
  int B, K, D;  // D = number of hidden-dim columns this launch covers (a column chunk)
  int col0;     // first column of the chunk in the tables
  int im_off;   // offset of the im / second component inside a table row (= hidden_dim)
  int ent_stride, rel_stride;
  int g_im_off;  // same three for the gradient buffers, which may be chunk-major (dense per chunk)
  int g_ent_stride, g_rel_stride;
  int k_per_cta;
  int tpg;  // threads per group (multiple of 32, divides 256)
  // Multi-record batches (the all-gathered global batch of the column-parallel multi-GPU step): the
  // arrays above describe record 0, record r lives rec_stride bytes further on, each holds rec_B
  // positives; stats of all records are summed for the global normaliser.  rec_B == 0: one record.
  int rec_B, n_rec;
  long long rec_stride;
  float phase_div;
  const float* modulus;  // pRotatE: device scalar (else null)
  // K7 (row-sharded entity table): table shards to read, gradient shards to add into (appended last)
  const float* shard[KGE_MAX_SHARDS];
  float* gshard[KGE_MAX_SHARDS];
  unsigned n_shards;
  int scalar_red;
  // by-entity backward (byent.cu), pass 1: per-positive gradient rows of the head / tail [B, NC*D] and of
  // the relation [B, RC*D]
  float* gh_buf;
  float* gt_buf;
  float* gr_buf;
  int shared_stats;  // multi-record launch whose `stats` is ONE already-global buffer (not one per record)
};
static_assert(sizeof(BwdParams) <= 4096, "kernel parameters are passed by value: 4 KB limit");

// Gradient row of an entity from a split id: in the owner's (possibly remote) gradient shard.
template <bool SHARD>
__device__ __forceinline__ float* grad_row(const BwdParams& p, int64_t sid) {
  if constexpr (SHARD) {
    return p.gshard[(unsigned)(sid >> 32)] + (sid & 0xffffffffll) * (int64_t)p.g_ent_stride;
  } else {
    return p.grad_ent + sid * (int64_t)p.g_ent_stride;
  }
}
// Entity-row gradient add: device scope locally, system scope when the row may be on a peer GPU.
template <int VEC, bool SHARD>
__device__ __forceinline__ void red_row(const BwdParams& p, float* dst, const float (&v)[VEC]) {
  if constexpr (SHARD) red_add_sys<VEC>(dst, v, p.scalar_red);
  else red_add<VEC>(dst, v);
}

constexpr int kTileK = 256;

// DQONLY (pass 1 of the by-entity backward, byent.cu): the candidates' row gradients are NOT scattered —
// pass 2 forms them entity by entity without atomics — and the head / tail gradient rows of each positive
// are stored to per-positive buffers instead of being added into the gradient table.
// BULK (K3b, the default for 16-byte-aligned unsharded launches): a candidate row's gradient is not fired as
// 16 RED.128 per lane but written to a shared-memory staging row and added into the gradient table by ONE bulk
// reduction (cp.reduce.async.bulk.global.shared::cta.add.f32, the TMA's reduce path).  ncu showed K3 limited by the
// L1/TEX pipe, which serialises vector REDs at a fraction of its load rate; the staging stores are plain STS.128
// and the reduction itself bypasses L1.  Loads stay LDG.128 (the all-TMA variant in score_tma.cuh lost to this
// kernel: 8 KB of shared memory per row in flight caps the resident warps).  One named barrier per row and group.
constexpr int kBulkSlots = 4;  // staging rows per thread group

// Named barrier of thread group `grp` (tpg threads).  Literal barrier ids: a register id would make ptxas
// reserve all 16 hardware barriers for the CTA.
__device__ __forceinline__ void group_barrier(int grp, int tpg) {
  switch (grp) {
    case 0: asm volatile("bar.sync 1, %0;" ::"r"(tpg) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"r"(tpg) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"r"(tpg) : "memory"); break;
    case 3: asm volatile("bar.sync 4, %0;" ::"r"(tpg) : "memory"); break;
    case 4: asm volatile("bar.sync 5, %0;" ::"r"(tpg) : "memory"); break;
    case 5: asm volatile("bar.sync 6, %0;" ::"r"(tpg) : "memory"); break;
    case 6: asm volatile("bar.sync 7, %0;" ::"r"(tpg) : "memory"); break;
    default: asm volatile("bar.sync 8, %0;" ::"r"(tpg) : "memory"); break;
  }
}

template <int M, bool HEAD, int VEC, bool SHARD = false, bool DQONLY = false, bool BULK = false>
__global__ void __launch_bounds__(kThreads, 3) score_bwd_kernel(BwdParams p) {
  using T = Traits<M>;
  static_assert(!BULK || (VEC == 4 && !SHARD && !DQONLY), "bulk reduction: aligned, local, scattering launches only");
  extern __shared__ __align__(16) float smem[];  // cross-group dq buffer: [G-1][NC][tpg*VEC] (+ BULK: staging rows)
  __shared__ int64_t s_idx[kTileK];
  __shared__ float s_coef[kTileK];

  const int tid = threadIdx.x;
  const int tpg = p.tpg, G = kThreads / tpg;
  const int grp = tid / tpg, lt = tid - grp * tpg;
  // BULK: this group's ring of staging rows ([re | im] like a gradient row) behind the dq buffer
  const int stage_floats = T::NC * p.D;
  float* stage = smem + (size_t)(G - 1) * T::NC * tpg * VEC + (size_t)grp * kBulkSlots * stage_floats;
  int slot = 0;
  int64_t i = blockIdx.x;
  size_t roff = 0;
  if (p.rec_B > 0) {
    const int64_t r = i / p.rec_B;
    i -= r * p.rec_B;
    roff = (size_t)r * (size_t)p.rec_stride;
  }
  const int64_t* smp = reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(p.sample) + roff) + 3 * i;
  const int64_t* negrow =
      p.neg ? reinterpret_cast<const int64_t*>(reinterpret_cast<const char*>(p.neg) + roff) + i * (int64_t)p.K : nullptr;
  const float* gnegrow =
      p.gneg ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.gneg) + roff) + i * (int64_t)p.K : nullptr;
  const int64_t hid = shard_split<SHARD>(smp[0], p.n_shards), rid = smp[1],
                tidx = shard_split<SHARD>(smp[2], p.n_shards);
  const float* hrow = ent_row<SHARD>(p, hid) + p.col0;
  const float* trow = ent_row<SHARD>(p, tidx) + p.col0;
  const float* fixed = HEAD ? trow : hrow;
  const float* relrow = p.rel + rid * (int64_t)p.rel_stride + p.col0;
  float scale = 1.f;
  if (p.stats) {
    float wsum = 0.f;
    for (int r = 0; r < ((p.rec_B > 0 && !p.shared_stats) ? p.n_rec : 1); ++r)
      wsum += __ldg(reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.stats) + (size_t)r * p.rec_stride) + 2);
    scale = (p.grad_loss ? __ldg(p.grad_loss) : 1.f) / (2.f * wsum);
  }
  scale *= load_modulus<M>(p);  // pRotatE: the element derivatives below are those of sum |sin|
  const bool do_pos = (p.gpos != nullptr) && blockIdx.y == 0;
  const float cpos =
      do_pos ? scale * __ldg(reinterpret_cast<const float*>(reinterpret_cast<const char*>(p.gpos) + roff) + i) : 0.f;
  const int j0 = p.neg ? blockIdx.y * p.k_per_cta : 0;
  const int j1 = p.neg ? min(p.K, j0 + p.k_per_cta) : 0;

  for (int cb = 0; cb * tpg * VEC < p.D; ++cb) {
    const int d = (cb * tpg + lt) * VEC;
    const bool active = d < p.D;
    // q is the only per-element state carried through the K loop; the fixed row and the relation
    // are re-read afterwards for the chain rule (L1/L2 hits) instead of pinning 16 registers.
    float q0[VEC] = {}, q1[VEC] = {};
    float dq0[VEC] = {}, dq1[VEC] = {};
    if (active) {
      float a0[VEC], a1[VEC] = {}, rr0[VEC], rr1[VEC] = {};
      ld_global<VEC>(fixed + d, a0);
      if constexpr (T::NC == 2) ld_global<VEC>(fixed + p.im_off + d, a1);
      ld_global<VEC>(relrow + d, rr0);
      if constexpr (T::RC == 2) ld_global<VEC>(relrow + p.im_off + d, rr1);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float r0, r1;
        rel_effective<M>(rr0[v], rr1[v], p.phase_div, r0, r1);
        make_query<M, HEAD>(a0[v], a1[v], r0, r1, q0[v], q1[v], p.phase_div);
      }
    }
    // ---- candidates: group g takes every G-th row of the tile
    for (int jt = j0; jt < j1; jt += kTileK) {
      const int n = min(kTileK, j1 - jt);
      __syncthreads();
      for (int k = tid; k < n; k += kThreads) {
        s_idx[k] = shard_split<SHARD>(negrow[jt + k], p.n_shards);
        s_coef[k] = scale * gnegrow[jt + k];
      }
      __syncthreads();
      if (active || BULK) {  // BULK: idle lanes of a group still take part in its per-row barrier
        constexpr int U = 4;
        for (int jj = grp; jj < n; jj += G * U) {
          float e0[U][VEC] = {}, e1[U][VEC] = {};
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int k = jj + u * G;
            if (k < n && active) {
              const float* row = ent_row<SHARD>(p, s_idx[k]) + p.col0;
              ld_global<VEC>(row + d, e0[u]);
              if constexpr (T::NC == 2) ld_global<VEC>(row + p.im_off + d, e1[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int k = jj + u * G;
            if (k < n) {
              const float c = s_coef[k];
              float* grow = grad_row<SHARD>(p, s_idx[k]);
              float g0[VEC] = {}, g1[VEC] = {};
              if (active) {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                  cand_bwd<M>(q0[v], q1[v], e0[u][v], T::NC == 2 ? e1[u][v] : 0.f, c, g0[v], g1[v],
                              dq0[v], dq1[v], p.phase_div);
              }
              if constexpr (BULK) {
                float* stg = stage + (size_t)slot * stage_floats;
                if (active) {
                  st_shared<VEC>(stg + d, g0);
                  if constexpr (T::NC == 2) st_shared<VEC>(stg + p.D + d, g1);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging row -> visible to the TMA
                // the reduction that last used the NEXT slot has finished reading it (<= kBulkSlots - 2 pending)
                if (lt == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kBulkSlots - 2) : "memory");
                group_barrier(grp, tpg);
                if (lt == 0) {
                  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(grow),
                               "r"((uint32_t)__cvta_generic_to_shared(stg)), "r"((uint32_t)(stage_floats * 4))
                               : "memory");
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                slot = (slot + 1 == kBulkSlots) ? 0 : slot + 1;
              } else if constexpr (!DQONLY) {
                if (active) {
                  red_row<VEC, SHARD>(p, grow + d, g0);
                  if constexpr (T::NC == 2) red_row<VEC, SHARD>(p, grow + p.g_im_off + d, g1);
                }
              }
            }
          }
        }
      }
    }
    // ---- fold the groups' query gradients into group 0
    if (G > 1) {
      __syncthreads();
      if (grp > 0 && active) {
        float* dst = smem + (size_t)(grp - 1) * T::NC * tpg * VEC + lt * VEC;
        st_shared<VEC>(dst, dq0);
        if constexpr (T::NC == 2) st_shared<VEC>(dst + tpg * VEC, dq1);
      }
      __syncthreads();
      if (grp == 0 && active) {
        for (int g = 1; g < G; ++g) {
          const float* src = smem + (size_t)(g - 1) * T::NC * tpg * VEC + lt * VEC;
          float x0[VEC], x1[VEC];
          ld_shared<VEC>(src, x0);
#pragma unroll
          for (int v = 0; v < VEC; ++v) dq0[v] += x0[v];
          if constexpr (T::NC == 2) {
            ld_shared<VEC>(src + tpg * VEC, x1);
#pragma unroll
            for (int v = 0; v < VEC; ++v) dq1[v] += x1[v];
          }
        }
      }
    }
    // ---- positive term + chain rule to the positive's own rows (group 0 only)
    if (grp == 0 && active) {
      float a0[VEC], a1[VEC] = {}, r0[VEC], r1[VEC];
      {
        float rr0[VEC], rr1[VEC] = {};
        ld_global<VEC>(fixed + d, a0);
        if constexpr (T::NC == 2) ld_global<VEC>(fixed + p.im_off + d, a1);
        ld_global<VEC>(relrow + d, rr0);
        if constexpr (T::RC == 2) ld_global<VEC>(relrow + p.im_off + d, rr1);
#pragma unroll
        for (int v = 0; v < VEC; ++v) rel_effective<M>(rr0[v], rr1[v], p.phase_div, r0[v], r1[v]);
      }
      float gt0[VEC] = {}, gt1[VEC] = {};  // -> tail row
      float gh0[VEC] = {}, gh1[VEC] = {};  // -> head row
      float gr0[VEC] = {}, gr1[VEC] = {};  // -> relation row (stored form)
      float dqp0[VEC] = {}, dqp1[VEC] = {};
      float h0[VEC] = {}, h1[VEC] = {};
      if (do_pos) {
        float t0[VEC], t1[VEC] = {};
        ld_global<VEC>(trow + d, t0);
        if constexpr (T::NC == 2) ld_global<VEC>(trow + p.im_off + d, t1);
        if constexpr (HEAD) {
          ld_global<VEC>(hrow + d, h0);
          if constexpr (T::NC == 2) ld_global<VEC>(hrow + p.im_off + d, h1);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float qp0, qp1;
            make_query<M, false>(h0[v], h1[v], r0[v], r1[v], qp0, qp1, p.phase_div);
            cand_bwd<M>(qp0, qp1, t0[v], t1[v], cpos, gt0[v], gt1[v], dqp0[v], dqp1[v], p.phase_div);
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            cand_bwd<M>(q0[v], q1[v], t0[v], t1[v], cpos, gt0[v], gt1[v], dq0[v], dq1[v], p.phase_div);
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float da0, da1, dr0, dr1;
        query_bwd<M, HEAD>(dq0[v], dq1[v], a0[v], a1[v], r0[v], r1[v], da0, da1, dr0, dr1, p.phase_div);
        if constexpr (HEAD) {
          gt0[v] += da0;
          gt1[v] += da1;
          if (do_pos) {
            float dh0, dh1, dpr0, dpr1;
            query_bwd<M, false>(dqp0[v], dqp1[v], h0[v], h1[v], r0[v], r1[v], dh0, dh1, dpr0, dpr1, p.phase_div);
            gh0[v] = dh0;
            gh1[v] = dh1;
            dr0 += dpr0;
            dr1 += dpr1;
          }
        } else {
          gh0[v] = da0;
          gh1[v] = da1;
        }
        rel_bwd<M>(dr0, dr1, r0[v], r1[v], p.phase_div, gr0[v], gr1[v]);
      }
      if constexpr (DQONLY) {
        float* bh = p.gh_buf + (int64_t)blockIdx.x * p.g_ent_stride;
        float* bt = p.gt_buf + (int64_t)blockIdx.x * p.g_ent_stride;
        if constexpr (VEC == 4) {
          *reinterpret_cast<float4*>(bh + d) = make_float4(gh0[0], gh0[1], gh0[2], gh0[3]);
          *reinterpret_cast<float4*>(bt + d) = make_float4(gt0[0], gt0[1], gt0[2], gt0[3]);
          if constexpr (T::NC == 2) {
            *reinterpret_cast<float4*>(bh + p.g_im_off + d) = make_float4(gh1[0], gh1[1], gh1[2], gh1[3]);
            *reinterpret_cast<float4*>(bt + p.g_im_off + d) = make_float4(gt1[0], gt1[1], gt1[2], gt1[3]);
          }
        }
        float* br = p.gr_buf + (int64_t)blockIdx.x * p.g_rel_stride;
        if constexpr (VEC == 4) {
          *reinterpret_cast<float4*>(br + d) = make_float4(gr0[0], gr0[1], gr0[2], gr0[3]);
          if constexpr (T::RC == 2)
            *reinterpret_cast<float4*>(br + p.g_im_off + d) = make_float4(gr1[0], gr1[1], gr1[2], gr1[3]);
        }
        continue;
      }
      float* gh = grad_row<SHARD>(p, hid);
      float* gt = grad_row<SHARD>(p, tidx);
      float* gr = p.grad_rel + rid * (int64_t)p.g_rel_stride;
      if (!HEAD || do_pos) {
        red_row<VEC, SHARD>(p, gh + d, gh0);
        if constexpr (T::NC == 2) red_row<VEC, SHARD>(p, gh + p.g_im_off + d, gh1);
      }
      if (HEAD || do_pos) {
        red_row<VEC, SHARD>(p, gt + d, gt0);
        if constexpr (T::NC == 2) red_row<VEC, SHARD>(p, gt + p.g_im_off + d, gt1);
      }
      red_add<VEC>(gr + d, gr0);
      if constexpr (T::RC == 2) red_add<VEC>(gr + p.g_im_off + d, gr1);
    }
  }
  if constexpr (BULK) {  // every bulk reduction this thread issued has been performed before the CTA retires
    if (lt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

}  // namespace kge
#ifndef KGE_NO_TMA  // (the host-side emulation of the kernels in tests/emu builds without the TMA variants)
#include "score_tma.cuh"  // K2-TMA / K3-TMA: the fused forward and backward with rows staged through the TMA
#endif
namespace kge {

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int validate_tables(const kge_tables_t* t) {
  if (!t || !t->entity || !t->relation) return KGE_E_NULL;
  if (t->model < KGE_TRANSE || t->model > KGE_PROTATE) return KGE_E_MODEL;
  if (t->model == KGE_PROTATE && !t->modulus) return KGE_E_NULL;
  if (t->hidden_dim <= 0 || t->n_entity <= 0 || t->n_relation <= 0) return KGE_E_SIZE;
  if ((reinterpret_cast<uintptr_t>(t->entity) | reinterpret_cast<uintptr_t>(t->relation)) & 3u)
    return KGE_E_ALIGN;
  return KGE_OK;
}

static bool can_vectorize(const kge_tables_t* t, const void* a = nullptr, const void* b = nullptr) {
  // rows (and the im half of complex rows) must start on 16-byte boundaries
  return (t->hidden_dim % 4 == 0) && aligned16(t->entity) && aligned16(t->relation) &&
         (!a || aligned16(a)) && (!b || aligned16(b));
}

static int sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

template <typename Kern>
static int set_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    if (bytes > 200 * 1024) return KGE_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return KGE_OK;
}

template <int M, bool HEAD, int VEC, bool FUSED, bool SHARD = false>
static int launch_neg(const FwdParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = score_neg_kernel<M, HEAD, VEC, FUSED, SHARD>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  kern<<<grid, kThreads, smem, st>>>(p);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

template <bool FUSED>
static int dispatch_neg(int model, int mode, bool vec, const FwdParams& p, dim3 grid, size_t smem,
                        cudaStream_t st) {
#define KGE_CASE(MM)                                                                        \
  case MM:                                                                                  \
    if (mode == KGE_HEAD_BATCH)                                                             \
      return vec ? launch_neg<MM, true, 4, FUSED>(p, grid, smem, st)                        \
                 : launch_neg<MM, true, 1, FUSED>(p, grid, smem, st);                       \
    return vec ? launch_neg<MM, false, 4, FUSED>(p, grid, smem, st)                         \
               : launch_neg<MM, false, 1, FUSED>(p, grid, smem, st);
  switch (model) {
    KGE_CASE(KGE_TRANSE)
    KGE_CASE(KGE_DISTMULT)
    KGE_CASE(KGE_COMPLEX)
    KGE_CASE(KGE_ROTATE)
    KGE_CASE(KGE_PROTATE)
  }
#undef KGE_CASE
  return KGE_E_MODEL;
}

// K7: the sharded instantiations exist for the 16-byte vector path only
template <bool FUSED>
static int dispatch_neg_sharded(int model, int mode, const FwdParams& p, dim3 grid, size_t smem,
                                cudaStream_t st) {
#define KGE_CASE(MM)                                                                              \
  case MM:                                                                                        \
    return mode == KGE_HEAD_BATCH ? launch_neg<MM, true, 4, FUSED, true>(p, grid, smem, st)       \
                                  : launch_neg<MM, false, 4, FUSED, true>(p, grid, smem, st);
  switch (model) {
    KGE_CASE(KGE_TRANSE)
    KGE_CASE(KGE_DISTMULT)
    KGE_CASE(KGE_COMPLEX)
    KGE_CASE(KGE_ROTATE)
    KGE_CASE(KGE_PROTATE)
  }
#undef KGE_CASE
  return KGE_E_MODEL;
}

// Argument checks shared by the sharded entry points; fills the kernel-side pointer tables.
static int validate_sharded(const kge_tables_t* t, const kge_shards_t* sh, bool need_grad) {
  if (!t || !t->relation || !sh) return KGE_E_NULL;
  if (t->model < KGE_TRANSE || t->model > KGE_PROTATE) return KGE_E_MODEL;
  if (t->model == KGE_PROTATE && !t->modulus) return KGE_E_NULL;
  if (t->hidden_dim <= 0 || t->n_entity <= 0 || t->n_relation <= 0 || t->n_entity > INT32_MAX) return KGE_E_SIZE;
  if (sh->n_shards < 1 || sh->n_shards > KGE_MAX_SHARDS) return KGE_E_SIZE;
  if (t->hidden_dim % 4 != 0) return KGE_E_UNSUPPORTED;
  if (!aligned16(t->relation)) return KGE_E_ALIGN;
  for (int s = 0; s < sh->n_shards; ++s) {
    if (!sh->entity[s] || (need_grad && !sh->grad_entity[s])) return KGE_E_NULL;
    if (!aligned16(sh->entity[s]) || (need_grad && !aligned16(sh->grad_entity[s]))) return KGE_E_ALIGN;
  }
  return KGE_OK;
}

static void fill_fwd(FwdParams& p, const kge_tables_t* t, const kge_shards_t* sh = nullptr) {
  if (sh) {
    for (int s = 0; s < sh->n_shards; ++s) p.shard[s] = sh->entity[s];
    p.n_shards = (unsigned)sh->n_shards;
  }
  p.ent = t->entity;
  p.rel = t->relation;
  p.D = t->hidden_dim;
  p.Dp = (t->hidden_dim + 3) & ~3;
  p.ent_stride = t->hidden_dim * entity_comps(t->model);
  p.rel_stride = t->hidden_dim * relation_comps(t->model);
  p.gamma = t->gamma;
  p.phase_div = host_phase_div(t->embedding_range);
  p.modulus = t->modulus;
}

template <int M, bool HEAD, int VEC, bool SHARD = false, bool DQONLY = false, bool BULK = false>
static int launch_bwd(const BwdParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  auto kern = score_bwd_kernel<M, HEAD, VEC, SHARD, DQONLY, BULK>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  kern<<<grid, kThreads, smem, st>>>(p);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

#ifndef KGE_NO_TMA
// K3-TMA launcher (score_tma.cuh): KGE_E_UNSUPPORTED when the shape does not fit (caller uses the scatter kernel).
template <int M, bool HEAD, int UMAX>
static int launch_bwd_tma(const BwdParams& p, int stages, int n_slices, size_t smem, int grid, cudaStream_t st) {
  auto kern = score_bwd_tma_kernel<M, HEAD, UMAX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, kThreads, smem, st>>>(p, stages, n_slices, nullptr);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

static int run_bwd_tma(const kge_tables_t* t, int mode, const BwdParams& p, int n_slices, cudaStream_t st) {
  const int D = t->hidden_dim;
  if (D > 1024 || D % 4 || t->model == KGE_PROTATE) return KGE_E_UNSUPPORTED;
  const size_t row_bytes = (size_t)p.ent_stride * 4;
  const size_t Kc = ((size_t)p.k_per_cta + 3) & ~(size_t)3;
  const size_t fixed = ((size_t)entity_comps(t->model) * ((D + 3) & ~3) + 2 * Kc) * 4 + kWarps * row_bytes;
  const size_t budget = 220 * 1024;
  if (fixed + kWarps * row_bytes + 64 > budget) return KGE_E_UNSUPPORTED;
  int stages = (int)((budget - fixed) / (kWarps * row_bytes + 64));
  if (stages > 8) stages = 8;
  const int want = env_int("KGE_TMA_STAGES", 0);
  if (want > 0 && want < stages) stages = want;
  const size_t smem = fixed + (size_t)stages * (kWarps * row_bytes + 64);
  int64_t items = (int64_t)p.B * n_slices;
  int grid = sm_count();
  if (grid > items) grid = (int)items;
  const int umax = D <= 256 ? 2 : (D <= 512 ? 4 : 8);
#define KGE_TMA_CASE2(MM, HH)                                                                                  \
  return umax == 2 ? launch_bwd_tma<MM, HH, 2>(p, stages, n_slices, smem, grid, st)                            \
                   : (umax == 4 ? launch_bwd_tma<MM, HH, 4>(p, stages, n_slices, smem, grid, st)               \
                                : launch_bwd_tma<MM, HH, 8>(p, stages, n_slices, smem, grid, st));
#define KGE_TMA_CASE(MM)                                                                                       \
  case MM:                                                                                                     \
    if (mode == KGE_HEAD_BATCH) { KGE_TMA_CASE2(MM, true) } else { KGE_TMA_CASE2(MM, false) }
  switch (t->model) {
    KGE_TMA_CASE(KGE_TRANSE)
    KGE_TMA_CASE(KGE_DISTMULT)
    KGE_TMA_CASE(KGE_COMPLEX)
    KGE_TMA_CASE(KGE_ROTATE)
  }
#undef KGE_TMA_CASE
#undef KGE_TMA_CASE2
  return KGE_E_MODEL;
}

#else
static int run_bwd_tma(const kge_tables_t*, int, const BwdParams&, int, cudaStream_t) { return KGE_E_UNSUPPORTED; }
#endif

// col0 / ncols select a column chunk of the hidden dim; ncols <= 0 means the whole row, in which case
// the gradient buffers have the tables' own layout.  With a chunk, grad_ent / grad_rel are dense
// [n_entity, NC*ncols] / [n_relation, RC*ncols] buffers holding just that chunk.
static int run_bwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B, const int64_t* neg,
                   int64_t K, const float* gpos, const float* gneg, const float* stats,
                   const float* grad_loss, float* grad_ent, float* grad_rel, cudaStream_t st,
                   int col0 = 0, int ncols = 0, int n_records = 1, long long record_stride = 0,
                   const kge_shards_t* sh = nullptr, float* gh_buf = nullptr, float* gt_buf = nullptr) {
  BwdParams p{};
  p.gh_buf = gh_buf;
  p.gt_buf = gt_buf;
  p.gr_buf = grad_rel;  // pass 1 of the by-entity backward stores per-positive relation rows there
  if (sh) {  // K7: grad_ent is unused, rows resolve through the shard tables
    for (int s = 0; s < sh->n_shards; ++s) {
      p.shard[s] = sh->entity[s];
      p.gshard[s] = sh->grad_entity[s];
    }
    p.n_shards = (unsigned)sh->n_shards;
    p.scalar_red = sh->scalar_red;
  }
  p.ent = t->entity;
  p.rel = t->relation;
  p.sample = sample;
  p.neg = neg;
  p.gpos = gpos;
  p.gneg = gneg;
  p.stats = stats;
  p.grad_loss = grad_loss;
  p.grad_ent = grad_ent;
  p.grad_rel = grad_rel;
  p.B = (int)B;
  p.K = neg ? (int)K : 0;
  if (n_records < -1) {  // |n_records| records that share one global stats buffer
    n_records = -n_records;
    p.shared_stats = 1;
  }
  if (n_records > 1) {
    p.rec_B = (int)B;
    p.n_rec = n_records;
    p.rec_stride = record_stride;
  }
  const int64_t total = B * (int64_t)(n_records > 1 ? n_records : 1);
  if (total > INT32_MAX) return KGE_E_SIZE;
  const bool chunked = ncols > 0;
  if (chunked && (col0 < 0 || col0 + ncols > t->hidden_dim)) return KGE_E_SIZE;
  p.D = chunked ? ncols : t->hidden_dim;
  p.col0 = chunked ? col0 : 0;
  p.im_off = t->hidden_dim;
  p.ent_stride = t->hidden_dim * entity_comps(t->model);
  p.rel_stride = t->hidden_dim * relation_comps(t->model);
  p.g_im_off = p.D;
  p.g_ent_stride = p.D * entity_comps(t->model);
  p.g_rel_stride = p.D * relation_comps(t->model);
  p.phase_div = host_phase_div(t->embedding_range);
  p.modulus = t->modulus;
  const bool vec = sh ? aligned16(grad_rel)  // validate_sharded checked the rest
                      : can_vectorize(t, grad_ent, grad_rel) && (p.col0 % 4 == 0) && (p.D % 4 == 0);
  if (sh && !vec) return KGE_E_ALIGN;
  const int VEC = vec ? 4 : 1;
  const int chunks = (p.D + VEC - 1) / VEC;
  int tpg = 32;
  while (tpg < kThreads && tpg < chunks) tpg <<= 1;
  p.tpg = tpg;
  const int G = kThreads / tpg;
  // K-slices: enough CTAs to fill the machine when B alone is small
  // K-slices (gridDim.y; CTAs are issued x-fastest, i.e. slice-major): (a) enough CTAs to fill the
  // machine when B is small, (b) with id-sorted rows slice s of every positive covers the same band
  // of the entity table, so table + gradient of the active band stay L2-resident (~64 MB budget).
  int ks = 1;
  if (p.K > 0) {
    const int fill = (4 * sm_count() + (int)total - 1) / (int)total;
    const double tbl = 2.0 * (double)t->n_entity * p.g_ent_stride * sizeof(float);  // table + grad touched
    const int band = (int)(tbl / (64.0 * 1024 * 1024)) + 1;
    const int want = fill > band ? fill : band;
    const int maxks = (p.K + 31) / 32;
    ks = want < 1 ? 1 : (want > maxks ? maxks : want);
    if (const char* e = getenv("KGE_KS")) ks = atoi(e) > 0 ? atoi(e) : ks;
  }
  if (gh_buf) ks = 1;  // pass 1 of the by-entity backward keeps a positive's whole dq in one CTA
  p.k_per_cta = p.K > 0 ? (p.K + ks - 1) / ks : 0;
  if (p.K > 0) ks = (p.K + p.k_per_cta - 1) / p.k_per_cta;
  if (!sh && !chunked && n_records <= 1 && !gh_buf && vec && neg && gneg && env_int("KGE_BWD_TMA", 0)) {
    const int rc_tma = run_bwd_tma(t, mode, p, ks, st);  // A/B switch: rows in AND gradients out through the TMA
    if (rc_tma != KGE_E_UNSUPPORTED) return rc_tma;
  }
  dim3 grid((unsigned)total, (unsigned)ks);
  const size_t smem = (size_t)(G - 1) * entity_comps(t->model) * tpg * VEC * sizeof(float);
  if (gh_buf) {
    if (!vec || !gt_buf || !gpos || !aligned16(gh_buf) || !aligned16(gt_buf)) return KGE_E_UNSUPPORTED;
#define KGE_CASE(MM)                                                                               \
  case MM:                                                                                         \
    return mode == KGE_HEAD_BATCH ? launch_bwd<MM, true, 4, false, true>(p, grid, smem, st)        \
                                  : launch_bwd<MM, false, 4, false, true>(p, grid, smem, st);
    switch (t->model) {
      KGE_CASE(KGE_TRANSE)
      KGE_CASE(KGE_DISTMULT)
      KGE_CASE(KGE_COMPLEX)
      KGE_CASE(KGE_ROTATE)
      KGE_CASE(KGE_PROTATE)
    }
#undef KGE_CASE
    return KGE_E_MODEL;
  }
  if (sh) {
#define KGE_CASE(MM)                                                                  \
  case MM:                                                                            \
    return mode == KGE_HEAD_BATCH ? launch_bwd<MM, true, 4, true>(p, grid, smem, st)  \
                                  : launch_bwd<MM, false, 4, true>(p, grid, smem, st);
    switch (t->model) {
      KGE_CASE(KGE_TRANSE)
      KGE_CASE(KGE_DISTMULT)
      KGE_CASE(KGE_COMPLEX)
      KGE_CASE(KGE_ROTATE)
      KGE_CASE(KGE_PROTATE)
    }
#undef KGE_CASE
    return KGE_E_MODEL;
  }
  // K3b: row gradients through shared-memory staging rows + one bulk reduction per row (see score_bwd_kernel)
  const bool bulk = vec && p.K > 0 && p.D <= kThreads * 4 && env_int("KGE_BWD_BULK", KGE_BWD_BULK_DEFAULT) != 0;
  if (bulk) {
    const size_t smem_b = smem + (size_t)G * kBulkSlots * entity_comps(t->model) * p.D * sizeof(float);
#define KGE_CASE(MM)                                                                                       \
  case MM:                                                                                                 \
    return mode == KGE_HEAD_BATCH ? launch_bwd<MM, true, 4, false, false, true>(p, grid, smem_b, st)       \
                                  : launch_bwd<MM, false, 4, false, false, true>(p, grid, smem_b, st);
    switch (t->model) {
      KGE_CASE(KGE_TRANSE)
      KGE_CASE(KGE_DISTMULT)
      KGE_CASE(KGE_COMPLEX)
      KGE_CASE(KGE_ROTATE)
      KGE_CASE(KGE_PROTATE)
    }
#undef KGE_CASE
    return KGE_E_MODEL;
  }
#define KGE_CASE(MM)                                                                       \
  case MM:                                                                                 \
    if (mode == KGE_HEAD_BATCH)                                                            \
      return vec ? launch_bwd<MM, true, 4>(p, grid, smem, st) : launch_bwd<MM, true, 1>(p, grid, smem, st); \
    return vec ? launch_bwd<MM, false, 4>(p, grid, smem, st) : launch_bwd<MM, false, 1>(p, grid, smem, st);
  switch (t->model) {
    KGE_CASE(KGE_TRANSE)
    KGE_CASE(KGE_DISTMULT)
    KGE_CASE(KGE_COMPLEX)
    KGE_CASE(KGE_ROTATE)
    KGE_CASE(KGE_PROTATE)
  }
#undef KGE_CASE
  return KGE_E_MODEL;
}

// pass 1 of the by-entity backward (byent.cu): dq per positive; the gradient rows of its head / tail /
// relation are STORED into gh_buf / gt_buf [B, NC*D] and gr_buf [B, RC*D] (no atomics anywhere)
int dq_pass_launch(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                   const float* coef_pos, const float* coef_neg, const float* stats, const float* grad_loss,
                   float* gh_buf, float* gt_buf, float* gr_buf, cudaStream_t st) {
  return run_bwd(t, mode, sample, B, neg, K, coef_pos, coef_neg, stats, grad_loss, /*grad_ent=*/gh_buf,
                 /*grad_rel=*/gr_buf, st, 0, 0, 1, 0, nullptr, gh_buf, gt_buf);
}

}  // namespace kge

using namespace kge;

extern "C" int kge_score_fwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                             const int64_t* neg, int64_t K, float* scores, kge_stream_t stream) {
  int rc = validate_tables(t);
  if (rc) return rc;
  if (!sample || !scores) return KGE_E_NULL;
  if (B < 0 || B > INT32_MAX || (neg && (K <= 0 || K > INT32_MAX))) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B == 0) return KGE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FwdParams p{};
  fill_fwd(p, t);
  p.sample = sample;
  p.B = (int)B;
  const bool vec = can_vectorize(t);
  if (!neg) {
    p.pos_score = scores;
    p.K = 0;
    const unsigned grid = (unsigned)((B + kWarps - 1) / kWarps);
#define KGE_CASE(MM)                                                        \
  case MM:                                                                  \
    if (vec) score_pos_kernel<MM, 4><<<grid, kThreads, 0, st>>>(p);         \
    else score_pos_kernel<MM, 1><<<grid, kThreads, 0, st>>>(p);             \
    break;
    switch (t->model) {
      KGE_CASE(KGE_TRANSE)
      KGE_CASE(KGE_DISTMULT)
      KGE_CASE(KGE_COMPLEX)
      KGE_CASE(KGE_ROTATE)
      KGE_CASE(KGE_PROTATE)
    }
#undef KGE_CASE
    KGE_LAUNCH_CHECK();
    return KGE_OK;
  }
  p.neg = neg;
  p.neg_score = scores;
  p.K = (int)K;
  const int want = (4 * sm_count() + (int)B - 1) / (int)B;
  const int maxks = (p.K + 63) / 64;
  int ks = want < 1 ? 1 : (want > maxks ? maxks : want);
  if (const char* e = getenv("KGE_KS")) ks = atoi(e) > 0 ? atoi(e) : ks;
  p.k_per_cta = (p.K + ks - 1) / ks;
  ks = (p.K + p.k_per_cta - 1) / p.k_per_cta;
  const size_t smem = (size_t)entity_comps(t->model) * p.Dp * sizeof(float);
  return dispatch_neg<false>(t->model, mode, vec, p, dim3((unsigned)B, (unsigned)ks), smem, st);
}

extern "C" int kge_score_bwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                             const int64_t* neg, int64_t K, const float* grad_scores, float* grad_entity,
                             float* grad_relation, kge_stream_t stream) {
  int rc = validate_tables(t);
  if (rc) return rc;
  if (!sample || !grad_scores || !grad_entity || !grad_relation) return KGE_E_NULL;
  if (B < 0 || B > INT32_MAX || (neg && (K <= 0 || K > INT32_MAX))) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B == 0) return KGE_OK;
  if (!neg)
    return run_bwd(t, KGE_TAIL_BATCH, sample, B, nullptr, 0, grad_scores, nullptr, nullptr, nullptr,
                   grad_entity, grad_relation, (cudaStream_t)stream);
  return run_bwd(t, mode, sample, B, neg, K, nullptr, grad_scores, nullptr, nullptr, grad_entity,
                 grad_relation, (cudaStream_t)stream);
}

extern "C" int kge_fused_bwd_chunk(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                                   const int64_t* neg, int64_t K, const float* coef_pos, const float* coef_neg,
                                   const float* stats, const float* grad_loss, int32_t col0, int32_t ncols,
                                   int32_t n_records, int64_t record_stride_bytes, float* grad_entity_chunk,
                                   float* grad_relation_chunk, kge_stream_t stream) {
  int rc = validate_tables(t);
  if (rc) return rc;
  if (!sample || !neg || !coef_pos || !coef_neg || !stats || !grad_entity_chunk || !grad_relation_chunk)
    return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX || ncols <= 0) return KGE_E_SIZE;
  const int n_abs = n_records < 0 ? -n_records : n_records;
  if (n_abs < 1 || n_abs > 64 || (n_abs > 1 && (record_stride_bytes <= 0 || record_stride_bytes % 8)))
    return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  return run_bwd(t, mode, sample, B, neg, K, coef_pos, coef_neg, stats, grad_loss, grad_entity_chunk,
                 grad_relation_chunk, (cudaStream_t)stream, col0, ncols, n_records, record_stride_bytes);
}

extern "C" size_t kge_loss_workspace_bytes(int64_t B) {
  return (size_t)(3 * (B > 0 ? B : 0)) * sizeof(float) + 16;
}

#ifndef KGE_NO_TMA
// K2-TMA launcher (score_tma.cuh).  Returns KGE_E_UNSUPPORTED when the shape does not fit the variant
// (the caller then uses the LDG kernel).
template <int M, bool HEAD, int UMAX, int MINB>
static int launch_neg_tma(const FwdParams& p, int stages, size_t smem, int grid, int* fail, cudaStream_t st) {
  auto kern = score_neg_tma_kernel<M, HEAD, UMAX, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, kThreads, smem, st>>>(p, stages, fail);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

static int run_fwd_tma(const kge_tables_t* t, int mode, const FwdParams& p, int* fail, cudaStream_t st) {
  const int D = t->hidden_dim;
  if (!can_vectorize(t) || D > 1024 || t->model == KGE_PROTATE) return KGE_E_UNSUPPORTED;
  const int minb = env_int("KGE_TMA_MINB", 1) >= 2 ? 2 : 1;
  const size_t row_bytes = (size_t)p.ent_stride * 4;
  const size_t Kp = ((size_t)p.K + 3) & ~(size_t)3;
  const size_t fixed = ((size_t)entity_comps(t->model) * p.Dp + 2 * Kp) * 4;
  const size_t budget = (size_t)(minb == 1 ? 220 : 108) * 1024;
  if (fixed + 64 + kWarps * row_bytes > budget) return KGE_E_UNSUPPORTED;
  int stages = (int)((budget - fixed) / (kWarps * row_bytes + 64));
  if (stages > 8) stages = 8;
  const int want = env_int("KGE_TMA_STAGES", 0);
  if (want > 0 && want < stages) stages = want;
  if (stages < 1) return KGE_E_UNSUPPORTED;
  const size_t smem = fixed + (size_t)stages * (kWarps * row_bytes + 64);
  int grid = sm_count() * minb;
  if (grid > p.B) grid = p.B;
  const int umax = D <= 256 ? 2 : (D <= 512 ? 4 : 8);
#define KGE_TMA_CASE3(MM, HH, UU)                                                                      \
  return minb == 2 ? launch_neg_tma<MM, HH, UU, 2>(p, stages, smem, grid, fail, st)                    \
                   : launch_neg_tma<MM, HH, UU, 1>(p, stages, smem, grid, fail, st);
#define KGE_TMA_CASE2(MM, HH)                                                                          \
  if (umax == 2) { KGE_TMA_CASE3(MM, HH, 2) } else if (umax == 4) { KGE_TMA_CASE3(MM, HH, 4) } else { KGE_TMA_CASE3(MM, HH, 8) }
#define KGE_TMA_CASE(MM)                                                                               \
  case MM:                                                                                             \
    if (mode == KGE_HEAD_BATCH) { KGE_TMA_CASE2(MM, true) } else { KGE_TMA_CASE2(MM, false) }
  switch (t->model) {
    KGE_TMA_CASE(KGE_TRANSE)
    KGE_TMA_CASE(KGE_DISTMULT)
    KGE_TMA_CASE(KGE_COMPLEX)
    KGE_TMA_CASE(KGE_ROTATE)
  }
#undef KGE_TMA_CASE
#undef KGE_TMA_CASE2
#undef KGE_TMA_CASE3
  return KGE_E_MODEL;
}

#else
static int run_fwd_tma(const kge_tables_t*, int, const FwdParams&, int*, cudaStream_t) { return KGE_E_UNSUPPORTED; }
#endif

extern "C" int kge_fused_fwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                             const int64_t* neg, int64_t K, const float* weight, float alpha,
                             float* pos_score, float* neg_score, float* coef_pos, float* coef_neg,
                             float* stats, void* workspace, kge_stream_t stream) {
  int rc = validate_tables(t);
  if (rc) return rc;
  if (!sample || !neg || !weight || !coef_pos || !coef_neg || !stats || !workspace) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  FwdParams p{};
  fill_fwd(p, t);
  p.sample = sample;
  p.neg = neg;
  p.weight = weight;
  p.pos_score = pos_score;
  p.neg_score = neg_score;
  p.coef_pos = coef_pos;
  p.coef_neg = coef_neg;
  p.stats = stats;
  p.ticket = reinterpret_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 16);
  p.B = (int)B;
  p.K = (int)K;
  p.k_per_cta = (int)K;
  p.alpha = alpha;
  if (env_int("KGE_FWD_TMA", 0)) {  // A/B switch: rows staged through the TMA (score_tma.cuh)
    rc = run_fwd_tma(t, mode, p, reinterpret_cast<int*>(workspace) + 1, (cudaStream_t)stream);
    if (rc != KGE_E_UNSUPPORTED) return rc;
  }
  const size_t smem = ((size_t)entity_comps(t->model) * p.Dp + (size_t)K) * sizeof(float);
  if (smem > 200 * 1024) return KGE_E_UNSUPPORTED;
  return dispatch_neg<true>(t->model, mode, can_vectorize(t), p, dim3((unsigned)B, 1), smem,
                            (cudaStream_t)stream);
}

extern "C" int kge_tma_fail_flag(void) {
#ifndef KGE_NO_TMA
  int v = 0, zero = 0;
  if (cudaMemcpyFromSymbol(&v, g_tma_fail, sizeof(int)) != cudaSuccess) return -1;
  if (v) cudaMemcpyToSymbol(g_tma_fail, &zero, sizeof(int));
  return v;
#else
  return 0;
#endif
}

extern "C" int kge_fused_bwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                             const int64_t* neg, int64_t K, const float* coef_pos, const float* coef_neg,
                             const float* stats, const float* grad_loss, float* grad_entity,
                             float* grad_relation, kge_stream_t stream) {
  int rc = validate_tables(t);
  if (rc) return rc;
  if (!sample || !neg || !coef_pos || !coef_neg || !stats || !grad_entity || !grad_relation)
    return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  return run_bwd(t, mode, sample, B, neg, K, coef_pos, coef_neg, stats, grad_loss, grad_entity,
                 grad_relation, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// K7: row-sharded entity table
// ------------------------------------------------------------------------------------------------
extern "C" int kge_score_fwd_sharded(const kge_tables_t* t, const kge_shards_t* sh, int mode,
                                     const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                                     float* scores, kge_stream_t stream) {
  int rc = validate_sharded(t, sh, false);
  if (rc) return rc;
  if (!sample || !scores) return KGE_E_NULL;
  if (B < 0 || B > INT32_MAX || (neg && (K <= 0 || K > INT32_MAX))) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B == 0) return KGE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FwdParams p{};
  fill_fwd(p, t, sh);
  p.sample = sample;
  p.B = (int)B;
  if (!neg) {
    p.pos_score = scores;
    p.K = 0;
    const unsigned grid = (unsigned)((B + kWarps - 1) / kWarps);
    switch (t->model) {
      case KGE_TRANSE: score_pos_kernel<KGE_TRANSE, 4, true><<<grid, kThreads, 0, st>>>(p); break;
      case KGE_DISTMULT: score_pos_kernel<KGE_DISTMULT, 4, true><<<grid, kThreads, 0, st>>>(p); break;
      case KGE_COMPLEX: score_pos_kernel<KGE_COMPLEX, 4, true><<<grid, kThreads, 0, st>>>(p); break;
      case KGE_ROTATE: score_pos_kernel<KGE_ROTATE, 4, true><<<grid, kThreads, 0, st>>>(p); break;
      case KGE_PROTATE: score_pos_kernel<KGE_PROTATE, 4, true><<<grid, kThreads, 0, st>>>(p); break;
    }
    KGE_LAUNCH_CHECK();
    return KGE_OK;
  }
  p.neg = neg;
  p.neg_score = scores;
  p.K = (int)K;
  const int want = (4 * sm_count() + (int)B - 1) / (int)B;
  const int maxks = (p.K + 63) / 64;
  int ks = want < 1 ? 1 : (want > maxks ? maxks : want);
  p.k_per_cta = (p.K + ks - 1) / ks;
  ks = (p.K + p.k_per_cta - 1) / p.k_per_cta;
  const size_t smem = (size_t)entity_comps(t->model) * p.Dp * sizeof(float);
  return dispatch_neg_sharded<false>(t->model, mode, p, dim3((unsigned)B, (unsigned)ks), smem, st);
}

extern "C" int kge_fused_fwd_sharded(const kge_tables_t* t, const kge_shards_t* sh, int mode,
                                     const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                                     const float* weight, float alpha, float* pos_score, float* neg_score,
                                     float* coef_pos, float* coef_neg, float* stats, void* workspace,
                                     kge_stream_t stream) {
  int rc = validate_sharded(t, sh, false);
  if (rc) return rc;
  if (!sample || !neg || !weight || !coef_pos || !coef_neg || !stats || !workspace) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  FwdParams p{};
  fill_fwd(p, t, sh);
  p.sample = sample;
  p.neg = neg;
  p.weight = weight;
  p.pos_score = pos_score;
  p.neg_score = neg_score;
  p.coef_pos = coef_pos;
  p.coef_neg = coef_neg;
  p.stats = stats;
  p.ticket = reinterpret_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 16);
  p.B = (int)B;
  p.K = (int)K;
  p.k_per_cta = (int)K;
  p.alpha = alpha;
  const size_t smem = ((size_t)entity_comps(t->model) * p.Dp + (size_t)K) * sizeof(float);
  if (smem > 200 * 1024) return KGE_E_UNSUPPORTED;
  return dispatch_neg_sharded<true>(t->model, mode, p, dim3((unsigned)B, 1), smem, (cudaStream_t)stream);
}

extern "C" int kge_fused_bwd_sharded(const kge_tables_t* t, const kge_shards_t* sh, int mode,
                                     const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                                     const float* coef_pos, const float* coef_neg, const float* stats,
                                     const float* grad_loss, float* grad_relation, kge_stream_t stream) {
  int rc = validate_sharded(t, sh, true);
  if (rc) return rc;
  if (!sample || !neg || !coef_pos || !coef_neg || !stats || !grad_relation) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  return run_bwd(t, mode, sample, B, neg, K, coef_pos, coef_neg, stats, grad_loss, nullptr, grad_relation,
                 (cudaStream_t)stream, 0, 0, 1, 0, sh);
}

// ------------------------------------------------------------------------------------------------
// pRotatE's trainable modulus (mkb/models/protate.py:72,91): score = gamma - modulus * A with
// A = sum_d |sin(phase)|, so dL/dmodulus = sum_k dL/dscore_k * (-A_k) = sum_k g_k (score_k - gamma) / modulus.
// One CTA, fixed summation order (deterministic); ADDS into grad_modulus[0].
// ------------------------------------------------------------------------------------------------
namespace kge {
__global__ void __launch_bounds__(kThreads) modulus_grad_kernel(const float* __restrict__ scores,
                                                                const float* __restrict__ grad_scores, int64_t n,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ grad_loss, float gamma,
                                                                const float* __restrict__ modulus,
                                                                float* grad_modulus) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int64_t k = threadIdx.x; k < n; k += kThreads) acc = fmaf(grad_scores[k], scores[k] - gamma, acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float scale = stats ? (grad_loss ? __ldg(grad_loss) : 1.f) / (2.f * __ldg(stats + 2)) : 1.f;
    grad_modulus[0] += scale * acc / __ldg(modulus);
  }
}
}  // namespace kge

extern "C" int kge_modulus_grad(const float* scores, const float* grad_scores, int64_t n, const float* stats,
                                const float* grad_loss, float gamma, const float* modulus, float* grad_modulus,
                                kge_stream_t stream) {
  if (!scores || !grad_scores || !modulus || !grad_modulus) return KGE_E_NULL;
  if (n < 0) return KGE_E_SIZE;
  if (n == 0) return KGE_OK;
  modulus_grad_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(scores, grad_scores, n, stats, grad_loss, gamma,
                                                               modulus, grad_modulus);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
