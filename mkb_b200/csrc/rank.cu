// K5: filtered all-entity ranking.
//
// Replaces, for head-/tail-batch link prediction, TestDataset.__getitem__'s per-query Python loop
// over all entities (mkb/datasets/base.py:196-241) and Evaluation.compute_score's
// score -> += filter_bias -> argsort -> position-of-the-positive (mkb/evaluation/evaluation.py:237-263).
//
// The rank of the positive needs no sort:
//     rank = 1 + #{e unfiltered : s_e > s_pos} + #{e < pos unfiltered : s_e == s_pos}
// (what a stable descending argsort yields; filtered candidates carry s_pos - 1e5 in the reference
// and can never outrank the positive).  So the kernel is a distance-"GEMM": a 64-query x 64-entity
// tile per CTA, 4x4 scores per thread in registers, both operand tiles staged through shared
// memory, hidden dim streamed in chunks; every score is accumulated sequentially over d with the
// same instruction sequence, and the positive's own score comes from the same routine, so the
// comparison s_e > s_pos is exact and self-consistent.
//
// Bound: RotatE is MUFU-bound (one sqrt per query x entity x dim), TransE FP32-bound, the two
// dot-product models FP32-FMA-bound in this fp32 formulation (tensor-core split-bf16 is a later
// row); none is HBM-bound: the entity table is re-read once per 64 queries out of L2.
#include <cstdlib>

#include "kge_common.cuh"

namespace kge {

constexpr int kTQ = 64, kTE = 64, kDK = 32, kPad = 68;  // tile sizes; padded row stride (floats)

// acc <- acc (+) term, one fixed instruction sequence shared by every caller
template <int M>
__device__ __forceinline__ float cand_acc(float acc, float q0, float q1, float e0, float e1, float pd) {
  if constexpr (M == KGE_PROTATE) {
    return __fadd_rn(acc, fabsf(sinf(__fsub_rn(q0, __fdiv_rn(e0, pd)))));
  } else if constexpr (M == KGE_TRANSE) {
    return __fadd_rn(acc, fabsf(__fsub_rn(e0, q0)));
  } else if constexpr (M == KGE_DISTMULT) {
    return __fmaf_rn(q0, e0, acc);
  } else if constexpr (M == KGE_COMPLEX) {
    return __fmaf_rn(q1, e1, __fmaf_rn(q0, e0, acc));
  } else {
    const float dx = __fsub_rn(q0, e0), dy = __fsub_rn(q1, e1);
    return __fadd_rn(acc, sqrt_approx(__fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
  }
}

struct RankParams {
  const float* ent;
  const float* rel;
  const int64_t* queries;
  kge_filter_csr_t filter;
  int has_filter;
  float* qmat;        // [Q][NC*D]
  float* pos_score;   // [Q]
  float* posrows;     // [Q][NC*D] copy of every query's positive entity row (tensor-core path only, else null)
  int64_t* seg;       // [Q][2]
  unsigned long long* ranks;
  float* scores_out;  // optional [Q][N]
  int64_t N;
  int Q, D;
  int ent_stride, rel_stride;
  float gamma, phase_div;
  const float* modulus;  // pRotatE: device scalar (else null)
  // K7 (row-sharded table): `ent` / `N` describe ONE shard; local row l is entity l * id_mul + id_add.
  // The two rows a query itself needs (fixed side, positive) may live on other shards: `shard`.
  int64_t id_mul, id_add, n_global;
  const float* shard[KGE_MAX_SHARDS];
  unsigned n_shards;  // 0: unsharded
};
static_assert(sizeof(RankParams) <= 4096, "kernel parameters are passed by value: 4 KB limit");

__device__ __forceinline__ const float* rk_entity_row(const RankParams& p, int64_t id) {
  if (p.n_shards == 0) return p.ent + id * (int64_t)p.ent_stride;
  const unsigned u = (unsigned)id, q = u / p.n_shards;
  return p.shard[u - q * p.n_shards] + (int64_t)q * p.ent_stride;
}

template <int M>
__device__ __forceinline__ float rk_modulus(const RankParams& p) {
  if constexpr (Traits<M>::kPhase) return __ldg(p.modulus);
  else return 1.f;
}

__device__ __forceinline__ int64_t rk_find_key(const int64_t* __restrict__ keys, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return (lo < n && __ldg(keys + lo) == key) ? lo : -1;
}
__device__ __forceinline__ bool rk_member(const int64_t* __restrict__ m, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = __ldg(m + mid);
    if (v < x) lo = mid + 1;
    else if (v > x) hi = mid;
    else return true;
  }
  return false;
}

// 1. per query: the query vector, the positive's score, the filter segment, rank := 1
template <int M, bool HEAD>
__global__ void __launch_bounds__(kThreads) rank_prepare_kernel(RankParams p) {
  using T = Traits<M>;
  const int qi = blockIdx.x;
  const int64_t h = p.queries[3 * qi], r = p.queries[3 * qi + 1], t = p.queries[3 * qi + 2];
  const float* fixed = rk_entity_row(p, HEAD ? t : h);
  const float* relrow = p.rel + r * (int64_t)p.rel_stride;
  float* q = p.qmat + (int64_t)qi * p.ent_stride;
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    const float a0 = fixed[d], a1 = T::NC == 2 ? fixed[p.D + d] : 0.f;
    float r0, r1, q0, q1;
    rel_effective<M>(relrow[d], T::RC == 2 ? relrow[p.D + d] : 0.f, p.phase_div, r0, r1);
    make_query<M, HEAD>(a0, a1, r0, r1, q0, q1, p.phase_div);
    q[d] = q0;
    if constexpr (T::NC == 2) q[p.D + d] = q1;
  }
  if (p.posrows) {  // the tensor-core path re-scores the positive through the SAME GEMM as the candidates
    const float* e = rk_entity_row(p, HEAD ? h : t);
    float* dst = p.posrows + (int64_t)qi * p.ent_stride;
    for (int d = threadIdx.x; d < p.ent_stride; d += blockDim.x) dst[d] = e[d];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float* e = rk_entity_row(p, HEAD ? h : t);
    float acc = 0.f;
    for (int d = 0; d < p.D; ++d)
      acc = cand_acc<M>(acc, q[d], T::NC == 2 ? q[p.D + d] : 0.f, e[d], T::NC == 2 ? e[p.D + d] : 0.f, p.phase_div);
    p.pos_score[qi] = finish_score<M>(acc, p.gamma, rk_modulus<M>(p));
    p.ranks[qi] = p.n_shards ? 0ull : 1ull;  // sharded: this shard's COUNT, the caller sums and adds 1
    int64_t lo = 0, hi = 0;
    if (p.has_filter) {
      const int64_t k = rk_find_key(p.filter.keys, p.filter.n_keys, r * p.n_global + (HEAD ? t : h));
      if (k >= 0) {
        lo = p.filter.offsets[k];
        hi = p.filter.offsets[k + 1];
      }
    }
    p.seg[2 * qi] = lo;
    p.seg[2 * qi + 1] = hi;
  }
}

// 2. 64 x 64 tile of scores, compare + count
template <int M, bool HEAD>
__global__ void __launch_bounds__(kThreads) rank_tile_kernel(RankParams p) {
  using T = Traits<M>;
  __shared__ __align__(16) float qs[T::NC][kDK][kPad];
  __shared__ __align__(16) float es[T::NC][kDK][kPad];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int q_base = blockIdx.x * kTQ;
  // entity tiles are folded over gridDim.y x gridDim.z (grid.y alone caps at 65 535 tiles = 4.19 M entities)
  const int64_t e_base = ((int64_t)blockIdx.y + (int64_t)blockIdx.z * gridDim.y) * kTE;
  if (e_base >= p.N) return;  // the whole CTA: the folded grid may overshoot

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int d0 = 0; d0 < p.D; d0 += kDK) {
    __syncthreads();
    const int d = d0 + lane;
    for (int row = warp; row < kTQ; row += kWarps) {
      const int qi = q_base + row;
      const int64_t ei = e_base + row;
      const bool vq = qi < p.Q && d < p.D, ve = ei < p.N && d < p.D;
      const float* qrow = p.qmat + (int64_t)qi * p.ent_stride;
      const float* erow = p.ent + ei * (int64_t)p.ent_stride;
#pragma unroll
      for (int c = 0; c < T::NC; ++c) {
        qs[c][lane][row] = vq ? qrow[c * p.D + d] : 0.f;
        es[c][lane][row] = ve ? __ldg(erow + c * p.D + d) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int dd = 0; dd < kDK; ++dd) {
      const float4 qa = *reinterpret_cast<const float4*>(&qs[0][dd][ty * 4]);
      const float4 ea = *reinterpret_cast<const float4*>(&es[0][dd][tx * 4]);
      float4 qb = make_float4(0.f, 0.f, 0.f, 0.f), eb = qb;
      if constexpr (T::NC == 2) {
        qb = *reinterpret_cast<const float4*>(&qs[1][dd][ty * 4]);
        eb = *reinterpret_cast<const float4*>(&es[1][dd][tx * 4]);
      }
      const float q0[4] = {qa.x, qa.y, qa.z, qa.w}, q1[4] = {qb.x, qb.y, qb.z, qb.w};
      const float e0[4] = {ea.x, ea.y, ea.z, ea.w}, e1[4] = {eb.x, eb.y, eb.z, eb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = cand_acc<M>(acc[a][b], q0[a], q1[a], e0[b], e1[b], p.phase_div);
    }
  }

#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int qi = q_base + ty * 4 + a;
    const bool vq = qi < p.Q;  // uniform over the 16 threads that share ty
    unsigned cnt = 0;
    // Which of this tile's 64 entities are filtered for query qi: ONE thread of the half-warp that shares the query
    // walks the query's (sorted) filter segment across the tile's id span and broadcasts a 64-bit mask — one
    // lower_bound per query and tile instead of a binary search per beating candidate (with untrained tables half
    // of all candidates beat the positive).  The score-dump path keeps the per-element lookup.
    unsigned long long fmask = 0ull;
    if (!p.scores_out) {
      unsigned m_lo = 0u, m_hi = 0u;
      if (vq && tx == 0) {
        const int64_t lo = p.seg[2 * qi], hi = p.seg[2 * qi + 1];
        if (hi > lo) {
          const int64_t last_row = (e_base + kTE - 1 < p.N ? e_base + kTE - 1 : p.N - 1);
          const int64_t first = e_base * p.id_mul + p.id_add, last = last_row * p.id_mul + p.id_add;
          int64_t a0 = lo, b0 = hi;
          while (a0 < b0) {
            const int64_t mid = (a0 + b0) >> 1;
            if (__ldg(p.filter.members + mid) < first) a0 = mid + 1;
            else b0 = mid;
          }
          for (int64_t k = a0; k < hi; ++k) {
            const int64_t m = __ldg(p.filter.members + k);
            if (m > last) break;
            const int64_t r = m - p.id_add;
            if (r % p.id_mul == 0) fmask |= 1ull << (int)(r / p.id_mul - e_base);
          }
        }
        m_lo = (unsigned)fmask;
        m_hi = (unsigned)(fmask >> 32);
      }
      m_lo = __shfl_sync(kFull, m_lo, lane & 16);  // lane 0 / 16 = tx 0 of this half-warp
      m_hi = __shfl_sync(kFull, m_hi, lane & 16);
      fmask = ((unsigned long long)m_hi << 32) | m_lo;
    }
    if (vq) {
      const float sp = p.pos_score[qi];
      const int64_t pos = p.queries[3 * qi + (HEAD ? 0 : 2)];
      const int64_t lo = p.seg[2 * qi], hi = p.seg[2 * qi + 1];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t el = e_base + tx * 4 + b;  // row of this shard (= entity id when unsharded)
        if (el < p.N) {
          const int64_t e = el * p.id_mul + p.id_add;
          const float s = finish_score<M>(acc[a][b], p.gamma, rk_modulus<M>(p));
          const bool beats = (s > sp) || (s == sp && e < pos);
          bool filtered = (fmask >> (tx * 4 + b)) & 1ull;
          if (p.scores_out && e != pos && hi > lo) filtered = rk_member(p.filter.members, lo, hi, e);
          if (beats && e != pos && !filtered) ++cnt;
          if (p.scores_out) p.scores_out[(int64_t)qi * p.n_global + e] = filtered ? sp + (-1e5f) : s;
        }
      }
    }
    // the 16 threads with the same ty sit in one half-warp: fold before the atomic
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if (vq && tx == 0 && cnt) atomicAdd(p.ranks + qi, (unsigned long long)cnt);
  }
}

// rank_tc.cu: tcgen05 (3xTF32) GEMM + fused rank count for the dot-product models
// posrows (optional): [Q, kd] copy of every query's positive row; when given, a first small launch re-scores the
// positives through the same tensor-core arithmetic and overwrites pos_score, so "s == s_pos" means what it means
// in the reference (positive and candidates scored by one instruction sequence).
int rank_tc_launch(const float* qmat, const float* ent, int64_t n_entity, int kd, const int64_t* queries, int Q,
                   const kge_filter_csr_t* filter, bool has_filter, float* pos_score, const int64_t* seg,
                   unsigned long long* ranks, float* scores_out, bool head, cudaStream_t st,
                   const float* posrows = nullptr, int k_split = 1);
bool rank_tc_eligible(const float* ent, int kd, int64_t n_entity);

}  // namespace kge

using namespace kge;

extern "C" size_t kge_rank_workspace_bytes(const kge_tables_t* t, int64_t Q) {
  if (!t || Q <= 0) return 0;
  const size_t row = (size_t)t->hidden_dim * entity_comps(t->model);
  const size_t pos = (t->model == KGE_COMPLEX || t->model == KGE_DISTMULT) ? row * sizeof(float) : 0;  // posrows
  return (size_t)Q * (row * sizeof(float) + sizeof(float) + 2 * sizeof(int64_t) + pos) + 64;
}

// (query tiles, entity tiles) -> grid: entity tiles beyond grid.y's 65 535 spill into grid.z
static dim3 fold_tiles(int64_t q_tiles, int64_t e_tiles) {
  const int64_t ymax = 32768;
  const int64_t y = e_tiles < ymax ? (e_tiles > 0 ? e_tiles : 1) : ymax;
  return dim3((unsigned)q_tiles, (unsigned)y, (unsigned)((e_tiles + y - 1) / y > 0 ? (e_tiles + y - 1) / y : 1));
}

static int run_rank(const kge_tables_t* t, const kge_shards_t* sh, int shard_index, int mode, const int64_t* queries,
                    int64_t Q, const kge_filter_csr_t* filter, int64_t* ranks, float* scores_out, void* workspace,
                    kge_stream_t stream) {
  if (!t || !t->relation || !queries || !ranks || !workspace) return KGE_E_NULL;
  if (!sh && !t->entity) return KGE_E_NULL;
  if (t->model < KGE_TRANSE || t->model > KGE_PROTATE) return KGE_E_MODEL;
  if (t->model == KGE_PROTATE && !t->modulus) return KGE_E_NULL;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (Q < 0 || Q > INT32_MAX / kTQ || t->hidden_dim <= 0 || t->n_entity <= 0) return KGE_E_SIZE;
  RankParams p{};
  p.N = t->n_entity;
  p.n_global = t->n_entity;
  p.id_mul = 1;
  p.id_add = 0;
  p.ent = t->entity;
  if (sh) {
    if (sh->n_shards < 1 || sh->n_shards > KGE_MAX_SHARDS || shard_index < 0 || shard_index >= sh->n_shards ||
        t->n_entity > INT32_MAX)
      return KGE_E_SIZE;
    for (int s = 0; s < sh->n_shards; ++s) {
      if (!sh->entity[s]) return KGE_E_NULL;
      p.shard[s] = sh->entity[s];
    }
    p.n_shards = (unsigned)sh->n_shards;
    p.id_mul = sh->n_shards;
    p.id_add = shard_index;
    p.ent = sh->entity[shard_index];
    p.N = (t->n_entity - shard_index + sh->n_shards - 1) / sh->n_shards;  // rows of this shard
  }
  if (Q == 0) return KGE_OK;
  p.rel = t->relation;
  p.queries = queries;
  p.has_filter = filter && filter->n_keys > 0;
  if (p.has_filter) p.filter = *filter;
  p.Q = (int)Q;
  p.D = t->hidden_dim;
  p.ent_stride = t->hidden_dim * entity_comps(t->model);
  p.rel_stride = t->hidden_dim * relation_comps(t->model);
  p.gamma = t->gamma;
  p.phase_div = host_phase_div(t->embedding_range);
  p.modulus = t->modulus;
  char* ws = reinterpret_cast<char*>(workspace);
  p.seg = reinterpret_cast<int64_t*>(ws);
  ws += (size_t)Q * 2 * sizeof(int64_t);
  p.qmat = reinterpret_cast<float*>(ws);
  ws += (size_t)Q * p.ent_stride * sizeof(float);
  p.pos_score = reinterpret_cast<float*>(ws);
  ws += ((size_t)Q * sizeof(float) + 15) & ~(size_t)15;
  p.ranks = reinterpret_cast<unsigned long long*>(ranks);
  p.scores_out = scores_out;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid = fold_tiles((Q + kTQ - 1) / kTQ, (p.N + kTE - 1) / kTE);
  if (grid.z > 65535) return KGE_E_SIZE;
  // dot-product models: GEMM on the tensor cores when the shape allows (otherwise the fp32 tiles);
  // the sharded variant keeps to the fp32 tiles (the tcgen05 kernel assumes row == entity id)
  const bool dot_model = !sh && (t->model == KGE_COMPLEX || t->model == KGE_DISTMULT);
  if (dot_model && rank_tc_eligible(p.ent, p.ent_stride, p.N)) p.posrows = reinterpret_cast<float*>(ws);
#define KGE_CASE(MM)                                                                              \
  case MM:                                                                                        \
    if (mode == KGE_HEAD_BATCH) rank_prepare_kernel<MM, true><<<(unsigned)Q, kThreads, 0, st>>>(p); \
    else rank_prepare_kernel<MM, false><<<(unsigned)Q, kThreads, 0, st>>>(p);                     \
    if (p.N == 0) break; /* a shard without rows (more shards than entities) */                   \
    if (dot_model &&                                                                              \
        rank_tc_launch(p.qmat, p.ent, p.N, p.ent_stride, queries, p.Q, filter, p.has_filter != 0, \
                       p.pos_score, p.seg, p.ranks, scores_out, mode == KGE_HEAD_BATCH, st, p.posrows) == KGE_OK) \
      break;                                                                                      \
    if (mode == KGE_HEAD_BATCH) rank_tile_kernel<MM, true><<<grid, kThreads, 0, st>>>(p);         \
    else rank_tile_kernel<MM, false><<<grid, kThreads, 0, st>>>(p);                               \
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

// out[M, N] = a[M, Kd] · b[N, Kd]^T in fp32-grade arithmetic: the ranking GEMM used as a plain "NT" GEMM
// (pooled.cu).  It IS the DistMult score matrix of "queries" a against "entities" b, so it runs on the
// tcgen05 3xTF32 kernel when the shape allows and on the fp32 tile kernel otherwise; the rank-counting
// epilogue works on zeroed dummies.  scratch: dot_nt_scratch_bytes(M) bytes.
namespace kge {
size_t dot_nt_scratch_bytes(int64_t M) { return (size_t)M * (3 * 8 + 2 * 8 + 8 + 4) + 64; }

int dot_nt_launch(const float* a, const float* b, int64_t M, int64_t N, int Kd, float* out, void* scratch,
                  cudaStream_t st) {
  if (M <= 0 || N <= 0 || Kd <= 0 || M > INT32_MAX / kTQ || N > INT32_MAX) return KGE_E_SIZE;
  // "positive" id -1 for every row: no column is ever treated as the positive (the tcgen05 epilogue would
  // substitute the prepared positive score there); everything else the rank epilogue reads is zero
  cudaError_t ce = cudaMemsetAsync(scratch, 0xFF, (size_t)M * 3 * 8, st);
  if (ce == cudaSuccess)
    ce = cudaMemsetAsync(reinterpret_cast<char*>(scratch) + (size_t)M * 3 * 8, 0,
                         dot_nt_scratch_bytes(M) - (size_t)M * 3 * 8, st);
  if (ce != cudaSuccess) return (int)ce;
  RankParams p{};
  char* ws = reinterpret_cast<char*>(scratch);
  p.queries = reinterpret_cast<const int64_t*>(ws);
  ws += (size_t)M * 3 * 8;
  p.seg = reinterpret_cast<int64_t*>(ws);
  ws += (size_t)M * 2 * 8;
  p.ranks = reinterpret_cast<unsigned long long*>(ws);
  ws += (size_t)M * 8;
  p.pos_score = reinterpret_cast<float*>(ws);
  p.qmat = const_cast<float*>(a);
  p.ent = b;
  p.N = N;
  p.n_global = N;
  p.id_mul = 1;
  p.Q = (int)M;
  p.D = Kd;
  p.ent_stride = Kd;
  p.scores_out = out;
  // small M x N (the pooled flow's 1024 x 512 x 2000 GEMMs are 16-64 tiles for 148 SMs): split K over grid.z so
  // that ~one wave of CTAs is busy; partial sums are added into `out` (zeroed here).  KGE_DOT_SPLITK=0: off.
  int k_split = 1;
  if (rank_tc_eligible(b, Kd, N) && aligned16(a)) {
    const int64_t tiles = ((M + 127) / 128) * ((N + 255) / 256);
    const char* env = getenv("KGE_DOT_SPLITK");
    if (tiles < 96 && !(env && atoi(env) == 0)) {
      k_split = (int)(148 / tiles);
      if (k_split > (Kd + 31) / 32) k_split = (Kd + 31) / 32;
      if (k_split > 1 && cudaMemsetAsync(out, 0, (size_t)M * N * sizeof(float), st) != cudaSuccess) k_split = 1;
    }
  }
  if (rank_tc_launch(a, b, N, Kd, p.queries, p.Q, nullptr, false, p.pos_score, p.seg, p.ranks, out, false, st, nullptr,
                     k_split) == KGE_OK)
    return KGE_OK;
  dim3 grid = fold_tiles((M + kTQ - 1) / kTQ, (N + kTE - 1) / kTE);
  if (grid.z > 65535) return KGE_E_SIZE;
  rank_tile_kernel<KGE_DISTMULT, false><<<grid, kThreads, 0, st>>>(p);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
}  // namespace kge

extern "C" int kge_rank_all(const kge_tables_t* t, int mode, const int64_t* queries, int64_t Q,
                            const kge_filter_csr_t* filter, int64_t* ranks, float* scores_out,
                            void* workspace, kge_stream_t stream) {
  return run_rank(t, nullptr, 0, mode, queries, Q, filter, ranks, scores_out, workspace, stream);
}

extern "C" int kge_rank_counts_sharded(const kge_tables_t* t, const kge_shards_t* shards, int32_t shard_index,
                                       int mode, const int64_t* queries, int64_t Q, const kge_filter_csr_t* filter,
                                       int64_t* counts, float* scores_out, void* workspace, kge_stream_t stream) {
  if (!shards) return KGE_E_NULL;
  return run_rank(t, shards, shard_index, mode, queries, Q, filter, counts, scores_out, workspace, stream);
}
