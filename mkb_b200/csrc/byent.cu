// By-entity backward with the optimizer fused in — an atomics-free alternative to K3 + dense Adam for the
// device-resident training step (replaces error.backward(); optimizer.step(); optimizer.zero_grad() at
// mkb/compose/pipeline.py:236-240 for the entity table; the relation table keeps K3's path).
//
// K3 scatters B*K candidate-row gradients into the dense gradient table with vector REDs (ncu: bound by
// the L1/TEX + L2 atomic path), after which the dense Adam pass streams table, gradient and both moments
// through HBM once more.  Here the same arithmetic is regrouped so that every table row has ONE owner:
//
//   pass 0  build_query_kernel       Q[i]   = query vector of positive i (h∘r, h+r, ...)        8 MB at cfg 2
//           entry lists by entity    a per-step CSR  entity -> {candidate pairs (i,j), head/tail roles}
//                                    (histogram, single-CTA scan, scatter; sorted inside pass 2)
//   pass 1  score_bwd_kernel<DQONLY> per positive: dq over its K candidates (row reads only, no REDs),
//                                    chain rule -> gradient rows of its own head / tail into gh[i], gt[i]
//   pass 2  byent_apply_kernel       per ENTITY e: g_e = sum over its pairs c_ij * ds/de(Q[i], e)
//                                    + sum of the gh / gt rows of the positives it heads / tails,
//                                    accumulated in registers in a fixed order (deterministic), then the
//                                    Adam update of row e in place: the entity gradient never exists in
//                                    memory, needs no zeroing, and table + moments cross HBM once.
//           byrel_apply_kernel       per RELATION r: sum of the gr rows of its positives in index order,
//                                    Adam in place (<= a few hundred rows).
//
// Row e is read and written only by its own CTA in pass 2, so the in-place update is race-free.  No
// atomic touches a float anywhere in the step => the whole training step is bit-reproducible.
#include "kge_common.cuh"

namespace kge {

int dq_pass_launch(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B, const int64_t* neg, int64_t K,
                   const float* coef_pos, const float* coef_neg, const float* stats, const float* grad_loss,
                   float* gh_buf, float* gt_buf, float* gr_buf, cudaStream_t st);

constexpr int kMaxBucket = 2048;  // entries of one entity sorted in shared memory (larger buckets: unsorted)

struct ByEntParams {
  float* ent;  // updated in place
  const float* rel;
  const int64_t* sample;
  const int64_t* neg;
  const float* coef_neg;
  const float* stats;
  const float* grad_loss;
  const float* modulus;
  float* qmat;               // [B, NC*D]
  const float* gh_buf;       // [B, NC*D]
  const float* gt_buf;       // [B, NC*D]
  float* exp_avg;            // [N, NC*D]
  float* exp_avg_sq;         // [N, NC*D]
  float* rel_w;              // relation table, updated in place
  const float* gr_buf;       // [B, RC*D]
  float* rel_exp_avg;        // [R, RC*D]
  float* rel_exp_avg_sq;
  unsigned int* counts;      // [N + 1]  histogram, then scatter cursors
  unsigned int* offsets;     // [N + 1]
  unsigned int* entries;     // [B*K + 2B]: j-th candidate of positive i -> i*K + j; head role -> BK + i; tail role -> BK + B + i
  int64_t N;
  int B, K, D;
  int ent_stride, rel_stride;
  float phase_div;
  float lr_bc1, inv_sqrt_bc2, b1, b2, eps;
};
static_assert(sizeof(ByEntParams) <= 4096, "kernel parameters are passed by value: 4 KB limit");

template <int M, bool HEAD>
__global__ void __launch_bounds__(kThreads) build_query_kernel(ByEntParams p) {
  using T = Traits<M>;
  const int64_t i = blockIdx.x;
  const int64_t h = p.sample[3 * i], r = p.sample[3 * i + 1], t = p.sample[3 * i + 2];
  const float* fixed = p.ent + (HEAD ? t : h) * (int64_t)p.ent_stride;
  const float* relrow = p.rel + r * (int64_t)p.rel_stride;
  float* q = p.qmat + i * (int64_t)p.ent_stride;
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float r0, r1, q0, q1;
    rel_effective<M>(relrow[d], T::RC == 2 ? relrow[p.D + d] : 0.f, p.phase_div, r0, r1);
    make_query<M, HEAD>(fixed[d], T::NC == 2 ? fixed[p.D + d] : 0.f, r0, r1, q0, q1, p.phase_div);
    q[d] = q0;
    if constexpr (T::NC == 2) q[p.D + d] = q1;
  }
}

// ---- per-step CSR entity -> entries -------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) byent_count_kernel(ByEntParams p) {
  const int64_t BK = (int64_t)p.B * p.K, total = BK + 2 * (int64_t)p.B;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
    int64_t e;
    if (x < BK) e = p.neg[x];
    else if (x < BK + p.B) e = p.sample[3 * (x - BK)];          // head role
    else e = p.sample[3 * (x - BK - p.B) + 2];                  // tail role
    atomicAdd(p.counts + e, 1u);
  }
}

// exclusive scan of counts[0..N) into offsets[0..N], counts reset to 0 (they become the scatter cursors)
__global__ void __launch_bounds__(1024) byent_scan_kernel(ByEntParams p) {
  __shared__ unsigned int warp_tot[32];
  __shared__ unsigned int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < p.N; base += 1024) {
    const int64_t k = base + tid;
    const unsigned v = k < p.N ? p.counts[k] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned up = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = warp_tot[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(kFull, wi, o);
        if (lane >= o) wi += up;
      }
      warp_tot[lane] = wi - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const unsigned excl = carry + warp_tot[warp] + incl - v;
    if (k < p.N) {
      p.offsets[k] = excl;
      p.counts[k] = 0u;
    }
    __syncthreads();
    if (tid == 1023) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) p.offsets[p.N] = carry;
}

__global__ void __launch_bounds__(kThreads) byent_scatter_kernel(ByEntParams p) {
  const int64_t BK = (int64_t)p.B * p.K, total = BK + 2 * (int64_t)p.B;
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
    int64_t e;
    if (x < BK) e = p.neg[x];
    else if (x < BK + p.B) e = p.sample[3 * (x - BK)];
    else e = p.sample[3 * (x - BK - p.B) + 2];
    const unsigned slot = atomicAdd(p.counts + e, 1u);
    p.entries[p.offsets[e] + slot] = (unsigned)x;
  }
}

// ---- pass 2: one CTA per entity -----------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(kThreads, 4) byent_apply_kernel(ByEntParams p) {
  using T = Traits<M>;
  __shared__ unsigned int s_ent[kMaxBucket];
  const int tid = threadIdx.x, nthr = blockDim.x;  // 32..256 threads: one 16-byte chunk of the row each
  const int64_t e = blockIdx.x;
  const unsigned lo = p.offsets[e], hi = p.offsets[e + 1];
  const int n = (int)(hi - lo);
  const bool sorted = n <= kMaxBucket;
  if (n > 0 && n <= 32) {
    // the scatter's order depends on atomics: sort the bucket so the sums below have ONE order.
    // The common case (a few dozen entries) is a shuffle-only bitonic network in warp 0: one barrier.
    if (tid < 32) {
      unsigned v = tid < n ? __ldg(p.entries + lo + tid) : 0xFFFFFFFFu;
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          const unsigned other = __shfl_xor_sync(kFull, v, stride);
          const bool up = (tid & size) == 0, lower = (tid & stride) == 0;
          v = (lower == up) ? min(v, other) : max(v, other);
        }
      }
      s_ent[tid] = v;
    }
    __syncthreads();
  } else if (sorted && n > 0) {
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int k = tid; k < p2; k += nthr) s_ent[k] = k < n ? p.entries[lo + k] : 0xFFFFFFFFu;
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (p2 >> 1); t += nthr) {
          const int a_i = 2 * t - (t & (stride - 1)), b_i = a_i + stride;
          const bool up = (a_i & size) == 0;
          const unsigned a = s_ent[a_i], b = s_ent[b_i];
          if ((a > b) == up) {
            s_ent[a_i] = b;
            s_ent[b_i] = a;
          }
        }
        __syncthreads();
      }
    }
  }
  const unsigned BK = (unsigned)p.B * (unsigned)p.K;
  float scale = (p.grad_loss ? __ldg(p.grad_loss) : 1.f) / (2.f * __ldg(p.stats + 2));
  if constexpr (T::kPhase) scale *= __ldg(p.modulus);
  float* row = p.ent + e * (int64_t)p.ent_stride;
  float* mrow = p.exp_avg + e * (int64_t)p.ent_stride;
  float* vrow = p.exp_avg_sq + e * (int64_t)p.ent_stride;
  for (int d = tid * 4; d < p.D; d += nthr * 4) {
    float e0[4], e1[4] = {}, g0[4] = {}, g1[4] = {};
    ld_global<4>(row + d, e0);
    if constexpr (T::NC == 2) ld_global<4>(row + p.D + d, e1);
    // U entries at a time: resolve U source rows first (a pair reads Q[i], a role reads gh[i] / gt[i]), issue
    // all their 16-byte loads back to back, then do the arithmetic — the loop is latency-bound otherwise
    constexpr int U = 4;
    for (int k0 = 0; k0 < n; k0 += U) {
      const float* src[U];
      float c[U];
      bool pair[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = min(k0 + u, n - 1);  // the tail re-reads the last entry, its result is discarded
        const unsigned x = sorted ? s_ent[k] : __ldg(p.entries + lo + k);
        pair[u] = x < BK;
        if (pair[u]) {
          src[u] = p.qmat + (int64_t)(x / (unsigned)p.K) * p.ent_stride;
          c[u] = scale * __ldg(p.coef_neg + x);
        } else {
          const unsigned i = x - BK;
          src[u] = i < (unsigned)p.B ? p.gh_buf + (int64_t)i * p.ent_stride
                                     : p.gt_buf + (int64_t)(i - p.B) * p.ent_stride;
          c[u] = 0.f;
        }
      }
      float a0[U][4], a1[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ld_global<4>(src[u] + d, a0[u]);
        if constexpr (T::NC == 2) ld_global<4>(src[u] + p.D + d, a1[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u < n) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            if (pair[u]) {
              float x0, x1, dq0 = 0.f, dq1 = 0.f;
              cand_bwd<M>(a0[u][v], T::NC == 2 ? a1[u][v] : 0.f, e0[v], e1[v], c[u], x0, x1, dq0, dq1, p.phase_div);
              g0[v] += x0;
              g1[v] += x1;
            } else {
              g0[v] += a0[u][v];
              if constexpr (T::NC == 2) g1[v] += a1[u][v];
            }
          }
        }
      }
    }
    // torch.optim.Adam on this row (same arithmetic as adam_kernel, api.cu), table updated in place
#pragma unroll
    for (int c = 0; c < T::NC; ++c) {
      float* g = c == 0 ? g0 : g1;
      float* pe = c == 0 ? e0 : e1;
      float4 mm = *reinterpret_cast<float4*>(mrow + c * p.D + d);
      float4 vv = *reinterpret_cast<float4*>(vrow + c * p.D + d);
      float* mf = reinterpret_cast<float*>(&mm);
      float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        mf[v] = p.b1 * mf[v] + (1.f - p.b1) * g[v];
        vf[v] = p.b2 * vf[v] + (1.f - p.b2) * g[v] * g[v];
        const float denom = sqrtf(vf[v]) * p.inv_sqrt_bc2 + p.eps;
        pe[v] -= p.lr_bc1 * (mf[v] / denom);
      }
      *reinterpret_cast<float4*>(row + c * p.D + d) = make_float4(pe[0], pe[1], pe[2], pe[3]);
      *reinterpret_cast<float4*>(mrow + c * p.D + d) = mm;
      *reinterpret_cast<float4*>(vrow + c * p.D + d) = vv;
    }
  }
}

// relation table: one CTA per relation, the gr rows of its positives summed in index order, Adam in place
__global__ void __launch_bounds__(kThreads) byrel_apply_kernel(ByEntParams p) {
  const int64_t r = blockIdx.x;
  float* row = p.rel_w + r * (int64_t)p.rel_stride;
  float* mrow = p.rel_exp_avg + r * (int64_t)p.rel_stride;
  float* vrow = p.rel_exp_avg_sq + r * (int64_t)p.rel_stride;
  for (int d = threadIdx.x * 4; d < p.rel_stride; d += blockDim.x * 4) {
    float g[4] = {};
    for (int i = 0; i < p.B; ++i) {
      if (__ldg(p.sample + 3 * (int64_t)i + 1) != r) continue;  // uniform over the CTA
      float a[4];
      ld_global<4>(p.gr_buf + (int64_t)i * p.rel_stride + d, a);
#pragma unroll
      for (int v = 0; v < 4; ++v) g[v] += a[v];
    }
    float4 pp = *reinterpret_cast<float4*>(row + d);
    float4 mm = *reinterpret_cast<float4*>(mrow + d);
    float4 vv = *reinterpret_cast<float4*>(vrow + d);
    float* pf = reinterpret_cast<float*>(&pp);
    float* mf = reinterpret_cast<float*>(&mm);
    float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      mf[v] = p.b1 * mf[v] + (1.f - p.b1) * g[v];
      vf[v] = p.b2 * vf[v] + (1.f - p.b2) * g[v] * g[v];
      pf[v] -= p.lr_bc1 * (mf[v] / (sqrtf(vf[v]) * p.inv_sqrt_bc2 + p.eps));
    }
    *reinterpret_cast<float4*>(row + d) = pp;
    *reinterpret_cast<float4*>(mrow + d) = mm;
    *reinterpret_cast<float4*>(vrow + d) = vv;
  }
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace kge

using namespace kge;

extern "C" size_t kge_byent_workspace_bytes(const kge_tables_t* t, int64_t B, int64_t K) {
  if (!t || B <= 0 || K <= 0) return 0;
  const size_t row = (size_t)t->hidden_dim * entity_comps(t->model) * sizeof(float);
  const size_t rrow = (size_t)t->hidden_dim * relation_comps(t->model) * sizeof(float);
  return 3 * align_up((size_t)B * row) + align_up((size_t)B * rrow) + 2 * align_up((size_t)(t->n_entity + 1) * 4) +
         align_up((size_t)(B * K + 2 * B) * 4) + 256;
}

extern "C" int kge_bwd_by_entity_adam(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                                      const int64_t* neg, int64_t K, const float* coef_pos, const float* coef_neg,
                                      const float* stats, const float* grad_loss, float* entity_table,
                                      float* exp_avg, float* exp_avg_sq, float* relation_table, float* rel_exp_avg,
                                      float* rel_exp_avg_sq, int64_t step, float lr, float beta1, float beta2,
                                      float eps, void* workspace, kge_stream_t stream) {
  if (!t || !t->entity || !t->relation || !sample || !neg || !coef_pos || !coef_neg || !stats || !entity_table ||
      !exp_avg || !exp_avg_sq || !relation_table || !rel_exp_avg || !rel_exp_avg_sq || !workspace)
    return KGE_E_NULL;
  if (t->model < KGE_TRANSE || t->model > KGE_PROTATE) return KGE_E_MODEL;
  if (t->model == KGE_PROTATE && !t->modulus) return KGE_E_NULL;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B <= 0 || K <= 0 || step < 1 || t->n_entity <= 0 || t->hidden_dim <= 0) return KGE_E_SIZE;
  if ((B * K + 2 * B) > (int64_t)0xFFFFFFF0u || t->n_entity > INT32_MAX) return KGE_E_SIZE;
  // the update is in place on the tables that are read
  if (entity_table != t->entity || relation_table != t->relation) return KGE_E_UNSUPPORTED;
  if (t->hidden_dim % 4 != 0) return KGE_E_UNSUPPORTED;
  if (!aligned16(entity_table) || !aligned16(exp_avg) || !aligned16(exp_avg_sq) || !aligned16(relation_table) ||
      !aligned16(rel_exp_avg) || !aligned16(rel_exp_avg_sq) || !aligned16(workspace))
    return KGE_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  ByEntParams p{};
  p.ent = entity_table;
  p.rel = t->relation;
  p.sample = sample;
  p.neg = neg;
  p.coef_neg = coef_neg;
  p.stats = stats;
  p.grad_loss = grad_loss;
  p.modulus = t->modulus;
  p.exp_avg = exp_avg;
  p.exp_avg_sq = exp_avg_sq;
  p.rel_w = relation_table;
  p.rel_exp_avg = rel_exp_avg;
  p.rel_exp_avg_sq = rel_exp_avg_sq;
  p.N = t->n_entity;
  p.B = (int)B;
  p.K = (int)K;
  p.D = t->hidden_dim;
  p.ent_stride = t->hidden_dim * entity_comps(t->model);
  p.rel_stride = t->hidden_dim * relation_comps(t->model);
  p.phase_div = host_phase_div(t->embedding_range);
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  p.lr_bc1 = (float)((double)lr / bc1);
  p.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  p.b1 = beta1;
  p.b2 = beta2;
  p.eps = eps;
  const size_t row = (size_t)p.ent_stride * sizeof(float);
  char* ws = reinterpret_cast<char*>(workspace);
  p.qmat = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)B * row);
  float* gh = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)B * row);
  float* gt = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)B * row);
  p.gh_buf = gh;
  p.gt_buf = gt;
  float* gr = reinterpret_cast<float*>(ws);
  ws += align_up((size_t)B * p.rel_stride * sizeof(float));
  p.gr_buf = gr;
  p.counts = reinterpret_cast<unsigned int*>(ws);
  ws += align_up((size_t)(p.N + 1) * 4);
  p.offsets = reinterpret_cast<unsigned int*>(ws);
  ws += align_up((size_t)(p.N + 1) * 4);
  p.entries = reinterpret_cast<unsigned int*>(ws);

  cudaError_t ce = cudaMemsetAsync(p.counts, 0, (size_t)(p.N + 1) * 4, st);
  if (ce != cudaSuccess) return (int)ce;
  const int64_t total = B * K + 2 * B;
  unsigned blocks = (unsigned)((total + kThreads * 4 - 1) / (kThreads * 4));
  if (blocks > 4096) blocks = 4096;
  const bool head = mode == KGE_HEAD_BATCH;
  byent_count_kernel<<<blocks, kThreads, 0, st>>>(p);
  byent_scan_kernel<<<1, 1024, 0, st>>>(p);
  byent_scatter_kernel<<<blocks, kThreads, 0, st>>>(p);
#define KGE_CASE(MM)                                                              \
  case MM:                                                                        \
    if (head) build_query_kernel<MM, true><<<(unsigned)B, kThreads, 0, st>>>(p);  \
    else build_query_kernel<MM, false><<<(unsigned)B, kThreads, 0, st>>>(p);      \
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
  int rc = dq_pass_launch(t, mode, sample, B, neg, K, coef_pos, coef_neg, stats, grad_loss, gh, gt, gr, st);
  if (rc) return rc;
  // one thread per 16-byte chunk of a row component, 32..256 threads per entity
  int apply_threads = ((p.D / 4 + 31) / 32) * 32;
  apply_threads = apply_threads < 32 ? 32 : (apply_threads > kThreads ? kThreads : apply_threads);
#define KGE_CASE(MM) \
  case MM: byent_apply_kernel<MM><<<(unsigned)p.N, apply_threads, 0, st>>>(p); break;
  switch (t->model) {
    KGE_CASE(KGE_TRANSE)
    KGE_CASE(KGE_DISTMULT)
    KGE_CASE(KGE_COMPLEX)
    KGE_CASE(KGE_ROTATE)
    KGE_CASE(KGE_PROTATE)
  }
#undef KGE_CASE
  int rel_threads = ((p.rel_stride / 4 + 31) / 32) * 32;
  rel_threads = rel_threads < 32 ? 32 : (rel_threads > kThreads ? kThreads : rel_threads);
  byrel_apply_kernel<<<(unsigned)t->n_relation, rel_threads, 0, st>>>(p);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
