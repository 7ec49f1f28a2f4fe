// Pooled negatives on the tensor cores — the dot-product models' training step when the batch shares ONE
// candidate pool, which is exactly the reference's sampler (mkb/sampling/negative_sampling.py:166: a single
// randint(n_entity, 2*size) pool per generate() call, every positive takes its first `size` survivors).
//
// With a shared pool of P rows the (1+K) scores of all B positives are entries of ONE matrix
//     S[B, P] = Q[B, E] · Pool[P, E]^T          (E = D for DistMult, 2D for ComplEx: Re<h,r,conj(t)> is a
//                                                plain dot product of the query h∘r / conj(r)∘t with the row)
// so the reference's B·K row gathers (2.1 GB of logical traffic at config 3) collapse to P = 512 rows and
// three small GEMMs on the tcgen05 3xTF32 kernel (rank_tc.cu; fp32 tiles when the shape does not allow):
//     forward   S  = Q · Pool^T                   -> per positive: gather its K scores by pool POSITION,
//                                                    self-adversarial terms, dS row (coefficients scattered
//                                                    back to pool positions)
//     backward  dQ    = dS · Pool                 -> chain rule to the positive's fixed entity + relation
//               dPool = dS^T · Q                  -> P row gradients added into the entity gradient
// Replaces, for that sampler and DistMult / ComplEx, the same reference code as K2 / K3
// (mkb/compose/pipeline.py:211-236, distmult.py:63-75, complex.py:65-85, losses/adversarial.py:21-30).
#include "kge_common.cuh"

namespace kge {

size_t dot_nt_scratch_bytes(int64_t M);
int dot_nt_launch(const float* a, const float* b, int64_t M, int64_t N, int Kd, float* out, void* scratch,
                  cudaStream_t st);

static size_t pl_align(size_t x) { return (x + 255) & ~(size_t)255; }

// workspace carved identically by forward and backward
struct PooledWs {
  float *q, *qt, *pool, *poolt, *s, *dst, *dq, *dpool, *pos, *gpos;
  void* scratch;
  size_t bytes;
};
static PooledWs carve(void* base, int64_t B, int64_t P, int64_t E) {
  PooledWs w{};
  char* p = reinterpret_cast<char*>(base);
  auto take = [&](size_t n) {
    char* r = p;
    p += pl_align(n);
    return r;
  };
  w.q = reinterpret_cast<float*>(take((size_t)B * E * 4));
  w.qt = reinterpret_cast<float*>(take((size_t)B * E * 4));
  w.pool = reinterpret_cast<float*>(take((size_t)P * E * 4));
  w.poolt = reinterpret_cast<float*>(take((size_t)P * E * 4));
  w.s = reinterpret_cast<float*>(take((size_t)B * P * 4));    // S, then dS
  w.dst = reinterpret_cast<float*>(take((size_t)B * P * 4));  // dS^T
  w.dq = reinterpret_cast<float*>(take((size_t)B * E * 4));
  w.dpool = reinterpret_cast<float*>(take((size_t)P * E * 4));
  w.pos = reinterpret_cast<float*>(take((size_t)B * 4));
  w.gpos = reinterpret_cast<float*>(take((size_t)B * 4));
  const int64_t m = B > P ? B : P;
  w.scratch = take(dot_nt_scratch_bytes(m));
  w.bytes = (size_t)(p - reinterpret_cast<char*>(base));
  return w;
}

// query vectors of the positives, Q[i] = make_query(fixed_i, rel_i)   (one CTA per positive)
template <int M, bool HEAD>
__global__ void __launch_bounds__(kThreads) pooled_query_kernel(const float* __restrict__ ent,
                                                                const float* __restrict__ rel,
                                                                const int64_t* __restrict__ sample, int D,
                                                                int ent_stride, int rel_stride, float* q) {
  using T = Traits<M>;
  const int64_t i = blockIdx.x;
  const int64_t h = sample[3 * i], r = sample[3 * i + 1], t = sample[3 * i + 2];
  const float* fixed = ent + (HEAD ? t : h) * (int64_t)ent_stride;
  const float* relrow = rel + r * (int64_t)rel_stride;
  float* out = q + i * (int64_t)ent_stride;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float q0, q1;
    make_query<M, HEAD>(fixed[d], T::NC == 2 ? fixed[D + d] : 0.f, relrow[d], T::RC == 2 ? relrow[D + d] : 0.f, q0, q1);
    out[d] = q0;
    if constexpr (T::NC == 2) out[D + d] = q1;
  }
}

__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const float* __restrict__ table,
                                                               const int64_t* __restrict__ ids, int width,
                                                               float* __restrict__ out) {
  const float* src = table + ids[blockIdx.x] * (int64_t)width;
  float* dst = out + (int64_t)blockIdx.x * width;
  for (int d = threadIdx.x; d < width; d += blockDim.x) dst[d] = __ldg(src + d);
}

// out[c, r] = in[r, c] through a padded 32 x 32 shared-memory tile (coalesced both ways)
__global__ void __launch_bounds__(kThreads) transpose_kernel(const float* __restrict__ in, int rows, int cols,
                                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int k = ty; k < 32; k += kWarps) {
    const int r = r0 + k, c = c0 + tx;
    tile[k][tx] = (r < rows && c < cols) ? in[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += kWarps) {
    const int c = c0 + k, r = r0 + tx;
    if (c < cols && r < rows) out[(int64_t)c * rows + r] = tile[tx][k];
  }
}

// per positive: its K scores out of S by pool position, self-adversarial terms, the dS row
__global__ void __launch_bounds__(kThreads) pooled_adv_kernel(float* s_ds, const int32_t* __restrict__ positions,
                                                              const float* __restrict__ pos_score,
                                                              const float* __restrict__ weight, int B, int K, int P,
                                                              float alpha, float* neg_score_out, float* coef_pos,
                                                              float* partials, unsigned int* ticket, float* stats) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[33];
  float* sc = smem;        // [K] scores, then coefficients
  float* ds = smem + K;    // [P]
  const int i = blockIdx.x, tid = threadIdx.x;
  float* srow = s_ds + (int64_t)i * P;
  const int32_t* prow = positions + (int64_t)i * K;
  for (int j = tid; j < K; j += kThreads) {
    const float v = srow[prow[j]];
    sc[j] = v;
    if (neg_score_out) neg_score_out[(int64_t)i * K + j] = v;
  }
  for (int k = tid; k < P; k += kThreads) ds[k] = 0.f;
  __syncthreads();
  const float w = weight[i], pos = pos_score[i];
  const float nt = adv_row_terms(sc, K, alpha, w, sc, red);  // in place: sc[j] <- w a_j sigmoid(n_j)
  __syncthreads();
  // a position may be taken several times (cyclic repetition of the survivors): equal scores, equal
  // coefficients, so the shared-memory atomic sum does not depend on the order
  for (int j = tid; j < K; j += kThreads) atomicAdd(&ds[prow[j]], sc[j]);
  __syncthreads();
  for (int k = tid; k < P; k += kThreads) srow[k] = ds[k];  // S row becomes the dS row
  if (tid == 0) {
    coef_pos[i] = -w * sigmoid(-pos);
    partials[i] = w * log_sigmoid(pos);
    partials[B + i] = w * nt;
    partials[2 * B + i] = w;
  }
  fold_partials(partials, B, ticket, gridDim.x, stats, red);
}

// dQ -> the positive's fixed entity row and relation row; also the scaled positive-score gradient
template <int M, bool HEAD>
__global__ void __launch_bounds__(kThreads) pooled_chain_kernel(const float* __restrict__ ent,
                                                                const float* __restrict__ rel,
                                                                const int64_t* __restrict__ sample,
                                                                const float* __restrict__ dq,
                                                                const float* __restrict__ coef_pos,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ grad_loss, int D,
                                                                int ent_stride, int rel_stride, float* grad_ent,
                                                                float* grad_rel, float* gpos_scaled) {
  using T = Traits<M>;
  const int64_t i = blockIdx.x;
  const float scale = (grad_loss ? __ldg(grad_loss) : 1.f) / (2.f * __ldg(stats + 2));
  const int64_t h = sample[3 * i], r = sample[3 * i + 1], t = sample[3 * i + 2];
  const int64_t fid = HEAD ? t : h;
  const float* fixed = ent + fid * (int64_t)ent_stride;
  const float* relrow = rel + r * (int64_t)rel_stride;
  const float* dqi = dq + i * (int64_t)ent_stride;
  float* gf = grad_ent + fid * (int64_t)ent_stride;
  float* gr = grad_rel + r * (int64_t)rel_stride;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float a0 = fixed[d], a1 = T::NC == 2 ? fixed[D + d] : 0.f;
    const float r0 = relrow[d], r1 = T::RC == 2 ? relrow[D + d] : 0.f;
    const float dq0 = scale * dqi[d], dq1 = T::NC == 2 ? scale * dqi[D + d] : 0.f;
    float da0, da1, dr0, dr1;
    query_bwd<M, HEAD>(dq0, dq1, a0, a1, r0, r1, da0, da1, dr0, dr1);
    const float v0[1] = {da0}, v1[1] = {da1}, w0[1] = {dr0}, w1[1] = {dr1};
    red_add<1>(gf + d, v0);
    if constexpr (T::NC == 2) red_add<1>(gf + D + d, v1);
    red_add<1>(gr + d, w0);
    if constexpr (T::RC == 2) red_add<1>(gr + D + d, w1);
  }
  if (threadIdx.x == 0) gpos_scaled[i] = scale * coef_pos[i];
}

// grad_ent[pool[p]] += scale * dPool[p]   (pool ids may repeat: atomics)
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(const float* __restrict__ rows,
                                                                const int64_t* __restrict__ ids, int width,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ grad_loss, float* table) {
  const float scale = (grad_loss ? __ldg(grad_loss) : 1.f) / (2.f * __ldg(stats + 2));
  const float* src = rows + (int64_t)blockIdx.x * width;
  float* dst = table + ids[blockIdx.x] * (int64_t)width;
  for (int d = threadIdx.x; d < width; d += blockDim.x) {
    const float v[1] = {scale * src[d]};
    red_add<1>(dst + d, v);
  }
}

static int check_pooled(const kge_tables_t* t, int mode, int64_t B, int64_t K, int64_t P) {
  if (!t || !t->entity || !t->relation) return KGE_E_NULL;
  if (t->model != KGE_DISTMULT && t->model != KGE_COMPLEX) return KGE_E_UNSUPPORTED;  // dot-product models only
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B <= 0 || K <= 0 || P <= 0 || B > INT32_MAX || K > INT32_MAX || P > INT32_MAX || t->hidden_dim <= 0)
    return KGE_E_SIZE;
  if ((size_t)(K + P) * sizeof(float) > 200 * 1024) return KGE_E_UNSUPPORTED;
  return KGE_OK;
}

static void transpose(const float* in, int rows, int cols, float* out, cudaStream_t st) {
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  transpose_kernel<<<grid, kThreads, 0, st>>>(in, rows, cols, out);
}

}  // namespace kge

using namespace kge;

extern "C" size_t kge_pooled_workspace_bytes(const kge_tables_t* t, int64_t B, int64_t K, int64_t P) {
  if (!t || B <= 0 || K <= 0 || P <= 0) return 0;
  const int64_t E = (int64_t)t->hidden_dim * entity_comps(t->model);
  return carve(nullptr, B, P, E).bytes + 256;
}

extern "C" int kge_pooled_dot_fwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                                  const int64_t* pool, int64_t P, const int32_t* positions, int64_t K,
                                  const float* weight, float alpha, float* pos_score, float* neg_score,
                                  float* coef_pos, float* stats, void* workspace, void* loss_workspace,
                                  kge_stream_t stream) {
  int rc = check_pooled(t, mode, B, K, P);
  if (rc) return rc;
  if (!sample || !pool || !positions || !weight || !coef_pos || !stats || !workspace || !loss_workspace)
    return KGE_E_NULL;
  if (!aligned16(workspace)) return KGE_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = t->hidden_dim;
  const int E = D * entity_comps(t->model), RS = D * relation_comps(t->model);
  const PooledWs w = carve(workspace, B, P, E);
  const bool head = mode == KGE_HEAD_BATCH;
  if (t->model == KGE_COMPLEX) {
    if (head) pooled_query_kernel<KGE_COMPLEX, true><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, D, E, RS, w.q);
    else pooled_query_kernel<KGE_COMPLEX, false><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, D, E, RS, w.q);
  } else {
    if (head) pooled_query_kernel<KGE_DISTMULT, true><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, D, E, RS, w.q);
    else pooled_query_kernel<KGE_DISTMULT, false><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, D, E, RS, w.q);
  }
  gather_rows_kernel<<<(unsigned)P, kThreads, 0, st>>>(t->entity, pool, E, w.pool);
  KGE_LAUNCH_CHECK();
  rc = dot_nt_launch(w.q, w.pool, B, P, E, w.s, w.scratch, st);  // S = Q · Pool^T
  if (rc) return rc;
  float* pos = pos_score ? pos_score : w.pos;
  rc = kge_score_fwd(t, KGE_TAIL_BATCH, sample, B, nullptr, 0, pos, stream);  // positives: the reference's own call
  if (rc) return rc;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(loss_workspace);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(loss_workspace) + 16);
  const size_t smem = (size_t)(K + P) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pooled_adv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  pooled_adv_kernel<<<(unsigned)B, kThreads, smem, st>>>(w.s, positions, pos, weight, (int)B, (int)K, (int)P, alpha,
                                                        neg_score, coef_pos, partials, ticket, stats);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_pooled_dot_bwd(const kge_tables_t* t, int mode, const int64_t* sample, int64_t B,
                                  const int64_t* pool, int64_t P, int64_t K, const float* coef_pos,
                                  const float* stats, const float* grad_loss, float* grad_entity,
                                  float* grad_relation, void* workspace, kge_stream_t stream) {
  int rc = check_pooled(t, mode, B, K, P);
  if (rc) return rc;
  if (!sample || !pool || !coef_pos || !stats || !grad_entity || !grad_relation || !workspace) return KGE_E_NULL;
  if (!aligned16(workspace)) return KGE_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = t->hidden_dim;
  const int E = D * entity_comps(t->model), RS = D * relation_comps(t->model);
  const PooledWs w = carve(workspace, B, P, E);
  const bool head = mode == KGE_HEAD_BATCH;
  // operands of the two backward GEMMs in "A · B^T" form
  transpose(w.pool, (int)P, E, w.poolt, st);  // Pool^T [E, P]
  transpose(w.q, (int)B, E, w.qt, st);        // Q^T    [E, B]
  transpose(w.s, (int)B, (int)P, w.dst, st);  // dS^T   [P, B]
  KGE_LAUNCH_CHECK();
  rc = dot_nt_launch(w.s, w.poolt, B, E, (int)P, w.dq, w.scratch, st);  // dQ = dS · Pool
  if (rc) return rc;
  rc = dot_nt_launch(w.dst, w.qt, P, E, (int)B, w.dpool, w.scratch, st);  // dPool = dS^T · Q
  if (rc) return rc;
  if (t->model == KGE_COMPLEX) {
    if (head) pooled_chain_kernel<KGE_COMPLEX, true><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, w.dq, coef_pos, stats, grad_loss, D, E, RS, grad_entity, grad_relation, w.gpos);
    else pooled_chain_kernel<KGE_COMPLEX, false><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, w.dq, coef_pos, stats, grad_loss, D, E, RS, grad_entity, grad_relation, w.gpos);
  } else {
    if (head) pooled_chain_kernel<KGE_DISTMULT, true><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, w.dq, coef_pos, stats, grad_loss, D, E, RS, grad_entity, grad_relation, w.gpos);
    else pooled_chain_kernel<KGE_DISTMULT, false><<<(unsigned)B, kThreads, 0, st>>>(t->entity, t->relation, sample, w.dq, coef_pos, stats, grad_loss, D, E, RS, grad_entity, grad_relation, w.gpos);
  }
  scatter_rows_kernel<<<(unsigned)P, kThreads, 0, st>>>(w.dpool, pool, E, stats, grad_loss, grad_entity);
  KGE_LAUNCH_CHECK();
  // the positive score's own gradient: the unfused positives-only backward with the scaled coefficients
  return kge_score_bwd(t, KGE_TAIL_BATCH, sample, B, nullptr, 0, w.gpos, grad_entity, grad_relation, stream);
}
