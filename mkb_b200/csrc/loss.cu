// Stand-alone self-adversarial loss (used when the caller keeps the reference's three-call
// sequence model(sample) / model(sample, neg, mode) / loss(...) instead of the fused kernel).
// Replaces mkb/losses/adversarial.py:21-30 and its autograd backward.
#include "kge_common.cuh"

namespace kge {

__global__ void __launch_bounds__(kThreads) adv_loss_fwd_kernel(const float* __restrict__ pos,
                                                                const float* __restrict__ neg,
                                                                const float* __restrict__ weight, int B,
                                                                int K, float alpha, float* partials,
                                                                unsigned int* ticket, float* stats) {
  __shared__ float red[33];
  const int i = blockIdx.x;
  const float w = weight[i];
  const float nt = adv_row_terms(neg + (int64_t)i * K, K, alpha, 0.f, nullptr, red);
  if (threadIdx.x == 0) {
    partials[i] = w * log_sigmoid(pos[i]);
    partials[B + i] = w * nt;
    partials[2 * B + i] = w;
  }
  fold_partials(partials, B, ticket, gridDim.x, stats, red);
}

__global__ void __launch_bounds__(kThreads) adv_loss_bwd_kernel(const float* __restrict__ pos,
                                                                const float* __restrict__ neg,
                                                                const float* __restrict__ weight, int B,
                                                                int K, float alpha,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ grad_loss,
                                                                float* grad_pos, float* grad_neg) {
  __shared__ float red[33];
  const int i = blockIdx.x;
  const float scale = (grad_loss ? __ldg(grad_loss) : 1.f) / (2.f * __ldg(stats + 2));
  const float w = weight[i] * scale;
  adv_row_terms(neg + (int64_t)i * K, K, alpha, w, grad_neg + (int64_t)i * K, red);
  if (threadIdx.x == 0) grad_pos[i] = -w * sigmoid(-pos[i]);
}

// ------------------------------------------------------------------------------------------------
// KL divergence between row softmaxes (distillation): replaces losses.KlDivergence.__call__
// (mkb/losses/kl_divergence.py:22-29)
//     mean_{i,j} q_ij (log q_ij - log p_ij),  p = softmax(student / T, dim=1),  q = softmax(teacher / T, dim=1)
// One CTA per row; row sums are folded by the last CTA in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
struct RowSoftmax {
  float mx, lse;  // max of x/T and log(sum exp(x/T - mx)); x/T rounded like the reference's division
};
__device__ __forceinline__ RowSoftmax row_softmax(const float* __restrict__ x, int K, float temp, float* red) {
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < K; j += blockDim.x) mx = fmaxf(mx, __fdiv_rn(x[j], temp));
  mx = block_max(mx, red);
  float se = 0.f;
  for (int j = threadIdx.x; j < K; j += blockDim.x) se += expf(__fdiv_rn(x[j], temp) - mx);
  se = block_sum(se, red);
  return RowSoftmax{mx, logf(se)};
}

__global__ void __launch_bounds__(kThreads) kl_fwd_kernel(const float* __restrict__ student,
                                                          const float* __restrict__ teacher, int B, int K,
                                                          float temp, float* partials, unsigned int* ticket,
                                                          float* loss) {
  __shared__ float red[33];
  __shared__ bool is_last;
  const int i = blockIdx.x;
  const float* s = student + (int64_t)i * K;
  const float* t = teacher + (int64_t)i * K;
  const RowSoftmax ps = row_softmax(s, K, temp, red);
  const RowSoftmax qs = row_softmax(t, K, temp, red);
  float acc = 0.f;
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    const float lq = __fdiv_rn(t[j], temp) - qs.mx - qs.lse;
    const float lp = __fdiv_rn(s[j], temp) - ps.mx - ps.lse;
    const float q = expf(lq);
    if (q > 0.f) acc = fmaf(q, lq - lp, acc);  // xlogy: q == 0 contributes 0
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[i] = acc;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float tot = 0.f;
  for (int k = threadIdx.x; k < B; k += blockDim.x) tot += __ldcg(partials + k);
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) {
    loss[0] = tot / ((float)B * (float)K);
    *ticket = 0u;
  }
}

__global__ void __launch_bounds__(kThreads) kl_bwd_kernel(const float* __restrict__ student,
                                                          const float* __restrict__ teacher, int B, int K,
                                                          float temp, const float* __restrict__ grad_loss,
                                                          float* grad_student, float* grad_teacher) {
  __shared__ float red[33];
  const int i = blockIdx.x;
  const float* s = student + (int64_t)i * K;
  const float* t = teacher + (int64_t)i * K;
  const RowSoftmax ps = row_softmax(s, K, temp, red);
  const RowSoftmax qs = row_softmax(t, K, temp, red);
  const float g = (grad_loss ? __ldg(grad_loss) : 1.f) / ((float)B * (float)K) / temp;
  float kl_row = 0.f;
  if (grad_teacher) {
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
      const float lq = __fdiv_rn(t[j], temp) - qs.mx - qs.lse;
      const float q = expf(lq);
      if (q > 0.f) kl_row = fmaf(q, lq - (__fdiv_rn(s[j], temp) - ps.mx - ps.lse), kl_row);
    }
    kl_row = block_sum(kl_row, red);
  }
  for (int j = threadIdx.x; j < K; j += blockDim.x) {
    const float lq = __fdiv_rn(t[j], temp) - qs.mx - qs.lse;
    const float lp = __fdiv_rn(s[j], temp) - ps.mx - ps.lse;
    const float q = expf(lq);
    grad_student[(int64_t)i * K + j] = g * (expf(lp) - q);
    if (grad_teacher) grad_teacher[(int64_t)i * K + j] = g * q * ((lq - lp) - kl_row);
  }
}

}  // namespace kge

using namespace kge;

extern "C" int kge_adv_loss_fwd(const float* pos_score, const float* neg_score, const float* weight,
                                int64_t B, int64_t K, float alpha, float* stats, void* workspace,
                                kge_stream_t stream) {
  if (!pos_score || !neg_score || !weight || !stats || !workspace) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX) return KGE_E_SIZE;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 16);
  adv_loss_fwd_kernel<<<(unsigned)B, kThreads, 0, (cudaStream_t)stream>>>(
      pos_score, neg_score, weight, (int)B, (int)K, alpha, partials, ticket, stats);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_adv_loss_bwd(const float* pos_score, const float* neg_score, const float* weight,
                                int64_t B, int64_t K, float alpha, const float* stats,
                                const float* grad_loss, float* grad_pos, float* grad_neg,
                                kge_stream_t stream) {
  if (!pos_score || !neg_score || !weight || !stats || !grad_pos || !grad_neg) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX) return KGE_E_SIZE;
  adv_loss_bwd_kernel<<<(unsigned)B, kThreads, 0, (cudaStream_t)stream>>>(
      pos_score, neg_score, weight, (int)B, (int)K, alpha, stats, grad_loss, grad_pos, grad_neg);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_kl_div_fwd(const float* student, const float* teacher, int64_t B, int64_t K, float T,
                              float* loss, void* workspace, kge_stream_t stream) {
  if (!student || !teacher || !loss || !workspace) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX || !(T > 0.f)) return KGE_E_SIZE;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 16);
  kl_fwd_kernel<<<(unsigned)B, kThreads, 0, (cudaStream_t)stream>>>(student, teacher, (int)B, (int)K, T,
                                                                    partials, ticket, loss);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_kl_div_bwd(const float* student, const float* teacher, int64_t B, int64_t K, float T,
                              const float* grad_loss, float* grad_student, float* grad_teacher,
                              kge_stream_t stream) {
  if (!student || !teacher || !grad_student) return KGE_E_NULL;
  if (B <= 0 || B > INT32_MAX || K <= 0 || K > INT32_MAX || !(T > 0.f)) return KGE_E_SIZE;
  kl_bwd_kernel<<<(unsigned)B, kThreads, 0, (cudaStream_t)stream>>>(student, teacher, (int)B, (int)K, T,
                                                                    grad_loss, grad_student, grad_teacher);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
