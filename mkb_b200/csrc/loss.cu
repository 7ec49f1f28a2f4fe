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
