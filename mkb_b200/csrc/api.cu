// Version / error text / device query + the dense Adam step.
#include "kge_common.cuh"

namespace kge {

// torch.optim.Adam (no weight decay, no amsgrad), single-tensor form, called by the user's
// optimizer at mkb/compose/pipeline.py:238; optimizer.zero_grad() (:240) is folded in.
__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ p, float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        int64_t n, float lr_bc1, float inv_sqrt_bc2,
                                                        float b1, float b2, float eps, int zero_grad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k * 4 < n; k += stride) {
    const int64_t e = k * 4;
    if (e + 4 <= n) {
      float4 gp = *reinterpret_cast<float4*>(g + e);
      float4 pp = *reinterpret_cast<float4*>(p + e);
      float4 mm = *reinterpret_cast<float4*>(m + e);
      float4 vv = *reinterpret_cast<float4*>(v + e);
      float* gf = reinterpret_cast<float*>(&gp);
      float* pf = reinterpret_cast<float*>(&pp);
      float* mf = reinterpret_cast<float*>(&mm);
      float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mf[c] = b1 * mf[c] + (1.f - b1) * gf[c];
        vf[c] = b2 * vf[c] + (1.f - b2) * gf[c] * gf[c];
        const float denom = sqrtf(vf[c]) * inv_sqrt_bc2 + eps;
        pf[c] -= lr_bc1 * (mf[c] / denom);
      }
      *reinterpret_cast<float4*>(p + e) = pp;
      *reinterpret_cast<float4*>(m + e) = mm;
      *reinterpret_cast<float4*>(v + e) = vv;
      if (zero_grad) *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int64_t c = e; c < n; ++c) {
        const float gg = g[c];
        const float mc = b1 * m[c] + (1.f - b1) * gg;
        const float vc = b2 * v[c] + (1.f - b2) * gg * gg;
        m[c] = mc;
        v[c] = vc;
        p[c] -= lr_bc1 * (mc / (sqrtf(vc) * inv_sqrt_bc2 + eps));
        if (zero_grad) g[c] = 0.f;
      }
    }
  }
}

// Same update for one column chunk: grad is a dense [rows, comps*ncols] chunk buffer, param / moments
// keep the table layout (row stride `stride`, second component at +im_off, chunk starts at col0).
__global__ void __launch_bounds__(kThreads) adam_chunk_kernel(float* __restrict__ p, float* __restrict__ g,
                                                              float* __restrict__ m, float* __restrict__ v,
                                                              int64_t rows, int comps, int ncols, int col0,
                                                              int stride, int im_off, float lr_bc1,
                                                              float inv_sqrt_bc2, float b1, float b2, float eps,
                                                              int zero_grad) {
  const int per_row = comps * ncols;  // multiple of 4
  const int64_t total4 = rows * per_row / 4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total4; k += step) {
    const int64_t e = k * 4;
    const int64_t row = e / per_row;
    const int rem = (int)(e - row * per_row);
    const int c = rem / ncols, d = rem - c * ncols;
    const int64_t pi = row * stride + (int64_t)c * im_off + col0 + d;
    float4 gp = *reinterpret_cast<float4*>(g + e);
    float4 pp = *reinterpret_cast<float4*>(p + pi);
    float4 mm = *reinterpret_cast<float4*>(m + pi);
    float4 vv = *reinterpret_cast<float4*>(v + pi);
    float* gf = reinterpret_cast<float*>(&gp);
    float* pf = reinterpret_cast<float*>(&pp);
    float* mf = reinterpret_cast<float*>(&mm);
    float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mf[q] = b1 * mf[q] + (1.f - b1) * gf[q];
      vf[q] = b2 * vf[q] + (1.f - b2) * gf[q] * gf[q];
      pf[q] -= lr_bc1 * (mf[q] / (sqrtf(vf[q]) * inv_sqrt_bc2 + eps));
    }
    *reinterpret_cast<float4*>(p + pi) = pp;
    *reinterpret_cast<float4*>(m + pi) = mm;
    *reinterpret_cast<float4*>(v + pi) = vv;
    if (zero_grad) *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Column-parallel multi-GPU step: this rank owns hidden-dim columns [col0, col0+ncols) of every row.
// grad / moments are dense slice buffers [rows, comps*ncols]; the parameter keeps the table layout and
// exists once per GPU.  The kernel reads the local replica, applies Adam, and stores the new values
// into EVERY replica (peer pointers mapped over NVLink): the update and its all-gather are one pass,
// the remote stores are fire-and-forget 16-byte writes.
struct Replicas {
  float* p[16];
};

__global__ void __launch_bounds__(kThreads) adam_slice_bcast_kernel(Replicas reps, int n_rep, int self,
                                                                    float* __restrict__ g, float* __restrict__ m,
                                                                    float* __restrict__ v, int64_t rows, int comps,
                                                                    int ncols, int col0, int stride, int im_off,
                                                                    float lr_bc1, float inv_sqrt_bc2, float b1,
                                                                    float b2, float eps, int zero_grad) {
  const int per_row = comps * ncols;
  const int64_t total4 = rows * per_row / 4;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const float* __restrict__ src = reps.p[self];
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total4; k += step) {
    const int64_t e = k * 4;
    const int64_t row = e / per_row;
    const int rem = (int)(e - row * per_row);
    const int c = rem / ncols, d = rem - c * ncols;
    const int64_t pi = row * stride + (int64_t)c * im_off + col0 + d;
    float4 gp = *reinterpret_cast<float4*>(g + e);
    float4 pp = *reinterpret_cast<const float4*>(src + pi);
    float4 mm = *reinterpret_cast<float4*>(m + e);
    float4 vv = *reinterpret_cast<float4*>(v + e);
    float* gf = reinterpret_cast<float*>(&gp);
    float* pf = reinterpret_cast<float*>(&pp);
    float* mf = reinterpret_cast<float*>(&mm);
    float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mf[q] = b1 * mf[q] + (1.f - b1) * gf[q];
      vf[q] = b2 * vf[q] + (1.f - b2) * gf[q] * gf[q];
      pf[q] -= lr_bc1 * (mf[q] / (sqrtf(vf[q]) * inv_sqrt_bc2 + eps));
    }
    *reinterpret_cast<float4*>(m + e) = mm;
    *reinterpret_cast<float4*>(v + e) = vv;
    if (zero_grad) *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < n_rep; ++r) *reinterpret_cast<float4*>(reps.p[r] + pi) = pp;
  }
}

}  // namespace kge

using namespace kge;

extern "C" int kge_abi_version(void) { return KGE_ABI_VERSION; }

extern "C" const char* kge_strerror(int code) {
  switch (code) {
    case KGE_OK: return "ok";
    case KGE_E_NULL: return "kge: required pointer is NULL";
    case KGE_E_SIZE: return "kge: invalid size";
    case KGE_E_MODEL: return "kge: unknown model";
    case KGE_E_MODE: return "kge: unknown mode";
    case KGE_E_ALIGN: return "kge: misaligned pointer";
    case KGE_E_UNSUPPORTED: return "kge: shape not supported by this kernel";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "kge: unknown error";
}

extern "C" int kge_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return (int)e;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return KGE_OK;
}

extern "C" int kge_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                             int64_t step, float lr, float beta1, float beta2, float eps, int zero_grad,
                             kge_stream_t stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return KGE_E_NULL;
  if (n < 0 || step < 1) return KGE_E_SIZE;
  if (n == 0) return KGE_OK;
  if (!aligned16(param) || !aligned16(grad) || !aligned16(exp_avg) || !aligned16(exp_avg_sq))
    return KGE_E_ALIGN;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float lr_bc1 = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (n / 4 + kThreads - 1) / kThreads + 1;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  adam_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr_bc1, inv_sqrt_bc2, beta1, beta2, eps, zero_grad);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_adam_step_chunk(float* param, float* grad_chunk, float* exp_avg, float* exp_avg_sq,
                                   int64_t rows, int32_t comps, int32_t ncols, int32_t col0, int32_t row_stride,
                                   int32_t im_off, int64_t step, float lr, float beta1, float beta2, float eps,
                                   int zero_grad, kge_stream_t stream) {
  if (!param || !grad_chunk || !exp_avg || !exp_avg_sq) return KGE_E_NULL;
  if (rows < 0 || step < 1 || comps < 1 || comps > 2 || ncols <= 0) return KGE_E_SIZE;
  if (rows == 0) return KGE_OK;
  if ((ncols | col0 | row_stride | im_off) % 4 != 0) return KGE_E_ALIGN;
  if (!aligned16(param) || !aligned16(grad_chunk) || !aligned16(exp_avg) || !aligned16(exp_avg_sq))
    return KGE_E_ALIGN;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (rows * comps * ncols / 4 + kThreads - 1) / kThreads;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  adam_chunk_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      param, grad_chunk, exp_avg, exp_avg_sq, rows, comps, ncols, col0, row_stride, im_off,
      (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps, zero_grad);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_adam_slice_bcast(float* const* param_replicas, int32_t n_replicas, int32_t self_index,
                                    float* grad_slice, float* exp_avg_slice, float* exp_avg_sq_slice,
                                    int64_t rows, int32_t comps, int32_t ncols, int32_t col0, int32_t row_stride,
                                    int32_t im_off, int64_t step, float lr, float beta1, float beta2, float eps,
                                    int zero_grad, kge_stream_t stream) {
  if (!param_replicas || !grad_slice || !exp_avg_slice || !exp_avg_sq_slice) return KGE_E_NULL;
  if (n_replicas < 1 || n_replicas > 16 || self_index < 0 || self_index >= n_replicas) return KGE_E_SIZE;
  if (rows < 0 || step < 1 || comps < 1 || comps > 2 || ncols <= 0) return KGE_E_SIZE;
  if (rows == 0) return KGE_OK;
  if ((ncols | col0 | row_stride | im_off) % 4 != 0) return KGE_E_ALIGN;
  Replicas reps{};
  for (int r = 0; r < n_replicas; ++r) {
    if (!param_replicas[r]) return KGE_E_NULL;
    if (!aligned16(param_replicas[r])) return KGE_E_ALIGN;
    reps.p[r] = param_replicas[r];
  }
  if (!aligned16(grad_slice) || !aligned16(exp_avg_slice) || !aligned16(exp_avg_sq_slice)) return KGE_E_ALIGN;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = (rows * comps * ncols / 4 + kThreads - 1) / kThreads;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  if (blocks < 1) blocks = 1;
  adam_slice_bcast_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      reps, n_replicas, self_index, grad_slice, exp_avg_slice, exp_avg_sq_slice, rows, comps, ncols, col0,
      row_stride, im_off, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps, zero_grad);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
