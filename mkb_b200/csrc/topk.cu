// K8: exact top-k of every row of a score matrix — the reduce behind utils.TopK
// (mkb/utils/top_k.py:226-234: argsort(model(sample), descending)[:k]) and the top-k negative sampling
// of the distillation loop (mkb/distillation/top_k_sampling.py:664-677), which argsort ALL N scores to
// keep a handful.
//
// One CTA per row, no sort of the row:
//   1. floats are mapped to order-preserving 32-bit keys (ascending key <=> descending score);
//   2. a 4-pass, 8-bit radix SELECT over a shared-memory histogram finds the key T of the k-th best
//      score and how many entries equal to T still belong to the answer;
//   3. one ordered sweep compacts {key < T} plus the first ties in column order into shared memory;
//   4. a bitonic sort of those k (key, column) pairs gives the output order: descending score, ties by
//      ascending column — the order of a stable descending argsort.
// The row is read 5 times (L2-resident for any entity table of the configs), k <= 1024.
#include "kge_common.cuh"

namespace kge {

constexpr int kMaxTopK = 1024;

__device__ __forceinline__ uint32_t desc_key(float f) {
  uint32_t u = __float_as_uint(f);
  if (u == 0x80000000u) u = 0u;                        // -0.0 == +0.0
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending u <=> ascending float
  return ~u;                                           // ascending key <=> descending float
}

__global__ void __launch_bounds__(kThreads) topk_rows_kernel(const float* __restrict__ scores, int64_t cols,
                                                             int64_t row_stride, int k, int sort_pow2,
                                                             int64_t* __restrict__ idx_out,
                                                             float* __restrict__ val_out) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long sel[kMaxTopK];
  __shared__ unsigned int s_bin, s_remaining, s_warp_eq[kWarps], s_lt_count, s_eq_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = scores + (int64_t)blockIdx.x * row_stride;

  // ---- 2. radix select of the k-th smallest key
  uint32_t prefix = 0, mask = 0;
  if (tid == 0) s_remaining = (unsigned)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;  // kThreads == 256 bins
    __syncthreads();
    for (int64_t j = tid; j < cols; j += kThreads) {
      const uint32_t key = desc_key(__ldg(row + j));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned cum = 0, b = 0;
      const unsigned need = s_remaining;
      for (; b < 255; ++b) {
        if (cum + hist[b] >= need) break;
        cum += hist[b];
      }
      s_bin = b;
      s_remaining = need - cum;
    }
    __syncthreads();
    prefix |= s_bin << shift;
    mask |= 0xffu << shift;
  }
  const uint32_t T = prefix;
  const unsigned take_eq = s_remaining;  // >= 1 entries equal to T belong to the answer
  const unsigned n_lt = (unsigned)k - take_eq;
  if (tid == 0) {
    s_lt_count = 0;
    s_eq_base = 0;
  }
  __syncthreads();

  // ---- 3. ordered compaction: every key < T (any order), the first take_eq keys == T in column order
  for (int64_t j0 = 0; j0 < cols; j0 += kThreads) {
    const int64_t j = j0 + tid;
    uint32_t key = 0xffffffffu;
    bool lt = false, eq = false;
    if (j < cols) {
      key = desc_key(__ldg(row + j));
      lt = key < T;
      eq = key == T;
    }
    const unsigned ballot = __ballot_sync(kFull, eq);
    if (lane == 0) s_warp_eq[warp] = __popc(ballot);
    __syncthreads();
    unsigned before = s_eq_base, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const unsigned c = s_warp_eq[w];
      if (w < warp) before += c;
      total += c;
    }
    before += __popc(ballot & ((1u << lane) - 1u));
    if (eq && before < take_eq) sel[n_lt + before] = ((unsigned long long)key << 32) | (unsigned long long)j;
    if (lt) sel[atomicAdd(&s_lt_count, 1u)] = ((unsigned long long)key << 32) | (unsigned long long)j;
    __syncthreads();
    if (tid == 0) s_eq_base += total;
    // s_eq_base is next read after the following iteration's first barrier
  }
  __syncthreads();

  // ---- 4. bitonic sort of the k selected (key, column) pairs
  for (int j = k + tid; j < sort_pow2; j += kThreads) sel[j] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= sort_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (sort_pow2 >> 1); t += kThreads) {
        const int lo_i = 2 * t - (t & (stride - 1));
        const int hi_i = lo_i + stride;
        const bool up = (lo_i & size) == 0;
        const unsigned long long a = sel[lo_i], b = sel[hi_i];
        if ((a > b) == up) {
          sel[lo_i] = b;
          sel[hi_i] = a;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += kThreads) {
    const int64_t col = (int64_t)(sel[j] & 0xffffffffull);
    idx_out[(int64_t)blockIdx.x * k + j] = col;
    if (val_out) val_out[(int64_t)blockIdx.x * k + j] = __ldg(row + col);
  }
}

}  // namespace kge

using namespace kge;

extern "C" int kge_topk_rows(const float* scores, int64_t rows, int64_t cols, int64_t row_stride, int32_t k,
                             int64_t* indices, float* values, kge_stream_t stream) {
  if (!scores || !indices) return KGE_E_NULL;
  if (rows < 0 || rows > INT32_MAX || cols <= 0 || cols > INT32_MAX || row_stride < cols) return KGE_E_SIZE;
  if (k < 1 || k > cols) return KGE_E_SIZE;
  if (k > kMaxTopK) return KGE_E_UNSUPPORTED;
  if (rows == 0) return KGE_OK;
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  topk_rows_kernel<<<(unsigned)rows, kThreads, 0, (cudaStream_t)stream>>>(scores, cols, row_stride, (int)k, p2, indices,
                                                                         values);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
