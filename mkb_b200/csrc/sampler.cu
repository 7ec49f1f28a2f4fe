// K4: negative sampling on the device.
//
// Replaces NegativeSampling.generate (mkb/sampling/negative_sampling.py:158-201) and the
// positive_triples dictionaries (:7-28).  The true-entity sets live on the GPU as a CSR keyed by
// relation * n_entity + fixed_entity (sorted keys -> binary search), members sorted per segment.
//
//   kge_sample_negatives : independent draws per output slot.  Slot (i, j) owns a Philox4x32-10
//       stream: key = seed, counter = (block_lo, block_hi, slot_lo, slot_hi) with
//       block = offset * 64 + attempt / 4 — the same (subsequence, offset) split curand's
//       Philox4_32_10 uses, written out here so the oracle can restate it bit for bit.
//       candidate = mulhi32(word, n_entity); members of the positive's true set are redrawn.
//   kge_filter_pool      : the reference's exact semantics for a host-drawn shared pool.
#include "kge_common.cuh"

namespace kge {

constexpr int kPhiloxBlocksPerCall = 64;  // 256 attempts per slot per call

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// index of `key` in sorted keys[0..n) or -1
__device__ __forceinline__ int64_t find_key(const int64_t* __restrict__ keys, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return (lo < n && __ldg(keys + lo) == key) ? lo : -1;
}

__device__ __forceinline__ bool is_member(const int64_t* __restrict__ m, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = __ldg(m + mid);
    if (v < x) lo = mid + 1;
    else if (v > x) hi = mid;
    else return true;
  }
  return false;
}

// Segment [seg_lo, seg_hi) of the positive's true set; bit0 of *status when the key is missing.
__device__ __forceinline__ void positive_segment(const kge_filter_csr_t& f, bool head_mode,
                                                 const int64_t* __restrict__ sample, int64_t i,
                                                 int64_t n_entity, int64_t& seg_lo, int64_t& seg_hi,
                                                 int32_t* status) {
  const int64_t h = sample[3 * i], r = sample[3 * i + 1], t = sample[3 * i + 2];
  const int64_t code = r * n_entity + (head_mode ? t : h);
  const int64_t k = find_key(f.keys, f.n_keys, code);
  if (k < 0) {
    seg_lo = seg_hi = 0;
    atomicOr(status, 1);
  } else {
    seg_lo = __ldg(f.offsets + k);
    seg_hi = __ldg(f.offsets + k + 1);
  }
}

constexpr int kMaxSortK = 2048;  // rows up to this length are sorted in shared memory

__global__ void __launch_bounds__(kThreads) sample_negatives_kernel(kge_filter_csr_t f, int head_mode,
                                                                    const int64_t* __restrict__ sample,
                                                                    int64_t K, int64_t n_entity,
                                                                    uint64_t seed, uint64_t offset,
                                                                    int sort_pow2,
                                                                    int64_t* __restrict__ out,
                                                                    int32_t* status) {
  extern __shared__ uint32_t s_ids[];  // sort_pow2 entries when sorting
  __shared__ int64_t s_seg[2];
  const int64_t i = blockIdx.x;
  if (threadIdx.x == 0) positive_segment(f, head_mode != 0, sample, i, n_entity, s_seg[0], s_seg[1], status);
  __syncthreads();
  const int64_t lo = s_seg[0], hi = s_seg[1];
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t n32 = (uint32_t)n_entity;
  for (int64_t j = threadIdx.x; j < K; j += blockDim.x) {
    const uint64_t slot = (uint64_t)i * (uint64_t)K + (uint64_t)j;
    int64_t cand = 0;
    bool found = false;
    for (int b = 0; b < kPhiloxBlocksPerCall && !found; ++b) {
      const uint64_t blk = offset * kPhiloxBlocksPerCall + b;
      const uint4 w = philox4x32_10(
          make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)slot, (uint32_t)(slot >> 32)), key);
      const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (!found) {
          cand = (int64_t)__umulhi(words[a], n32);
          found = !is_member(f.members, lo, hi, cand);
        }
      }
    }
    if (!found) {  // pathological true set: walk to the next non-member
      for (int64_t s = 0; s < n_entity && !found; ++s) {
        cand = cand + 1 == n_entity ? 0 : cand + 1;
        found = !is_member(f.members, lo, hi, cand);
      }
      if (!found) atomicOr(status, 2);
    }
    if (sort_pow2) s_ids[j] = (uint32_t)cand;
    else out[i * K + j] = cand;
  }
  if (sort_pow2) {
    // Ascending bitonic sort of the row: the set of negatives is unchanged (the loss is a symmetric
    // function of the row), but every CTA of the scoring kernels then walks the entity table in the
    // same direction, which keeps the active band of the table (and of its gradient) L2-resident.
    for (int j = (int)K + threadIdx.x; j < sort_pow2; j += blockDim.x) s_ids[j] = 0xFFFFFFFFu;
    __syncthreads();
    for (int size = 2; size <= sort_pow2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < (sort_pow2 >> 1); t += blockDim.x) {
          const int lo_i = 2 * t - (t & (stride - 1));
          const int hi_i = lo_i + stride;
          const bool up = (lo_i & size) == 0;
          const uint32_t a = s_ids[lo_i], b = s_ids[hi_i];
          if ((a > b) == up) {
            s_ids[lo_i] = b;
            s_ids[hi_i] = a;
          }
        }
        __syncthreads();
      }
    }
    for (int64_t j = threadIdx.x; j < K; j += blockDim.x) out[i * K + j] = (int64_t)s_ids[j];
  }
}

// Reference pool semantics: one warp per positive, ballot-compaction of the survivors in pool
// order, cyclic repetition when fewer than K survive (negative_sampling.py:176-195).
__global__ void __launch_bounds__(kThreads) filter_pool_kernel(kge_filter_csr_t f, int head_mode,
                                                               const int64_t* __restrict__ sample,
                                                               int64_t B, int64_t K, int64_t n_entity,
                                                               const int64_t* __restrict__ pool,
                                                               int64_t pool_size,
                                                               int64_t* __restrict__ out, int32_t* status,
                                                               int32_t* __restrict__ positions) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (i >= B) return;
  int64_t lo = 0, hi = 0;
  if (lane == 0) positive_segment(f, head_mode != 0, sample, i, n_entity, lo, hi, status);
  lo = __shfl_sync(kFull, lo, 0);
  hi = __shfl_sync(kFull, hi, 0);
  int64_t* dst = out + i * K;
  int32_t* pdst = positions ? positions + i * K : nullptr;  // index into the pool of every chosen negative
  int64_t cnt = 0;
  for (int64_t base = 0; base < pool_size && cnt < K; base += 32) {
    const int64_t k = base + lane;
    const int64_t x = k < pool_size ? __ldg(pool + k) : 0;
    const bool keep = k < pool_size && !is_member(f.members, lo, hi, x);
    const unsigned mask = __ballot_sync(kFull, keep);
    const int64_t pos = cnt + __popc(mask & ((1u << lane) - 1u));
    if (keep && pos < K) {
      dst[pos] = x;
      if (pdst) pdst[pos] = (int32_t)k;
    }
    cnt += __popc(mask);
  }
  __syncwarp();
  if (cnt == 0) {
    if (lane == 0) atomicOr(status, 4);
    for (int64_t k = lane; k < K; k += 32) {
      dst[k] = 0;
      if (pdst) pdst[k] = 0;
    }
    return;
  }
  // survivors repeat cyclically: dst[k] = dst[k mod cnt]  (np.concatenate of re-filtered pools)
  for (int64_t k = cnt + lane; k < K; k += 32) {
    dst[k] = dst[k % cnt];
    if (pdst) pdst[k] = pdst[k % cnt];
  }
}

}  // namespace kge

using namespace kge;

static int check_filter(const kge_filter_csr_t* f) {
  if (!f) return KGE_E_NULL;
  if (f->n_keys < 0) return KGE_E_SIZE;
  if (f->n_keys > 0 && (!f->keys || !f->offsets || !f->members)) return KGE_E_NULL;
  return KGE_OK;
}

extern "C" int kge_sample_negatives(const kge_filter_csr_t* filter, int mode, const int64_t* sample,
                                    int64_t B, int64_t K, int64_t n_entity, uint64_t seed,
                                    uint64_t offset, int sort_rows, int64_t* negatives, int32_t* status,
                                    kge_stream_t stream) {
  int rc = check_filter(filter);
  if (rc) return rc;
  if (!sample || !negatives || !status) return KGE_E_NULL;
  if (B < 0 || K <= 0 || n_entity <= 0 || n_entity > 0xFFFFFFFFLL || B > INT32_MAX) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B == 0) return KGE_OK;
  const int threads = K >= 256 ? 256 : (int)((K + 31) / 32 * 32);
  int pow2 = 0;
  if (sort_rows && K > 1 && K <= kMaxSortK) {
    pow2 = 2;
    while (pow2 < K) pow2 <<= 1;
  }
  sample_negatives_kernel<<<(unsigned)B, threads, (size_t)pow2 * sizeof(uint32_t), (cudaStream_t)stream>>>(
      *filter, mode == KGE_HEAD_BATCH, sample, K, n_entity, seed, offset, pow2, negatives, status);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_filter_pool(const kge_filter_csr_t* filter, int mode, const int64_t* sample, int64_t B,
                               int64_t K, int64_t n_entity, const int64_t* pool, int64_t pool_size,
                               int64_t* negatives, int32_t* status, kge_stream_t stream) {
  return kge_filter_pool_positions(filter, mode, sample, B, K, n_entity, pool, pool_size, negatives, nullptr, status,
                                   stream);
}

extern "C" int kge_filter_pool_positions(const kge_filter_csr_t* filter, int mode, const int64_t* sample, int64_t B,
                                         int64_t K, int64_t n_entity, const int64_t* pool, int64_t pool_size,
                                         int64_t* negatives, int32_t* positions, int32_t* status,
                                         kge_stream_t stream) {
  int rc = check_filter(filter);
  if (rc) return rc;
  if (!sample || !negatives || !status || !pool) return KGE_E_NULL;
  if (B < 0 || K <= 0 || n_entity <= 0 || pool_size <= 0 || B > INT32_MAX) return KGE_E_SIZE;
  if (mode != KGE_TAIL_BATCH && mode != KGE_HEAD_BATCH) return KGE_E_MODE;
  if (B == 0) return KGE_OK;
  filter_pool_kernel<<<(unsigned)((B + kWarps - 1) / kWarps), kThreads, 0, (cudaStream_t)stream>>>(
      *filter, mode == KGE_HEAD_BATCH, sample, B, K, n_entity, pool, pool_size, negatives, status, positions);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
