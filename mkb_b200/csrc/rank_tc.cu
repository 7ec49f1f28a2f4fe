// K6: tensor-core ranking for the dot-product models (ComplEx, DistMult).
//
// For these two models the all-entity score matrix is a plain GEMM,  S[q, e] = sum_k Q[q, k] * E[e, k],
// with Q the query vectors (h∘r, conj(r)∘t, h*r, r*t — built by rank_prepare_kernel) and E the entity
// table itself (mkb/models/complex.py:74-85, distmult.py:68-73 with all N entities as candidates,
// evaluation/evaluation.py:237).  GEMM-shaped work belongs on the 5th-generation tensor cores:
//
//   * operands: fp32 tiles (128 queries x 32 k, 256 entities x 32 k) brought in by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 2-stage shared-memory ring;
//   * precision: 3xTF32.  x = hi + lo with hi = x truncated to TF32 (what the tensor core reads when
//     handed an fp32 word) and lo = x - hi (exact).  S ≈ hi·hi + hi·lo + lo·hi, accumulated in fp32:
//     per-product error ~2^-20, i.e. fp32-grade, so scores meet the 1e-4 bar and ranks only move on
//     fp32-level near-ties.  Four warps form the lo tiles in shared memory (element-wise, layout
//     agnostic) while the previous stage is being multiplied;
//   * MMA: tcgen05.mma.cta_group::1.kind::tf32, M=128, N=256, K=8, issued by one thread, accumulator
//     in TMEM (256 columns x 128 lanes fp32), completion signalled with tcgen05.commit -> mbarrier;
//   * epilogue fused with the ranking: TMEM -> registers (tcgen05.ld 32x32b), compare with the
//     positive's score, filter through the CSR, count; the [Q, N] score matrix never exists.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 lo-splitters during the main loop and the epilogue afterwards (warp w owns TMEM lanes
// 32*(w%4)..+31).
#include <cuda.h>

#include <cstdlib>

#include "kge_common.cuh"

namespace kge {

constexpr int TC_M = 128, TC_N = 256, TC_K = 32;
constexpr int TC_A_BYTES = TC_M * TC_K * 4;  // 16 KB
constexpr int TC_B_BYTES = TC_N * TC_K * 4;  // 32 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;  // A, A_lo, B, B_lo = 96 KB
constexpr int tc_smem_bytes(int stages) { return stages * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }

struct RankTcParams {
  const int64_t* queries;
  kge_filter_csr_t filter;
  const float* pos_score;  // [Q]
  float* pos_out;          // diag pass: where the re-scored positives go
  int diag;                // 1: the "entity" operand is posrows [Q, Kd]; tile m only needs columns 128m .. 128m+127
  int k_split;             // > 1 (plain GEMM use only): blockIdx.z owns kb_per_split k-blocks, results are ADDED into
  int kb_per_split;        //     scores_out with fp32 REDs (the caller zeroes it) — fills the SMs when M x N is small
  const int64_t* seg;      // [Q][2]
  unsigned long long* ranks;
  float* scores_out;
  int64_t N;
  int Q, Kd;
  int head;  // 1: the positive is the head (head-batch)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128-byte swizzle: rows are 128 B, 8-row atoms are 1024 B apart (SBO); LBO is unused.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                     // leading byte offset (ignored), bits [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (Blackwell), bits [46,48)
  d |= static_cast<uint64_t>(2) << 61;                     // SWIZZLE_128B, bits [61,64)
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ bool tc_member(const int64_t* __restrict__ m, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = __ldg(m + mid);
    if (v < x) lo = mid + 1;
    else if (v > x) hi = mid;
    else return true;
  }
  return false;
}

// TC_STAGES = 2, one CTA per SM (192 KB): the measured default.  TC_STAGES = 1 with TWO CTAs per SM (2 x 96 KB of
// shared memory, 2 x 256 TMEM columns): each CTA runs TMA -> split -> MMA serially, the pair interleaves on the SM's
// tensor pipe, and one CTA's prologue / epilogue overlaps the other's main loop (A/B: KGE_TC_STAGES=1).
template <int TC_STAGES>
__global__ void __launch_bounds__(256, TC_STAGES == 1 ? 2 : 1)
rank_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_e,
               RankTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES] TMA landed
  uint64_t* split = bars + TC_STAGES;       // [STAGES] lo tiles written
  uint64_t* empty = bars + 2 * TC_STAGES;   // [STAGES] MMAs finished reading the stage
  uint64_t* tmem_full = bars + 3 * TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_base = blockIdx.x * TC_M;
  const bool ksplit = p.k_split > 1;
  const int64_t e_base = p.diag ? (int64_t)(blockIdx.x >> 1) * TC_N  // the column tile holding this block's diagonal
                         : ksplit ? (int64_t)blockIdx.y * TC_N       // grid.z is the k-split index
                                 : ((int64_t)blockIdx.y + (int64_t)blockIdx.z * gridDim.y) * TC_N;  // folded over y, z
  if (e_base >= p.N) return;  // the whole CTA, before any barrier / TMEM set-up
  const int all_kb = (p.Kd + TC_K - 1) / TC_K;
  const int kb0 = ksplit ? (int)blockIdx.z * p.kb_per_split : 0;  // this CTA's k-blocks: [kb0, kb0 + num_kb)
  const int num_kb = ksplit ? min(p.kb_per_split, all_kb - kb0) : all_kb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_e)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TC_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % TC_STAGES;
        if (kb >= TC_STAGES) mbar_wait(&empty[s], ((kb / TC_STAGES) - 1) & 1);
        uint8_t* st = smem + s * TC_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], TC_A_BYTES + TC_B_BYTES);
        tma_load_2d(st, &map_q, &full[s], (kb0 + kb) * TC_K, q_base);
        tma_load_2d(st + 2 * TC_A_BYTES, &map_e, &full[s], (kb0 + kb) * TC_K, (int)e_base);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) |
                           ((uint32_t)(TC_M >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % TC_STAGES;
      mbar_wait(&split[s], (kb / TC_STAGES) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a = smem_u32(smem + s * TC_STAGE_BYTES);
        const uint64_t da = umma_desc_k_sw128(a), dal = umma_desc_k_sw128(a + TC_A_BYTES);
        const uint64_t db = umma_desc_k_sw128(a + 2 * TC_A_BYTES);
        const uint64_t dbl = umma_desc_k_sw128(a + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
        for (int k = 0; k < TC_K / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // 32 bytes per K=8 step, in 16-byte units
          umma_tf32(tmem_base, da + adv, db + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_tf32(tmem_base, da + adv, dbl + adv, idesc, 1u);
          umma_tf32(tmem_base, dal + adv, db + adv, idesc, 1u);
        }
        umma_commit(&empty[s]);
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== lo-splitters, then epilogue =====
    const int t = threadIdx.x - 128;  // 0..127
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % TC_STAGES;
      mbar_wait(&full[s], (kb / TC_STAGES) & 1);
      float4* st = reinterpret_cast<float4*>(smem + s * TC_STAGE_BYTES);
      // A (1024 float4) -> A_lo, B (2048 float4) -> B_lo; same (swizzled) position in the sibling tile
#pragma unroll 4
      for (int i = t; i < TC_A_BYTES / 16; i += 128) {
        const float4 x = st[i];
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
        st[i + TC_A_BYTES / 16] = l;
      }
      float4* sb = st + 2 * TC_A_BYTES / 16;
#pragma unroll 4
      for (int i = t; i < TC_B_BYTES / 16; i += 128) {
        const float4 x = sb[i];
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
        sb[i + TC_B_BYTES / 16] = l;
      }
      // generic-proxy writes must be visible to the tensor core's (async-proxy) reads
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&split[s]);
    }

    // ----- epilogue: one query row per thread -----
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int w4 = warp & 3;
    const int row = w4 * 32 + lane;
    const int qi = q_base + row;
    const bool vq = qi < p.Q;
    float sp = 0.f;
    int64_t pos = -1, lo = 0, hi = 0;
    if (vq) {
      sp = p.pos_score[qi];
      pos = p.queries[3 * qi + (p.head ? 0 : 2)];
      lo = p.seg[2 * qi];
      hi = p.seg[2 * qi + 1];
    }
    unsigned cnt = 0;
    if (p.diag) {  // S[qi, qi]: the positive of query qi scored by the same MMAs as every candidate
      const int col = qi - (int)e_base;
      float s_diag = 0.f;
      for (int c = 0; c < TC_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j == col) s_diag = __uint_as_float(v[j]);
      }
      if (vq) p.pos_out[qi] = s_diag;
    }
    // Ranking without a score dump: per 32-column chunk a bit mask of the candidates that beat the positive, from
    // which the FILTERED ones are cleared by walking this query's (sorted) filter segment once across the tile —
    // one lower_bound per tile and thread instead of one binary search per beating candidate (with untrained
    // tables half of all candidates beat the positive: ~100 us of dependent L2 loads per tile, more than the MMAs).
    int64_t cur = lo;
    if (!p.diag && !p.scores_out && vq && hi > lo) {
      int64_t a = lo, b = hi;
      while (a < b) {
        const int64_t mid = (a + b) >> 1;
        if (__ldg(p.filter.members + mid) < e_base) a = mid + 1;
        else b = mid;
      }
      cur = a;
    }
    for (int c = 0; c < ((p.diag || p.scores_out) ? 0 : TC_N / 32); ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(c * 32), v);
      if (vq) {
        const int64_t c0 = e_base + c * 32;
        unsigned beats = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int64_t e = c0 + j;
          const float s = __uint_as_float(v[j]);
          if (e < p.N && e != pos && ((s > sp) || (s == sp && e < pos))) beats |= 1u << j;
        }
        while (cur < hi) {  // filtered entities inside this chunk never count
          const int64_t m = __ldg(p.filter.members + cur);
          if (m >= c0 + 32) break;
          beats &= ~(1u << (int)(m - c0));
          ++cur;
        }
        cnt += __popc(beats);
      }
    }
    // With a score dump: element by element, and the 32 x 32 block of every chunk goes through shared memory so that
    // the global stores (or, with split-K, the fp32 REDs) of a warp are 128-byte rows instead of 32 scattered words
    // (the stage buffers are free: tmem_full says every MMA has finished reading them).
    float* tbuf = reinterpret_cast<float*>(smem) + w4 * (32 * 33);
    for (int c = 0; c < ((p.diag || !p.scores_out) ? 0 : TC_N / 32); ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(c * 32), v);
      __syncwarp();  // the previous chunk's reads of tbuf are done
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int64_t e = e_base + c * 32 + j;
        float outv = 0.f;
        if (vq && e < p.N) {
          const float s = __uint_as_float(v[j]);
          const bool beats = (s > sp) || (s == sp && e < pos);
          bool filtered = false;
          if (e != pos && hi > lo) filtered = tc_member(p.filter.members, lo, hi, e);
          if (beats && e != pos && !filtered) ++cnt;
          outv = ksplit ? s : (filtered ? sp + (-1e5f) : (e == pos ? sp : s));
        }
        tbuf[lane * 33 + j] = outv;
      }
      __syncwarp();
      const int64_t e = e_base + c * 32 + lane;  // lane = column now
      if (e < p.N) {
        for (int r = 0; r < 32; ++r) {
          const int qr = q_base + w4 * 32 + r;
          if (qr >= p.Q) break;
          float* dst = p.scores_out + (int64_t)qr * p.N + e;
          if (ksplit) atomicAdd(dst, tbuf[r * 33 + lane]);  // partial sum of this k-split (RED.ADD.F32)
          else *dst = tbuf[r * 33 + lane];
        }
      }
    }
    if (vq && cnt) atomicAdd(p.ranks + qi, (unsigned long long)cnt);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_N)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<EncodeTiledFn>(fn);
}

static bool make_map(EncodeTiledFn enc, CUtensorMap* map, const float* base, int64_t rows, int kd, int box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)kd, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kd * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)TC_K, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns KGE_OK when the tensor-core kernel was launched, KGE_E_UNSUPPORTED when the caller should use
// the fp32 tile kernel instead (shape/alignment not eligible, or disabled with KGE_RANK_TC=0).
bool rank_tc_eligible(const float* ent, int kd, int64_t n_entity) {
  if (const char* e = getenv("KGE_RANK_TC"))
    if (atoi(e) == 0) return false;
  static EncodeTiledFn enc = encode_fn();
  return enc && kd % 4 == 0 && aligned16(ent) && n_entity <= INT32_MAX;
}

int rank_tc_launch(const float* qmat, const float* ent, int64_t n_entity, int kd, const int64_t* queries, int Q,
                   const kge_filter_csr_t* filter, bool has_filter, float* pos_score, const int64_t* seg,
                   unsigned long long* ranks, float* scores_out, bool head, cudaStream_t st, const float* posrows,
                   int k_split) {
  if (!rank_tc_eligible(ent, kd, n_entity) || !aligned16(qmat)) return KGE_E_UNSUPPORTED;
  static EncodeTiledFn enc = encode_fn();
  CUtensorMap mq, me, mp;
  if (!make_map(enc, &mq, qmat, Q, kd, TC_M) || !make_map(enc, &me, ent, n_entity, kd, TC_N))
    return KGE_E_UNSUPPORTED;
  if (posrows && (!aligned16(posrows) || !make_map(enc, &mp, posrows, Q, kd, TC_N))) return KGE_E_UNSUPPORTED;
  RankTcParams p{};
  p.queries = queries;
  if (has_filter) p.filter = *filter;
  p.pos_score = pos_score;
  p.seg = seg;
  p.ranks = ranks;
  p.scores_out = scores_out;
  p.N = n_entity;
  p.Q = Q;
  p.Kd = kd;
  p.head = head ? 1 : 0;
  static const int stages = (getenv("KGE_TC_STAGES") && atoi(getenv("KGE_TC_STAGES")) == 1) ? 1 : 2;
  auto kern = stages == 1 ? rank_tc_kernel<1> : rank_tc_kernel<2>;
  const int smem_bytes = tc_smem_bytes(stages);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return (int)e;
  if (posrows) {  // diag pass: positives re-scored by the tensor cores, one CTA per 128 queries
    RankTcParams pd = p;
    pd.diag = 1;
    pd.N = Q;
    pd.pos_out = pos_score;
    kern<<<dim3((unsigned)((Q + TC_M - 1) / TC_M)), 256, smem_bytes, st>>>(mq, mp, pd);
    KGE_LAUNCH_CHECK();
  }
  const int64_t e_tiles = (n_entity + TC_N - 1) / TC_N, ty = e_tiles < 32768 ? e_tiles : 32768;
  dim3 grid((unsigned)((Q + TC_M - 1) / TC_M), (unsigned)ty, (unsigned)((e_tiles + ty - 1) / ty));
  if (k_split > 1 && scores_out && !posrows && grid.z == 1) {  // split-K: the caller has zeroed scores_out
    const int all_kb = (kd + TC_K - 1) / TC_K;
    p.kb_per_split = (all_kb + k_split - 1) / k_split;
    p.k_split = (all_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
    if (p.k_split > 1) grid.z = (unsigned)p.k_split;
    else p.k_split = 0;
  }
  kern<<<grid, 256, smem_bytes, st>>>(mq, me, p);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

}  // namespace kge
