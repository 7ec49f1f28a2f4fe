// K2-TMA: the fused gather -> score -> adversarial-loss forward with candidate rows staged through the TMA
// (cp.async.bulk global -> shared memory, completion on an mbarrier) instead of per-lane LDG.128.
// Included by score.cu (shares FwdParams and the device helpers); selected by kge_fused_fwd when
// KGE_FWD_TMA is set — see DESIGN.md §4 for the A/B against the LDG kernel and which one is the default.
//
// Why a second kernel: ncu shows the LDG forward limited by the L1/TEX pipe (83 %), which carries the
// global row loads AND the shared-memory reads of the query.  Here
//   * a candidate row ([re|im] = NC*D floats, one contiguous 4-8 KB segment) arrives with ONE bulk copy
//     issued by one lane — no per-lane address generation, no L1 tag stage, no registers held by loads in
//     flight; each warp owns a ring of `stages` row buffers, so `8 * stages` rows (up to 192 KB) are in
//     flight per SM independent of occupancy;
//   * the query lives in REGISTERS (each lane always handles the same hidden-dim columns), so the only
//     shared-memory traffic per scored row is the row itself: 8 KB instead of 8 KB (LDG) + 4 KB (query);
//   * the kernel is persistent: grid = resident CTAs, each walks positives i = blockIdx.x, += gridDim.x.
// Same arithmetic, in the same order per lane, as score_neg_kernel<M, HEAD, 4, true>: scores are
// bit-identical (tested).
#pragma once

namespace kge {

__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a copy that never completes (a bug, not a steady-state event) must not hang the GPU box.
__device__ __forceinline__ bool tma_mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
// One contiguous global segment -> shared memory; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// UMAX = ceil(D / 128): 16-byte column chunks per lane.  MINB = resident CTAs per SM the kernel is compiled
// for (1: up to 255 registers; 2: 128).
template <int M, bool HEAD, int UMAX, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) score_neg_tma_kernel(FwdParams p, int stages, int* fail) {
  using T = Traits<M>;
  constexpr int NC = T::NC;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[33];
  const int Kp = (p.K + 3) & ~3;
  float* q = smem;                                  // [NC][Dp]
  float* sc = q + NC * p.Dp;                        // [Kp]
  int* ids = reinterpret_cast<int*>(sc + Kp);       // [Kp]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ids + Kp);  // [kWarps][stages]
  const uint32_t row_bytes = (uint32_t)p.ent_stride * 4u;
  char* ring = reinterpret_cast<char*>(bars + kWarps * stages);  // [kWarps][stages][row_bytes]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kWarps * stages) tma_mbar_init(tma_smem_u32(bars + tid), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const uint32_t my_bars = tma_smem_u32(bars + warp * stages);
  char* my_ring = ring + (size_t)warp * stages * row_bytes;
  const uint32_t my_ring_u32 = tma_smem_u32(my_ring);
  unsigned it0 = 0;  // rows this warp has consumed so far (stage = it % stages, parity = (it / stages) & 1)
  bool ok = true;

  for (int64_t i = blockIdx.x; i < p.B; i += gridDim.x) {
    const int64_t hid = p.sample[3 * i + 0], rid = p.sample[3 * i + 1], tidx = p.sample[3 * i + 2];
    const float* fixed = p.ent + (HEAD ? tidx : hid) * (int64_t)p.ent_stride;
    const float* relrow = p.rel + rid * (int64_t)p.rel_stride;
    const int64_t* negrow = p.neg + i * (int64_t)p.K;
    for (int j = tid; j < p.K; j += kThreads) ids[j] = (int)negrow[j];

    // 1. query -> shared memory and the positive's score (same code as score_neg_kernel)
    float pacc = 0.f;
    for (int d = tid * 4; d < p.D; d += kThreads * 4) {
      float a0[4], a1[4] = {}, rr0[4], rr1[4] = {}, q0[4], q1[4];
      ld_global<4>(fixed + d, a0);
      if constexpr (NC == 2) ld_global<4>(fixed + p.D + d, a1);
      ld_global<4>(relrow + d, rr0);
      if constexpr (T::RC == 2) ld_global<4>(relrow + p.D + d, rr1);
      float r0[4], r1[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        rel_effective<M>(rr0[v], rr1[v], p.phase_div, r0[v], r1[v]);
        make_query<M, HEAD>(a0[v], a1[v], r0[v], r1[v], q0[v], q1[v], p.phase_div);
      }
      st_shared<4>(q + d, q0);
      if constexpr (NC == 2) st_shared<4>(q + p.Dp + d, q1);
      const float* hrow = p.ent + hid * (int64_t)p.ent_stride;
      const float* trow = p.ent + tidx * (int64_t)p.ent_stride;
      float t0[4], t1[4] = {};
      ld_global<4>(trow + d, t0);
      if constexpr (NC == 2) ld_global<4>(trow + p.D + d, t1);
      if constexpr (HEAD) {
        float h0[4], h1[4] = {};
        ld_global<4>(hrow + d, h0);
        if constexpr (NC == 2) ld_global<4>(hrow + p.D + d, h1);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          float qp0, qp1;
          make_query<M, false>(h0[v], h1[v], r0[v], r1[v], qp0, qp1, p.phase_div);
          pacc += cand_term<M>(qp0, qp1, t0[v], t1[v], p.phase_div);
        }
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) pacc += cand_term<M>(q0[v], q1[v], t0[v], t1[v], p.phase_div);
      }
    }
    const float pos = finish_score<M>(block_sum(pacc, red), p.gamma, load_modulus<M>(p));
    if (tid == 0 && p.pos_score) p.pos_score[i] = pos;
    __syncthreads();  // q, ids visible

    // 2. this lane's query columns -> registers
    float qr0[UMAX][4], qr1[UMAX][4];
#pragma unroll
    for (int u = 0; u < UMAX; ++u) {
      const int d = (u * 32 + lane) * 4;
#pragma unroll
      for (int v = 0; v < 4; ++v) qr0[u][v] = qr1[u][v] = 0.f;
      if (d < p.D) {
        ld_shared<4>(q + d, qr0[u]);
        if constexpr (NC == 2) ld_shared<4>(q + p.Dp + d, qr1[u]);
      }
    }

    // 3. candidates: warp w scores rows j = w, w + 8, ... through its ring of `stages` row buffers
    const int n = (p.K - warp + kWarps - 1) / kWarps;
    if (lane == 0) {
      for (int s = 0; s < stages && s < n; ++s) {
        const unsigned it = it0 + s, st = it % stages;
        tma_mbar_expect_tx(my_bars + st * 8, row_bytes);
        tma_bulk_load(my_ring_u32 + st * row_bytes, p.ent + (int64_t)ids[warp + s * kWarps] * p.ent_stride, row_bytes,
                      my_bars + st * 8);
      }
    }
    for (int m = 0; m < n; ++m) {
      const unsigned it = it0 + m, st = it % stages;
      if (ok) ok = tma_mbar_wait(my_bars + st * 8, (it / stages) & 1u);
      const float* row = reinterpret_cast<const float*>(my_ring + (size_t)st * row_bytes);
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < UMAX; ++u) {
        const int d = (u * 32 + lane) * 4;
        if (d < p.D) {
          float e0[4], e1[4] = {};
          ld_shared<4>(row + d, e0);
          if constexpr (NC == 2) ld_shared<4>(row + p.D + d, e1);
#pragma unroll
          for (int v = 0; v < 4; ++v) acc += cand_term<M>(qr0[u][v], qr1[u][v], e0[v], e1[v], p.phase_div);
        }
      }
      __syncwarp();  // every lane has read the slot before it is handed back to the TMA
      if (lane == 0 && m + stages < n) {
        tma_mbar_expect_tx(my_bars + st * 8, row_bytes);
        tma_bulk_load(my_ring_u32 + st * row_bytes, p.ent + (int64_t)ids[warp + (m + stages) * kWarps] * p.ent_stride,
                      row_bytes, my_bars + st * 8);
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const int j = warp + m * kWarps;
        const float s = finish_score<M>(acc, p.gamma, load_modulus<M>(p));
        if (p.neg_score) p.neg_score[i * (int64_t)p.K + j] = s;
        sc[j] = s;
      }
    }
    it0 += (unsigned)n;

    // 4. self-adversarial terms of this positive (losses/adversarial.py:22-30)
    __syncthreads();
    const float w = p.weight[i];
    const float nt = adv_row_terms(sc, p.K, p.alpha, w, p.coef_neg + i * (int64_t)p.K, red);
    if (tid == 0) {
      p.coef_pos[i] = -w * sigmoid(-pos);
      p.partials[i] = w * log_sigmoid(pos);
      p.partials[p.B + i] = w * nt;
      p.partials[2 * p.B + i] = w;
    }
    __syncthreads();  // sc / ids / q are rewritten by the next positive
  }
  if (!ok && fail) atomicOr(fail, 1);
  // 5. last CTA folds the partials in a fixed order (deterministic loss)
  fold_partials(p.partials, p.B, p.ticket, gridDim.x, p.stats, red);
}

}  // namespace kge
