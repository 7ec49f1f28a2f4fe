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

// Raised by any TMA kernel whose bounded mbarrier wait gave up (kge_tma_fail_flag() reads and clears it).
__device__ int g_tma_fail = 0;

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
  if (!ok) {
    atomicOr(&g_tma_fail, 1);
    if (fail) atomicOr(fail, 1);
  }
  // 5. last CTA folds the partials in a fixed order (deterministic loss)
  fold_partials(p.partials, p.B, p.ticket, gridDim.x, p.stats, red);
}


// ------------------------------------------------------------------------------------------------
// K3-TMA: the fused backward with BOTH directions of the candidate-row traffic on the TMA:
//   in : candidate row global -> shared memory by cp.async.bulk (as K2-TMA);
//   out: the row's gradient is formed in a shared-memory staging row and added into the dense gradient table
//        by ONE bulk reduction, cp.reduce.async.bulk.global.shared::cta.add.f32 (8 KB per instruction), instead
//        of 16 RED.128 per lane.  ncu shows the scatter kernel limited by the L1/TEX pipe (85 %): LDG.128 and
//        RED.128 both go through it; bulk copies and bulk reductions do not.
// One warp per candidate row (its lanes keep the query and the running query gradient of their columns in
// registers); persistent CTAs walk the work items (K-slice, positive) slice-major, so the band sweep of the
// scatter kernel (id-sorted negatives: all CTAs touch the same band of table + gradient at a time) is kept.
// After its rows, a warp parks its partial dq in its staging row; the CTA folds the 8 partials and runs the
// same chain rule as score_bwd_kernel (positive term, head / tail / relation rows, vector REDs).
// Unchunked, unsharded, 16-byte-vector path only (single-GPU and all-reduce flows).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_bulk_reduce_add_f32(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int M, bool HEAD, int UMAX>
__global__ void __launch_bounds__(kThreads, 1) score_bwd_tma_kernel(BwdParams p, int stages, int n_slices, int* fail) {
  using T = Traits<M>;
  constexpr int NC = T::NC;
  extern __shared__ __align__(16) float smem[];
  const int Dp = (p.D + 3) & ~3;
  const int Kc = (p.k_per_cta + 3) & ~3;
  const uint32_t row_bytes = (uint32_t)p.ent_stride * 4u;
  float* q = smem;                                          // [NC][Dp]
  float* s_coef = q + NC * Dp;                              // [Kc]
  int* s_ids = reinterpret_cast<int*>(s_coef + Kc);         // [Kc]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ids + Kc);  // [kWarps][stages]
  char* ring = reinterpret_cast<char*>(bars + kWarps * stages);  // in : [kWarps][stages][row_bytes]
  char* outr = ring + (size_t)kWarps * stages * row_bytes;       // out: [kWarps][row_bytes]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kWarps * stages) tma_mbar_init(tma_smem_u32(bars + tid), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const uint32_t my_bars = tma_smem_u32(bars + warp * stages);
  char* my_ring = ring + (size_t)warp * stages * row_bytes;
  const uint32_t my_ring_u32 = tma_smem_u32(my_ring);
  float* my_out = reinterpret_cast<float*>(outr + (size_t)warp * row_bytes);
  const uint32_t my_out_u32 = tma_smem_u32(my_out);
  unsigned it0 = 0;
  bool ok = true;

  float scale = 1.f;
  if (p.stats) scale = (p.grad_loss ? __ldg(p.grad_loss) : 1.f) / (2.f * __ldg(p.stats + 2));

  const int64_t n_items = (int64_t)p.B * n_slices;
  for (int64_t wi = blockIdx.x; wi < n_items; wi += gridDim.x) {
    const int slice = (int)(wi / p.B);
    const int64_t i = wi - (int64_t)slice * p.B;
    const int64_t hid = p.sample[3 * i + 0], rid = p.sample[3 * i + 1], tidx = p.sample[3 * i + 2];
    const float* hrow = p.ent + hid * (int64_t)p.ent_stride;
    const float* trow = p.ent + tidx * (int64_t)p.ent_stride;
    const float* fixed = HEAD ? trow : hrow;
    const float* relrow = p.rel + rid * (int64_t)p.rel_stride;
    const int j0 = slice * p.k_per_cta;
    const int j1 = min(p.K, j0 + p.k_per_cta);
    const int nrows = j1 - j0;
    const bool do_pos = (p.gpos != nullptr) && slice == 0;
    const float cpos = do_pos ? scale * __ldg(p.gpos + i) : 0.f;

    for (int k = tid; k < nrows; k += kThreads) {
      s_ids[k] = (int)p.neg[i * (int64_t)p.K + j0 + k];
      s_coef[k] = scale * p.gneg[i * (int64_t)p.K + j0 + k];
    }
    // query of this positive -> shared memory (thread t: columns 4t..4t+3), as in score_bwd_kernel
    const int d = tid * 4;
    const bool active = d < p.D;
    float a0[4] = {}, a1[4] = {}, r0[4] = {}, r1[4] = {};
    if (active) {
      float rr0[4], rr1[4] = {}, q0[4], q1[4];
      ld_global<4>(fixed + d, a0);
      if constexpr (NC == 2) ld_global<4>(fixed + p.im_off + d, a1);
      ld_global<4>(relrow + d, rr0);
      if constexpr (T::RC == 2) ld_global<4>(relrow + p.im_off + d, rr1);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        rel_effective<M>(rr0[v], rr1[v], p.phase_div, r0[v], r1[v]);
        make_query<M, HEAD>(a0[v], a1[v], r0[v], r1[v], q0[v], q1[v], p.phase_div);
      }
      st_shared<4>(q + d, q0);
      if constexpr (NC == 2) st_shared<4>(q + Dp + d, q1);
    }
    __syncthreads();

    // this lane's query columns -> registers; running dq of the same columns
    float qr0[UMAX][4], qr1[UMAX][4], dq0[UMAX][4], dq1[UMAX][4];
#pragma unroll
    for (int u = 0; u < UMAX; ++u) {
      const int dd = (u * 32 + lane) * 4;
#pragma unroll
      for (int v = 0; v < 4; ++v) qr0[u][v] = qr1[u][v] = dq0[u][v] = dq1[u][v] = 0.f;
      if (dd < p.D) {
        ld_shared<4>(q + dd, qr0[u]);
        if constexpr (NC == 2) ld_shared<4>(q + Dp + dd, qr1[u]);
      }
    }

    // candidate rows of this slice: warp w takes k = w, w + 8, ...
    const int n = nrows > warp ? (nrows - warp + kWarps - 1) / kWarps : 0;
    if (lane == 0) {
      for (int s = 0; s < stages && s < n; ++s) {
        const unsigned st = (it0 + s) % stages;
        tma_mbar_expect_tx(my_bars + st * 8, row_bytes);
        tma_bulk_load(my_ring_u32 + st * row_bytes, p.ent + (int64_t)s_ids[warp + s * kWarps] * p.ent_stride, row_bytes,
                      my_bars + st * 8);
      }
    }
    for (int m = 0; m < n; ++m) {
      const unsigned it = it0 + m, st = it % stages;
      const int k = warp + m * kWarps;
      if (ok) ok = tma_mbar_wait(my_bars + st * 8, (it / stages) & 1u);
      const float* row = reinterpret_cast<const float*>(my_ring + (size_t)st * row_bytes);
      const float c = s_coef[k];
      if (lane == 0) tma_wait_group_read0();  // the previous bulk reduction has finished reading the staging row
      __syncwarp();
#pragma unroll
      for (int u = 0; u < UMAX; ++u) {
        const int dd = (u * 32 + lane) * 4;
        if (dd < p.D) {
          float e0[4], e1[4] = {}, g0[4], g1[4];
          ld_shared<4>(row + dd, e0);
          if constexpr (NC == 2) ld_shared<4>(row + p.D + dd, e1);
#pragma unroll
          for (int v = 0; v < 4; ++v)
            cand_bwd<M>(qr0[u][v], qr1[u][v], e0[v], e1[v], c, g0[v], g1[v], dq0[u][v], dq1[u][v], p.phase_div);
          st_shared<4>(my_out + dd, g0);
          if constexpr (NC == 2) st_shared<4>(my_out + p.D + dd, g1);
        }
      }
      tma_fence_proxy_async();  // generic-proxy writes of the staging row -> visible to the async proxy
      __syncwarp();
      if (lane == 0) {
        tma_bulk_reduce_add_f32(p.grad_ent + (int64_t)s_ids[k] * p.g_ent_stride, my_out_u32, row_bytes);
        tma_commit_group();
        if (m + stages < n) {
          tma_mbar_expect_tx(my_bars + st * 8, row_bytes);
          tma_bulk_load(my_ring_u32 + st * row_bytes, p.ent + (int64_t)s_ids[warp + (m + stages) * kWarps] * p.ent_stride,
                        row_bytes, my_bars + st * 8);
        }
      }
    }
    it0 += (unsigned)n;

    // park this warp's partial dq in its staging row, fold the 8 partials per column
    if (lane == 0) tma_wait_group_read0();
    __syncwarp();
#pragma unroll
    for (int u = 0; u < UMAX; ++u) {
      const int dd = (u * 32 + lane) * 4;
      if (dd < p.D) {
        st_shared<4>(my_out + dd, dq0[u]);
        if constexpr (NC == 2) st_shared<4>(my_out + p.D + dd, dq1[u]);
      }
    }
    __syncthreads();
    if (active) {
      float s0[4] = {}, s1[4] = {};
      for (int w = 0; w < kWarps; ++w) {
        const float* src = reinterpret_cast<const float*>(outr + (size_t)w * row_bytes);
        float x0[4], x1[4] = {};
        ld_shared<4>(src + d, x0);
        if constexpr (NC == 2) ld_shared<4>(src + p.D + d, x1);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          s0[v] += x0[v];
          s1[v] += x1[v];
        }
      }
      // ---- positive term + chain rule to the positive's own rows (same algebra as score_bwd_kernel)
      float gt0[4] = {}, gt1[4] = {}, gh0[4] = {}, gh1[4] = {}, gr0[4] = {}, gr1[4] = {};
      float dqp0[4] = {}, dqp1[4] = {}, h0[4] = {}, h1[4] = {};
      if (do_pos) {
        float t0[4], t1[4] = {};
        ld_global<4>(trow + d, t0);
        if constexpr (NC == 2) ld_global<4>(trow + p.im_off + d, t1);
        if constexpr (HEAD) {
          ld_global<4>(hrow + d, h0);
          if constexpr (NC == 2) ld_global<4>(hrow + p.im_off + d, h1);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            float qp0, qp1;
            make_query<M, false>(h0[v], h1[v], r0[v], r1[v], qp0, qp1, p.phase_div);
            cand_bwd<M>(qp0, qp1, t0[v], t1[v], cpos, gt0[v], gt1[v], dqp0[v], dqp1[v], p.phase_div);
          }
        } else {
          float q0[4], q1[4] = {};
          ld_shared<4>(q + d, q0);
          if constexpr (NC == 2) ld_shared<4>(q + Dp + d, q1);
#pragma unroll
          for (int v = 0; v < 4; ++v)
            cand_bwd<M>(q0[v], q1[v], t0[v], t1[v], cpos, gt0[v], gt1[v], s0[v], s1[v], p.phase_div);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float da0, da1, dr0, dr1;
        query_bwd<M, HEAD>(s0[v], s1[v], a0[v], a1[v], r0[v], r1[v], da0, da1, dr0, dr1, p.phase_div);
        if constexpr (HEAD) {
          gt0[v] += da0;
          gt1[v] += da1;
          if (do_pos) {
            float dh0, dh1, dpr0, dpr1;
            query_bwd<M, false>(dqp0[v], dqp1[v], h0[v], h1[v], r0[v], r1[v], dh0, dh1, dpr0, dpr1, p.phase_div);
            gh0[v] = dh0;
            gh1[v] = dh1;
            dr0 += dpr0;
            dr1 += dpr1;
          }
        } else {
          gh0[v] = da0;
          gh1[v] = da1;
        }
        rel_bwd<M>(dr0, dr1, r0[v], r1[v], p.phase_div, gr0[v], gr1[v]);
      }
      float* gh = p.grad_ent + hid * (int64_t)p.g_ent_stride;
      float* gt = p.grad_ent + tidx * (int64_t)p.g_ent_stride;
      float* gr = p.grad_rel + rid * (int64_t)p.g_rel_stride;
      if (!HEAD || do_pos) {
        red_add<4>(gh + d, gh0);
        if constexpr (NC == 2) red_add<4>(gh + p.g_im_off + d, gh1);
      }
      if (HEAD || do_pos) {
        red_add<4>(gt + d, gt0);
        if constexpr (NC == 2) red_add<4>(gt + p.g_im_off + d, gt1);
      }
      red_add<4>(gr + d, gr0);
      if constexpr (T::RC == 2) red_add<4>(gr + p.g_im_off + d, gr1);
    }
    __syncthreads();  // q, ids, coefficients and the staging rows are rewritten by the next item
  }
  if (lane == 0) tma_wait_group0();  // every bulk reduction of this warp has been performed
  if (!ok) {
    atomicOr(&g_tma_fail, 2);
    if (fail) atomicOr(fail, 1);
  }
}

}  // namespace kge
