// Shared device helpers for the KGE kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kge_b200.h"

#define KGE_LAUNCH_CHECK()                          \
  do {                                              \
    cudaError_t _e = cudaGetLastError();            \
    if (_e != cudaSuccess) return (int)_e;          \
  } while (0)

namespace kge {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

// ----------------------------------------------------------------------------------------------
// Model traits.  Entity rows hold NC components of D floats ([re|im] for the complex models),
// relation rows RC components (mkb/models/{transe,distmult,complex,rotate}.py constructors).
// ----------------------------------------------------------------------------------------------
template <int M>
struct Traits {
  static constexpr int NC = (M == KGE_COMPLEX || M == KGE_ROTATE) ? 2 : 1;
  static constexpr int RC = (M == KGE_COMPLEX) ? 2 : 1;
  static constexpr bool kDistance = (M == KGE_TRANSE || M == KGE_ROTATE);  // score = gamma - acc
  static constexpr bool kPhase = (M == KGE_PROTATE);  // score = gamma - modulus * acc, rows are phases * pd
};

__host__ __device__ inline int entity_comps(int model) {
  return (model == KGE_COMPLEX || model == KGE_ROTATE) ? 2 : 1;  // pRotatE rows are plain D-vectors
}
__host__ __device__ inline int relation_comps(int model) { return model == KGE_COMPLEX ? 2 : 1; }

// ----------------------------------------------------------------------------------------------
// Reductions
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
// Block-wide sum, result broadcast to every thread.  red: __shared__ float[33].  Fixed tree =>
// deterministic for a given block size.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect red from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    x = warp_sum(x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float x = lane < (int)(blockDim.x >> 5) ? red[lane] : -INFINITY;
    x = warp_max(x);
    if (lane == 0) red[32] = x;
  }
  __syncthreads();
  return red[32];
}

// ----------------------------------------------------------------------------------------------
// Vector access.  VEC = 4: 16-byte aligned rows (LDG.128 / RED.128); VEC = 1: any shape.
// ----------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld_global(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else {
    v[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void ld_shared(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 x = *reinterpret_cast<const float4*>(p);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void st_shared(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *p = v[0];
  }
}
// Fire-and-forget global float add (RED, no return value).  VEC = 4 -> one 16-byte vector
// reduction (red.global.add.v4.f32, sm_90+).
template <int VEC>
__device__ __forceinline__ void red_add(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]),
                 "f"(v[1]), "f"(v[2]), "f"(v[3])
                 : "memory");
  } else {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
  }
}

// Same reduction at system scope: the target row may live in a PEER GPU's HBM (row-sharded tables,
// K7), where the add is performed by the owner's L2 over NVLink.  scalar != 0 issues four 4-byte REDs
// instead of one 16-byte one (a debugging switch for fabrics that would not take the vector form).
template <int VEC>
__device__ __forceinline__ void red_add_sys(float* p, const float (&v)[VEC], int scalar) {
  if constexpr (VEC == 4) {
    if (!scalar) {
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]),
                   "f"(v[1]), "f"(v[2]), "f"(v[3])
                   : "memory");
      return;
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k)
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p + k), "f"(v[k]) : "memory");
}

// ----------------------------------------------------------------------------------------------
// Row-sharded entity table (K7).  Entity e lives on shard e % G at local row e / G (block-cyclic, so
// hub entities spread evenly); every shard is reachable through a base pointer that is either local
// HBM or a peer mapping over NVLink.  Ids are split ONCE into (shard, local row) and carried as
// shard << 32 | row; the unsharded instantiations keep the plain id and compile to the same code as
// before the sharded variants existed.
// ----------------------------------------------------------------------------------------------
template <bool SHARD>
__device__ __forceinline__ int64_t shard_split(int64_t id, unsigned n_shards) {
  if constexpr (SHARD) {
    const unsigned u = (unsigned)id, q = u / n_shards;
    return (int64_t)q | ((int64_t)(u - q * n_shards) << 32);
  } else {
    return id;
  }
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// torch's logsigmoid: min(x,0) - log1p(exp(-|x|))  (mkb/losses/adversarial.py:22,25)
__device__ __forceinline__ float log_sigmoid(float x) {
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid(float x) {
  const float e = expf(-fabsf(x));
  return x >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
}

// ----------------------------------------------------------------------------------------------
// Per-element model algebra.  (r0, r1) is the *effective* relation pair: (r, -) for
// TransE/DistMult, (re_r, im_r) for ComplEx, (cos θ, sin θ) for RotatE.
// The query q is the part of the score that does not depend on the candidate entity e:
//   tail-batch (and positives): q = f(head, relation), candidate = tail
//   head-batch:                 q = f(tail, relation), candidate = head
// Query arithmetic uses explicit _rn intrinsics (no FMA contraction) so each element rounds
// exactly like the reference's separate mul / add / sub ATen ops.
// ----------------------------------------------------------------------------------------------
template <int M>
__device__ __forceinline__ void rel_effective(float r_in0, float r_in1, float phase_div, float& r0,
                                              float& r1) {
  if constexpr (M == KGE_ROTATE) {
    // phase = relation / (embedding_range / pi)   (rotate.py:79-81)
    const float phase = __fdiv_rn(r_in0, phase_div);
    sincosf(phase, &r1, &r0);
  } else if constexpr (M == KGE_PROTATE) {
    r0 = __fdiv_rn(r_in0, phase_div);  // phase_relation   (protate.py:80)
    r1 = 0.f;
  } else {
    r0 = r_in0;
    r1 = r_in1;
  }
}

// pRotatE (protate.py:79-91): every row is turned into a phase x / phase_div first;
//   tail: (ph_h + ph_r) - ph_t'          head: ph_h' + (ph_r - ph_t) = -((ph_t - ph_r) - ph_h')
// so with q = ph_a + ph_r (tail) or ph_a - ph_r (head) the element is |sin(q - ph_e)| in both modes
// (negation is exact and sin is odd).  `pd` (the phase divisor) is only read by that model.
template <int M, bool HEAD>
__device__ __forceinline__ void make_query(float a0, float a1, float r0, float r1, float& q0,
                                           float& q1, float pd = 1.f) {
  if constexpr (M == KGE_PROTATE) {
    const float pa = __fdiv_rn(a0, pd);
    q0 = HEAD ? __fsub_rn(pa, r0) : __fadd_rn(pa, r0);
    q1 = 0.f;
  } else if constexpr (M == KGE_TRANSE) {
    // tail: (h + r) - t'   head: h' + (r - t) = h' - (t - r)      (transe.py:70-73)
    q0 = HEAD ? __fsub_rn(a0, r0) : __fadd_rn(a0, r0);
    q1 = 0.f;
  } else if constexpr (M == KGE_DISTMULT) {
    q0 = __fmul_rn(a0, r0);  // (h*r)*t'  |  h'*(r*t)                (distmult.py:68-71)
    q1 = 0.f;
  } else {
    if constexpr (!HEAD) {  // a = head: a∘r                     (complex.py:80-82, rotate.py:90-91)
      q0 = __fsub_rn(__fmul_rn(a0, r0), __fmul_rn(a1, r1));
      q1 = __fadd_rn(__fmul_rn(a0, r1), __fmul_rn(a1, r0));
    } else {  // a = tail: conj(r)∘a                             (complex.py:75-76, rotate.py:84-85)
      q0 = __fadd_rn(__fmul_rn(r0, a0), __fmul_rn(r1, a1));
      q1 = __fsub_rn(__fmul_rn(r0, a1), __fmul_rn(r1, a0));
    }
  }
}

// One element's contribution to the reduction over the hidden dim.
template <int M>
__device__ __forceinline__ float cand_term(float q0, float q1, float e0, float e1, float pd = 1.f) {
  if constexpr (M == KGE_PROTATE) {
    return fabsf(sinf(__fsub_rn(q0, __fdiv_rn(e0, pd))));
  } else if constexpr (M == KGE_TRANSE) {
    return fabsf(e0 - q0);
  } else if constexpr (M == KGE_DISTMULT) {
    return q0 * e0;
  } else if constexpr (M == KGE_COMPLEX) {
    return fmaf(q1, e1, q0 * e0);
  } else {
    const float dx = q0 - e0, dy = q1 - e1;
    return sqrt_approx(fmaf(dx, dx, dy * dy));
  }
}

template <int M>
__device__ __forceinline__ float finish_score(float acc, float gamma, float modulus = 1.f) {
  if constexpr (Traits<M>::kPhase) return __fsub_rn(gamma, __fmul_rn(acc, modulus));  // protate.py:91
  return Traits<M>::kDistance ? gamma - acc : acc;
}

// Backward of one element given c = dL/dscore: ge = c * ds/de (goes to the candidate's row),
// dq += c * ds/dq.  pRotatE: the caller passes c already multiplied by the modulus.
template <int M>
__device__ __forceinline__ void cand_bwd(float q0, float q1, float e0, float e1, float c, float& ge0,
                                         float& ge1, float& dq0, float& dq1, float pd = 1.f) {
  if constexpr (M == KGE_PROTATE) {
    // s = gamma - m |sin x|, x = q - e / pd:  ds/dx = -m sign(sin x) cos x
    float sn, cs;
    sincosf(__fsub_rn(q0, __fdiv_rn(e0, pd)), &sn, &cs);
    const float g = (sn > 0.f) ? c * cs : ((sn < 0.f) ? -c * cs : 0.f);
    ge0 = __fdiv_rn(g, pd);  // dx/de = -1/pd
    ge1 = 0.f;
    dq0 -= g;
  } else if constexpr (M == KGE_TRANSE) {
    const float x = e0 - q0;
    const float sg = (x > 0.f) ? c : ((x < 0.f) ? -c : 0.f);  // c * sign(x), sign(0) = 0
    ge0 = -sg;  // s = gamma - |e - q|
    ge1 = 0.f;
    dq0 += sg;
  } else if constexpr (M == KGE_DISTMULT) {
    ge0 = c * q0;
    ge1 = 0.f;
    dq0 = fmaf(c, e0, dq0);
  } else if constexpr (M == KGE_COMPLEX) {
    ge0 = c * q0;
    ge1 = c * q1;
    dq0 = fmaf(c, e0, dq0);
    dq1 = fmaf(c, e1, dq1);
  } else {
    const float dx = q0 - e0, dy = q1 - e1;
    const float m2 = fmaf(dx, dx, dy * dy);
    const float cim = (m2 > 0.f) ? c * rsqrt_approx(m2) : 0.f;  // c / |x|, 0 at |x| = 0
    ge0 = cim * dx;  // s = gamma - |q - e|  =>  ds/de = +x/|x|, ds/dq = -x/|x|
    ge1 = cim * dy;
    dq0 -= ge0;
    dq1 -= ge1;
  }
}

// Chain dq through q = make_query(a, r): da (fixed entity row), dr (effective relation pair).
template <int M, bool HEAD>
__device__ __forceinline__ void query_bwd(float dq0, float dq1, float a0, float a1, float r0, float r1,
                                          float& da0, float& da1, float& dr0, float& dr1, float pd = 1.f) {
  if constexpr (M == KGE_PROTATE) {  // q = a / pd +- ph_r
    da0 = __fdiv_rn(dq0, pd);
    da1 = 0.f;
    dr0 = HEAD ? -dq0 : dq0;
    dr1 = 0.f;
  } else if constexpr (M == KGE_TRANSE) {
    da0 = dq0;
    da1 = 0.f;
    dr0 = HEAD ? -dq0 : dq0;
    dr1 = 0.f;
  } else if constexpr (M == KGE_DISTMULT) {
    da0 = dq0 * r0;
    da1 = 0.f;
    dr0 = dq0 * a0;
    dr1 = 0.f;
  } else {
    if constexpr (!HEAD) {  // q0 = a0 r0 - a1 r1 ; q1 = a0 r1 + a1 r0
      da0 = fmaf(dq0, r0, dq1 * r1);
      da1 = fmaf(dq1, r0, -dq0 * r1);
      dr0 = fmaf(dq0, a0, dq1 * a1);
      dr1 = fmaf(dq1, a0, -dq0 * a1);
    } else {  // q0 = r0 a0 + r1 a1 ; q1 = r0 a1 - r1 a0
      da0 = fmaf(dq0, r0, -dq1 * r1);
      da1 = fmaf(dq0, r1, dq1 * r0);
      dr0 = fmaf(dq0, a0, dq1 * a1);
      dr1 = fmaf(dq0, a1, -dq1 * a0);
    }
  }
}

// Effective-relation gradient -> stored relation row gradient.
template <int M>
__device__ __forceinline__ void rel_bwd(float dr0, float dr1, float r0, float r1, float phase_div,
                                        float& g0, float& g1) {
  if constexpr (M == KGE_ROTATE) {
    // (r0, r1) = (cos θ, sin θ): dθ = -sin θ * dcos + cos θ * dsin ; θ = r / phase_div
    g0 = __fdiv_rn(fmaf(r0, dr1, -r1 * dr0), phase_div);
    g1 = 0.f;
  } else if constexpr (M == KGE_PROTATE) {
    g0 = __fdiv_rn(dr0, phase_div);  // ph_r = r / pd
    g1 = 0.f;
  } else {
    g0 = dr0;
    g1 = dr1;
  }
}

// ----------------------------------------------------------------------------------------------
// Self-adversarial terms of ONE positive from its K candidate scores (CTA-wide; every thread
// calls it).  sc may live in shared or global memory.  (mkb/losses/adversarial.py:24-26)
//   returns  nterm = sum_j softmax_j(alpha n)_j * logsigmoid(-n_j)      (same value in all threads)
//   writes   coef[j] = cscale * softmax_j * sigmoid(n_j)   when coef != nullptr
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float adv_row_terms(const float* sc, int K, float alpha, float cscale,
                                               float* coef, float* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  float mx = -INFINITY;
  for (int j = tid; j < K; j += nt) mx = fmaxf(mx, alpha * sc[j]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int j = tid; j < K; j += nt) se += expf(alpha * sc[j] - mx);
  se = block_sum(se, red);
  const float inv = 1.f / se;
  float acc = 0.f;
  for (int j = tid; j < K; j += nt) {
    const float n = sc[j];
    const float a = expf(alpha * n - mx) * inv;
    acc = fmaf(a, log_sigmoid(-n), acc);
    if (coef) coef[j] = cscale * a * sigmoid(n);
  }
  return block_sum(acc, red);
}

// Last-CTA fold of the per-positive partial sums [3][B] in a fixed order => deterministic loss.
// Every thread of every CTA calls it after its own partials are written by thread 0.
__device__ __forceinline__ void fold_partials(const float* partials, int B, unsigned int* ticket,
                                              unsigned int n_ctas, float* stats, float* red) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == n_ctas - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int k = threadIdx.x; k < B; k += blockDim.x) {
    s0 += __ldcg(partials + k);
    s1 += __ldcg(partials + B + k);
    s2 += __ldcg(partials + 2 * B + k);
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    stats[0] = s0;
    stats[1] = s1;
    stats[2] = s2;
    stats[3] = -(s0 + s1) / (2.f * s2);  // (positive_loss + negative_loss) / 2, adversarial.py:28-30
    *ticket = 0u;
  }
}

inline float host_phase_div(float embedding_range) {
  // torch: relation / (self.embedding_range.item() / self.pi): the divisor is a Python double
  // that ATen casts to the tensor's dtype (fp32) before the division.
  return (float)((double)embedding_range / 3.141592653589793);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace kge
