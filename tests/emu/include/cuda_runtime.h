// TEST INFRASTRUCTURE — not product code, never linked into libkge_b200.so.
//
// A minimal single-threaded emulation of the CUDA execution model, just large enough to compile
// mkb_b200/csrc/{score,loss,sampler,rank,api}.cu with g++ and run the kernels on host memory:
// every CUDA thread of a block is a ucontext fiber, blocks run one after another,
// __syncthreads / warp shuffles / ballots are rendezvous points between fibers, "global memory" is
// ordinary host memory, atomics are plain read-modify-writes.  It exists so the KERNEL LOGIC
// (indexing, tiling, reductions, the loss algebra, the sharded addressing) can be checked against the
// oracle in the CPU test suite of a container that has no GPU; it says nothing about performance and
// is not a fallback: the package refuses to run without the real CUDA library.
//
// tests/emu/build_emu.py rewrites the three CUDA-only constructs g++ cannot parse — `k<<<...>>>(...)`
// launches, `extern __shared__` arrays and the inline-PTX statements — and compiles with this file
// standing in for <cuda_runtime.h>.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define KGE_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static  // blocks run sequentially, all fibers of a block share one address space

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 {
  float x, y, z, w;
};
struct alignas(8) float2 {
  float x, y;
};
struct alignas(16) uint4 {
  unsigned x, y, z, w;
};
struct alignas(8) uint2 {
  unsigned x, y;
};
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  int multiProcessorCount, major, minor;
};
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = 4;  // a 4-"SM" device keeps the grid-size heuristics small
  return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  p->multiProcessorCount = 4;
  p->major = 10;
  p->minor = 0;
  return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
  return cudaSuccess;
}
inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}

namespace emu {

struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  int warp, lane;
  bool done;
};
struct Warp {
  int live, count;
  unsigned gen;
  unsigned long long slot[32];
};
struct State {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<char> stacks;
  int live = 0, bar_count = 0;
  unsigned bar_gen = 0;
  ucontext_t sched;
  Fiber* cur = nullptr;
  std::function<void()> body;
  uint3 block_idx{0, 0, 0};
  dim3 block_dim, grid_dim;
  std::vector<char> dyn;
  void* dyn_smem = nullptr;
  long launches = 0;
};
inline State S;
constexpr size_t kStack = 256 * 1024;

inline void yield() { swapcontext(&S.cur->ctx, &S.sched); }

inline void release_if_complete() {  // called after a fiber exits: barriers may now be satisfied
  if (S.live > 0 && S.bar_count >= S.live) {
    S.bar_count = 0;
    ++S.bar_gen;
  }
}
inline void entry() {
  S.body();
  Fiber* f = S.cur;
  f->done = true;
  --S.live;
  Warp& w = S.warps[f->warp];
  --w.live;
  if (w.live > 0 && w.count >= w.live) {
    w.count = 0;
    ++w.gen;
  }
  release_if_complete();
  // returning follows uc_link back to the scheduler
}

inline void run_block(int nthreads) {
  const int nw = (nthreads + 31) / 32;
  S.fibers.assign(nthreads, Fiber{});
  S.warps.assign(nw, Warp{});
  if (S.stacks.size() < (size_t)nthreads * kStack) S.stacks.resize((size_t)nthreads * kStack);
  S.live = nthreads;
  S.bar_count = 0;
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = S.fibers[t];
    f.tid = uint3{(unsigned)(t % S.block_dim.x), (unsigned)((t / S.block_dim.x) % S.block_dim.y),
                  (unsigned)(t / (S.block_dim.x * S.block_dim.y))};
    f.warp = t / 32;
    f.lane = t % 32;
    f.done = false;
    ++S.warps[f.warp].live;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = S.stacks.data() + (size_t)t * kStack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &S.sched;
    makecontext(&f.ctx, (void (*)())entry, 0);
  }
  long spins = 0;
  while (S.live > 0) {
    for (int t = 0; t < nthreads; ++t) {
      Fiber& f = S.fibers[t];
      if (f.done) continue;
      S.cur = &f;
      swapcontext(&S.sched, &f.ctx);
    }
    if (++spins > 50000000L) {
      fprintf(stderr, "cuda_emu: block (%u,%u) does not terminate (deadlocked barrier?)\n", S.block_idx.x,
              S.block_idx.y);
      abort();
    }
  }
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, F&& body) {
  S.grid_dim = grid;
  S.block_dim = block;
  S.dyn.assign(smem + 64, 0);
  S.dyn_smem = (void*)(((uintptr_t)S.dyn.data() + 63) & ~(uintptr_t)63);
  S.body = body;
  ++S.launches;
  const int nthreads = (int)(block.x * block.y * block.z);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        S.block_idx = uint3{x, y, z};
        run_block(nthreads);
      }
}

inline void warp_barrier() {
  Warp& w = S.warps[S.cur->warp];
  const unsigned my = w.gen;
  if (++w.count >= w.live) {
    w.count = 0;
    ++w.gen;
  } else {
    while (w.gen == my) yield();
  }
}
template <class T>
inline T warp_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
  Warp& w = S.warps[S.cur->warp];
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.slot[S.cur->lane] = raw;
  warp_barrier();
  raw = w.slot[src_lane & 31];
  warp_barrier();  // nobody overwrites a slot before every lane has read
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}

}  // namespace emu

#define threadIdx (emu::S.cur->tid)
#define blockIdx (emu::S.block_idx)
#define blockDim (emu::S.block_dim)
#define gridDim (emu::S.grid_dim)

inline void __syncthreads() {
  using namespace emu;
  const unsigned my = S.bar_gen;
  if (++S.bar_count >= S.live) {
    S.bar_count = 0;
    ++S.bar_gen;
  } else {
    while (S.bar_gen == my) yield();
  }
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline void __threadfence() {}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  return emu::warp_exchange(v, src);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return emu::warp_exchange(v, emu::S.cur->lane ^ lane_mask);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
  const int src = emu::S.cur->lane + delta;
  return emu::warp_exchange(v, src < 32 ? src : emu::S.cur->lane);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int delta) {
  const int src = emu::S.cur->lane - delta;
  return emu::warp_exchange(v, src >= 0 ? src : emu::S.cur->lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  using namespace emu;
  Warp& w = S.warps[S.cur->warp];
  w.slot[S.cur->lane] = pred ? 1ull : 0ull;
  warp_barrier();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) {
    const int t = S.cur->warp * 32 + l;
    if (t < (int)S.fibers.size() && !S.fibers[t].done && w.slot[l]) m |= 1u << l;
  }
  warp_barrier();
  return m;
}

template <class T>
inline T __ldg(const T* p) {
  return *p;
}
template <class T>
inline T __ldcg(const T* p) {
  return *p;
}
template <class T>
inline T atomicAdd(T* p, T v) {
  const T old = *p;
  *p = old + v;
  return old;
}
template <class T>
inline T atomicOr(T* p, T v) {
  const T old = *p;
  *p = old | v;
  return old;
}
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
using std::max;
using std::min;
