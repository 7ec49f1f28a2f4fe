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
#include <time.h>
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
namespace emu {
inline cudaError_t last_error = 0;  // set by emu::launch when a launch configuration breaks a CUDA limit
}
inline cudaError_t cudaGetLastError() {
  const cudaError_t e = emu::last_error;
  emu::last_error = 0;
  return e;
}
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
namespace emu {
// kernels that opted in to more than 48 KB of dynamic shared memory (cudaFuncAttributeMaxDynamicSharedMemorySize)
inline std::vector<std::pair<const void*, size_t>> smem_optin;
template <class R, class... A>
inline const void* fn_addr(R (*f)(A...)) {  // a kernel name or a variable holding a kernel pointer
  return reinterpret_cast<const void*>(f);
}
inline size_t optin_of(const void* k) {
  size_t best = 0;
  for (auto& e : smem_optin)
    if (e.first == k) best = e.second;
  return best;
}
}  // namespace emu
template <class F>
inline cudaError_t cudaFuncSetAttribute(F f, cudaFuncAttribute, int bytes) {
  if (bytes > 232448) return 1;  // cudaErrorInvalidValue
  emu::smem_optin.push_back({reinterpret_cast<const void*>(f), (size_t)bytes});
  return cudaSuccess;
}
inline const char* cudaGetErrorString(cudaError_t e) {
  return e == 9 ? "invalid configuration argument (emulated)" : "emulated CUDA error";
}
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
  int site[32];  // source line of the warp primitive each lane is waiting in
};
// Schedule perturbation (race check).  Fibers switch only at barriers / shuffles, so a fixed thread order hides
// every missing __syncthreads whose reader happens to run after its writer.  KGE_EMU_ORDER=reverse|random runs
// the threads of a block (and the blocks of a grid) in the opposite / a shuffled order: a kernel without data
// races produces the same result under all three; KGE_EMU_SEED seeds the shuffle.
struct Config {
  int order = 0;  // 0 forward, 1 reverse, 2 random
  unsigned long long rng = 0x9E3779B97F4A7C15ull;
  Config() {
    const char* o = getenv("KGE_EMU_ORDER");
    if (o && !strcmp(o, "reverse")) order = 1;
    if (o && !strcmp(o, "random")) order = 2;
    if (const char* sd = getenv("KGE_EMU_SEED")) rng ^= strtoull(sd, nullptr, 10) * 0xD1342543DE82EF95ull;
  }
  unsigned next() {  // splitmix64
    unsigned long long z = (rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (unsigned)((z ^ (z >> 31)) >> 16);
  }
  void permute(std::vector<int>& v) {
    if (order == 1) std::reverse(v.begin(), v.end());
    if (order == 2)
      for (size_t i = v.size(); i > 1; --i) std::swap(v[i - 1], v[next() % i]);
  }
};
inline Config C;

struct State {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<char> stacks;
  int live = 0, bar_count = 0, bar_site = 0;
  unsigned bar_gen = 0;
  ucontext_t sched;
  Fiber* cur = nullptr;
  std::function<void()> body;
  uint3 block_idx{0, 0, 0};
  dim3 block_dim, grid_dim;
  std::vector<char> dyn;
  void* dyn_smem = nullptr;
  size_t dyn_bytes = 0;
  std::vector<void*> poisoned;  // static __shared__ objects already poisoned for the running block
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
  // shared memory holds garbage when a block starts: poison it so reads of unwritten words show (NaN / -1)
  if (S.dyn_bytes) memset(S.dyn_smem, 0xFF, S.dyn_bytes);
  S.poisoned.clear();
  std::vector<int> order(nthreads);
  for (int t = 0; t < nthreads; ++t) order[t] = t;
  C.permute(order);
  long spins = 0;
  while (S.live > 0) {
    if (C.order == 2 && spins) C.permute(order);
    // random mode: warps (and lanes) also advance at different RATES — half of the warps and a quarter of the
    // remaining lanes sit a round out, so warps drift apart as far as the barriers allow
    unsigned skip_warp = 0, skip_lane = 0;
    if (C.order == 2) {
      skip_warp = C.next();
      skip_lane = C.next() & C.next();
    }
    for (int oi = 0; oi < nthreads; ++oi) {
      const int t = order[oi];
      Fiber& f = S.fibers[t];
      if (f.done) continue;
      if (C.order == 2 && (((skip_warp >> (f.warp & 31)) & 1u) || ((skip_lane >> f.lane) & 1u))) continue;
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
inline void launch(const void* kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t, F&& body) {
  if (smem > 48 * 1024 && optin_of(kernel) < smem) {  // a launch above 48 KB needs the opt-in attribute first
    fprintf(stderr, "cuda_emu: %zu bytes of dynamic shared memory without cudaFuncSetAttribute opt-in\n", smem);
    last_error = 1;
    return;
  }
  // the launch limits of sm_100 (cudaErrorInvalidConfiguration = 9 on the device; here too, and nothing runs)
  const unsigned long long nthreads_ll = (unsigned long long)block.x * block.y * block.z;
  if (grid.x == 0 || grid.y == 0 || grid.z == 0 || nthreads_ll == 0 || nthreads_ll > 1024 || block.x > 1024 ||
      block.y > 1024 || block.z > 64 || grid.x > 2147483647u || grid.y > 65535u || grid.z > 65535u ||
      smem > 232448) {
    fprintf(stderr, "cuda_emu: invalid launch configuration grid=(%u,%u,%u) block=(%u,%u,%u) smem=%zu\n", grid.x,
            grid.y, grid.z, block.x, block.y, block.z, smem);
    last_error = 9;
    return;
  }
  S.grid_dim = grid;
  S.block_dim = block;
  S.dyn.assign(smem + 64, 0);
  S.dyn_smem = (void*)(((uintptr_t)S.dyn.data() + 63) & ~(uintptr_t)63);
  S.dyn_bytes = smem;
  S.body = body;
  ++S.launches;
  const int nthreads = (int)nthreads_ll;
  const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
  if (C.order == 0) {
    for (unsigned z = 0; z < grid.z; ++z)
      for (unsigned y = 0; y < grid.y; ++y)
        for (unsigned x = 0; x < grid.x; ++x) {
          S.block_idx = uint3{x, y, z};
          run_block(nthreads);
        }
  } else {  // blocks of a grid have no guaranteed order either
    std::vector<int> border((size_t)nblocks);
    for (size_t b = 0; b < border.size(); ++b) border[b] = (int)b;
    C.permute(border);
    for (int b : border) {
      S.block_idx = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / (grid.x * grid.y))};
      run_block(nthreads);
    }
  }
}

// static __shared__ objects (build_emu.py appends a call after each declaration): first fiber of a block to
// reach the declaration fills it with 0xFF
inline void poison_shared(void* p, size_t n) {
  for (void* q : S.poisoned)
    if (q == p) return;
  S.poisoned.push_back(p);
  memset(p, 0xFF, n);
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
// Every lane named in `mask` must execute the SAME primitive (undefined behaviour on the device otherwise):
// the calling lane has to be in the mask, every live lane of the warp has to be in it (this emulation
// rendezvouses all live lanes), and all lanes must have come from the same source line.
inline void check_site(unsigned mask, int line) {
  Warp& w = S.warps[S.cur->warp];
  const int wi = S.cur->warp;
  if (!((mask >> S.cur->lane) & 1u)) {
    fprintf(stderr, "cuda_emu: lane %d calls a *_sync primitive (line %d) with mask %08x that excludes it\n",
            S.cur->lane, line, mask);
    abort();
  }
  for (int l = 0; l < 32; ++l) {
    const int t = wi * 32 + l;
    if (t >= (int)S.fibers.size() || S.fibers[t].done) continue;
    if (!((mask >> l) & 1u)) {
      fprintf(stderr, "cuda_emu: *_sync primitive at line %d: live lane %d is not in mask %08x (partial masks are "
                      "not emulated)\n", line, l, mask);
      abort();
    }
    if (w.site[l] != line) {
      fprintf(stderr, "cuda_emu: divergent warp primitive: lane %d is at line %d, lane %d at line %d\n",
              S.cur->lane, line, l, w.site[l]);
      abort();
    }
  }
}
template <class T>
inline T warp_exchange(T v, int src_lane, unsigned mask, int line) {
  static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
  Warp& w = S.warps[S.cur->warp];
  unsigned long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.slot[S.cur->lane] = raw;
  w.site[S.cur->lane] = line;
  warp_barrier();
  check_site(mask, line);
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

inline void __syncthreads(int line = __builtin_LINE()) {
  using namespace emu;
  const unsigned my = S.bar_gen;
  // every thread of the block must reach the SAME barrier (a barrier in divergent code is undefined on the device)
  if (S.bar_count == 0) S.bar_site = line;
  else if (S.bar_site != line) {
    fprintf(stderr, "cuda_emu: block (%u,%u): threads wait at different __syncthreads (lines %d and %d)\n",
            S.block_idx.x, S.block_idx.y, S.bar_site, line);
    abort();
  }
  if (++S.bar_count >= S.live) {
    S.bar_count = 0;
    ++S.bar_gen;
  } else {
    while (S.bar_gen == my) yield();
  }
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline void __threadfence() {}
inline void __threadfence_system() {}
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int line = __builtin_LINE()) {
  return emu::warp_exchange(v, src, mask, line);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int line = __builtin_LINE()) {
  return emu::warp_exchange(v, emu::S.cur->lane ^ lane_mask, mask, line);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, int delta, int line = __builtin_LINE()) {
  const int src = emu::S.cur->lane + delta;
  return emu::warp_exchange(v, src < 32 ? src : emu::S.cur->lane, mask, line);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, int delta, int line = __builtin_LINE()) {
  const int src = emu::S.cur->lane - delta;
  return emu::warp_exchange(v, src >= 0 ? src : emu::S.cur->lane, mask, line);
}
inline unsigned __ballot_sync(unsigned mask, int pred, int line = __builtin_LINE()) {
  using namespace emu;
  Warp& w = S.warps[S.cur->warp];
  w.slot[S.cur->lane] = pred ? 1ull : 0ull;
  w.site[S.cur->lane] = line;
  warp_barrier();
  check_site(mask, line);
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
