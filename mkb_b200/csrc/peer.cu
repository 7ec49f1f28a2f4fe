// NVLink peer-memory handshakes of the multi-GPU step: push a buffer into every peer, raise a flag in every
// peer, wait for every peer's flag.  They replace the three latency-bound NCCL collectives the round-1
// column-parallel step issued per ~0.8 ms step (an all-reduce of 3 floats, an all-gather of the ~3 MB step
// records and a 4-byte all-reduce used as a barrier): all buffers live in peer-mapped (symmetric) memory, a
// handshake is one tiny kernel on the step's stream, nothing leaves the GPU.
//
// Memory ordering.  A flag write is  __threadfence_system(); volatile store  — a release at system scope:
// every write that happens-before it (this kernel's own, and by stream order those of the kernels launched
// before it on the same stream, e.g. the peer stores of kge_adam_slice_bcast / kge_peer_copy) is visible to
// whoever observes the flag.  A flag read is  volatile load; __threadfence_system()  — the matching acquire.
// Flags are monotonically increasing step numbers (never reset), so a waiter that is several steps late
// still sees ">= value".
#include "kge_common.cuh"

namespace kge {

struct PeerPtrs {
  void* p[16];
};

// src -> dst[r] + off for every r != self: one 16-byte load, n-1 16-byte fire-and-forget NVLink stores.
__global__ void __launch_bounds__(kThreads) peer_copy_kernel(const uint4* __restrict__ src, PeerPtrs dst, int n,
                                                             int self, int64_t off16, int64_t n16) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n16; k += step) {
    const uint4 v = src[k];
    for (int r = 0; r < n; ++r)
      if (r != self) reinterpret_cast<uint4*>(dst.p[r])[off16 + k] = v;
  }
}

// flags[r][slot] = value on every peer r (self included: the local waiter polls its own array only).
__global__ void peer_signal_kernel(PeerPtrs flags, int n, int slot, unsigned value) {
  const int r = threadIdx.x;
  if (r < n) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(flags.p[r]) + slot) = value;
  }
}

// Thread r spins until flags[r] >= value.  A peer that never arrives must not hang the GPU (and the box):
// after timeout_ns the kernel gives up and raises *status, which the host checks at its next sync point.
__global__ void peer_wait_kernel(const unsigned* flags, int n, unsigned value, long long timeout_ns, int* status) {
  const int r = threadIdx.x;
  if (r < n) {
    const volatile unsigned* f = flags + r;
    long long t0 = 0;
    unsigned spins = 0;
    while ((int)(*f - value) < 0) {  // wrap-safe ">="
      if ((++spins & 1023u) == 0) {
        long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) {
          if (status) atomicOr(status, 1 << (r & 15));
          break;
        }
      }
    }
    __threadfence_system();
  }
}

}  // namespace kge

using namespace kge;

static int fill_ptrs(void* const* host, int n, PeerPtrs& out) {
  if (!host) return KGE_E_NULL;
  if (n < 1 || n > 16) return KGE_E_SIZE;
  for (int r = 0; r < n; ++r) {
    if (!host[r]) return KGE_E_NULL;
    out.p[r] = host[r];
  }
  return KGE_OK;
}

extern "C" int kge_peer_copy(const void* src, void* const* dst_peers, int32_t n_peers, int32_t self,
                             int64_t dst_offset_bytes, int64_t bytes, kge_stream_t stream) {
  PeerPtrs d{};
  if (!src) return KGE_E_NULL;
  if (int rc = fill_ptrs(dst_peers, n_peers, d)) return rc;
  if (self < 0 || self >= n_peers || bytes < 0 || dst_offset_bytes < 0) return KGE_E_SIZE;
  if (bytes % 16 || dst_offset_bytes % 16 || !aligned16(src)) return KGE_E_ALIGN;
  for (int r = 0; r < n_peers; ++r)
    if (!aligned16(d.p[r])) return KGE_E_ALIGN;
  if (bytes == 0 || n_peers == 1) return KGE_OK;
  const int64_t n16 = bytes / 16;
  int64_t blocks = (n16 + kThreads - 1) / kThreads;
  const int64_t cap = 2 * 148;  // a 3 MB record: NVLink, not the SMs, is the limit
  if (blocks > cap) blocks = cap;
  peer_copy_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(src), d, n_peers, self, dst_offset_bytes / 16, n16);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_peer_signal(void* const* flag_peers, int32_t n_peers, int32_t slot, uint32_t value,
                               kge_stream_t stream) {
  PeerPtrs f{};
  if (int rc = fill_ptrs(flag_peers, n_peers, f)) return rc;
  if (slot < 0) return KGE_E_SIZE;
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, n_peers, slot, value);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}

extern "C" int kge_peer_wait(const uint32_t* flags, int32_t n, uint32_t value, int64_t timeout_ns, int32_t* status,
                             kge_stream_t stream) {
  if (!flags) return KGE_E_NULL;
  if (n < 1 || n > 16 || timeout_ns <= 0) return KGE_E_SIZE;
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags, n, value, (long long)timeout_ns, status);
  KGE_LAUNCH_CHECK();
  return KGE_OK;
}
