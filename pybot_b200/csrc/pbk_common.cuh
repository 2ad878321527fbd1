// Shared helpers for the libpybot_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "../../include/pybot_b200.h"

namespace pbk {

// ----------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);   // sets message, returns PBK_ECUDA

#define PBK_REQUIRE(cond, code, ...)            \
  do {                                          \
    if (!(cond)) {                              \
      ::pbk::set_error(__VA_ARGS__);            \
      return (code);                            \
    }                                           \
  } while (0)

#define PBK_CUDA(expr)                                        \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::pbk::cuda_fail(_e, #expr); \
  } while (0)

#define PBK_CHECK_LAUNCH(name)                                      \
  do {                                                              \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return ::pbk::cuda_fail(_e, name);       \
  } while (0)

static inline bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// Cached per-device facts (SM count etc.); returns 0 or a pbk_status.
struct DeviceInfo {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  int max_smem_optin;
  int l2_bytes;
};
int current_device_info(const DeviceInfo** out);

static inline cudaStream_t as_stream(pbk_stream_t s) {
  return reinterpret_cast<cudaStream_t>(s);
}

// ------------------------------------------------ streaming loads / stores
// Inputs on this path are read exactly once: bypass L1 allocation.  Outputs
// are written exactly once: streaming (evict-first) stores.
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u1(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_f2(void* p, float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};"
               :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_d1(void* p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// ------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completing `bytes` on `bar`.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// L2 eviction policies for streams that are touched exactly once (evict_first) or that
// should stay resident while such a stream passes through (evict_last).  Measured on the
// C2/C4/depth kernels: evict_first hints on the TMA bulk loads / stores change nothing
// (scripts/ab_variants.sh); evict_last on the BA point gathers is worth 2 us.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_store_1d_hint(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
// ---- warp-private TMA ring ----------------------------------------------------
// Every warp owns NST shared-memory slots of one tile each and NST
// mbarriers.  Lane 0 issues a 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) for
// tile k+NST-1 while the warp computes tile k, so each warp keeps NST-1 tiles
// of HBM reads in flight without holding them in registers.  Only the single
// ragged tile at the end of the array (whose byte count is not a multiple of
// 16) is loaded with plain loads.
template <int TILE_BYTES, int NST>
struct WarpRing {
  char* slots;
  uint64_t* bars;
  __device__ __forceinline__ void init(char* warp_smem, uint64_t* warp_bars, int lane) {
    slots = warp_smem;
    bars = warp_bars;
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  __device__ __forceinline__ char* slot(int k) const { return slots + (k % NST) * TILE_BYTES; }
  // lane 0 only; `src` 16 B aligned, full tile
  __device__ __forceinline__ void issue(int k, const char* src) const {
    mbar_expect_tx(&bars[k % NST], TILE_BYTES);
    tma_load_1d(slot(k), src, TILE_BYTES, &bars[k % NST]);
  }
  // lane 0 only; ragged tile of `bytes` (multiple of 16, > 0)
  __device__ __forceinline__ void issue_bytes(int k, const char* src, uint32_t bytes) const {
    mbar_expect_tx(&bars[k % NST], bytes);
    tma_load_1d(slot(k), src, bytes, &bars[k % NST]);
  }
  // lane 0 only; full tile with an L2 eviction-policy hint
  __device__ __forceinline__ void issue_hint(int k, const char* src, uint64_t pol) const {
    mbar_expect_tx(&bars[k % NST], TILE_BYTES);
    tma_load_1d_hint(slot(k), src, TILE_BYTES, &bars[k % NST], pol);
  }
  // lane 0 only; slot k is filled by hand (ragged tile): completes its barrier phase so that
  // the parity bookkeeping of later uses of the slot stays in step
  __device__ __forceinline__ void skip(int k) const { mbar_arrive(&bars[k % NST]); }
  __device__ __forceinline__ void wait(int k) const { mbar_wait(&bars[k % NST], (uint32_t)((k / NST) & 1)); }
};

// 1-D TMA bulk store shared -> global (bulk-group completion).  Writers of the
// shared-memory source must execute fence_proxy_async_smem() and then
// synchronise with the issuing thread before the copy is issued.
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the N most recently committed groups have finished reading shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read_but() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// ... and have completed entirely
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }


// Launch facts of one kernel (one template instantiation), resolved ONCE per device: the
// shared-memory opt-in (cudaFuncSetAttribute) and the occupancy query are host-side driver
// calls of several microseconds each - per call they were a visible part of a 30-80 us step.
// A launch site keeps a `static KernelCache` and calls setup(); thread-safe.
constexpr int kMaxDevices = 64;
struct KernelCache {
  std::atomic<int> ready[kMaxDevices];
  int per_sm[kMaxDevices];
  std::mutex mu;
  // returns 0 or a pbk_status; *blocks_per_sm = resident CTAs per SM for (threads, dyn_smem)
  template <typename K>
  int setup(const DeviceInfo* di, K kernel, int threads, int dyn_smem, int* blocks_per_sm) {
    const int dev = di->device;
    if (!ready[dev].load(std::memory_order_acquire)) {
      std::lock_guard<std::mutex> lock(mu);
      if (!ready[dev].load(std::memory_order_relaxed)) {
        if (dyn_smem > 32 * 1024)        // static + dynamic beyond 48 KB needs the opt-in as well
          PBK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, (size_t)dyn_smem) != cudaSuccess ||
            n < 1)
          n = 1;
        per_sm[dev] = n;
        ready[dev].store(1, std::memory_order_release);
      }
    }
    if (blocks_per_sm) *blocks_per_sm = per_sm[dev];
    return PBK_OK;
  }
};

// Persistent grid = what is actually resident: SMs x occupancy of the kernel
// (a grid larger than that runs a ragged second wave), capped by the work.
static inline int resident_grid(const DeviceInfo* di, int per_sm, long long work_items) {
  long long g = (long long)di->sm_count * (per_sm < 1 ? 1 : per_sm);
  if (g > work_items) g = work_items;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------ fused multi-GPU exchanges
// Bound of the cross-GPU spin waits in SM clocks (default ~2 s; pbk_set_exchange_timeout_ms).
long long exchange_timeout_clk();
// A wait that ran into the bound sets *status (device memory or mapped pinned host memory)
// to one of these; the host wrapper raises.
constexpr uint32_t kStatusPeerTimeout = 1u;
__device__ __forceinline__ void raise_status(uint32_t* status, uint32_t code) {
  if (status != nullptr) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(status), "r"(code) : "memory");
    __threadfence_system();
  }
}

}  // namespace pbk
