// Issue-rate probe of tcgen05.mma kind::i8 (no data movement, no epilogue): one CTA per SM, operands
// = whatever shared memory holds, `iters` x 8 MMAs of shape M=128, N in {64,128,256}, K=32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _bin/utcimma_probe utcimma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a, uint32_t sbo, uint32_t lbo = 128) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
template <int N, int MODE>      // MODE 0: back-to-back; 1: tcgen05.commit after every 8 MMAs; 2: + rotating over 4 A blocks / 2 B tiles, SBO 2304
__global__ void __launch_bounds__(128, 1) probe(int iters, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (MODE == 2 ? 6 * 128 * 288 : (128 + N) * 256) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 0 && lane == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + (MODE == 2 ? 4 * 128 * 288 : 128 * 256);
    const uint32_t z = 0;
    constexpr uint32_t sbo = MODE == 2 ? 2304 : 2048;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = MODE == 2 ? a0 + (it & 3) * 128 * 288 : a0, b = MODE == 2 ? b0 + ((it >> 2) & 1) * 128 * 288 : b0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t acc = k > 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tm + (uint32_t)((it & 3) * (N > 128 ? 0 : N))),
                     "l"(MODE == 3 ? desc(a + k * 4096, 128, 2048) : desc(a + k * 256, sbo)), "l"(MODE == 3 ? desc(b + k * 4096, 128, 2048) : desc(b + k * 256, sbo)), "r"(idesc), "r"(acc), "r"(z), "r"(z), "r"(z), "r"(z)
                     : "memory");
      }
      if (MODE >= 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
  if (threadIdx.x == 0) out[blockIdx.x] = tm;
}

template <int N, int MODE = 0>
void run(int sms, uint32_t* out) {
  const int smem = MODE == 2 ? 6 * 128 * 288 : (128 + N) * 256, iters = 4000;
  cudaFuncSetAttribute(probe<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<N, MODE><<<sms, 128, smem>>>(10, out);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<N, MODE><<<sms, 128, smem>>>(iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double macs = (double)iters * 8 * 128.0 * N * 32 * sms;
  printf("mode %d M=128 N=%3d K=32: %8.3f ms  %7.1f TOP/s  %7.0f MAC/clk/SM at 1.965 GHz  (%s)\n", MODE, N, best, 2 * macs / best / 1e9,
         macs / (best * 1e-3) / sms / 1.965e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; cudaMalloc(&out, sms * 4);
  run<64>(sms, out); run<128>(sms, out); run<256>(sms, out);
  run<128, 1>(sms, out); run<128, 2>(sms, out); run<128, 3>(sms, out);
  cudaDeviceSynchronize();
  return 0;
}
