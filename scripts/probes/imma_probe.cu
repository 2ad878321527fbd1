// Throughput probe: legacy mma.sync m16n8k32 s8 (IMMA.16832) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe imma_probe.cu && ./imma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) imma_kernel(int iters, int* sink) {
  uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 ^ 0x5a5a5a5au, a2 = a0 + 7, a3 = a1 + 3;
  uint32_t b0 = blockIdx.x * 0x01000193u + threadIdx.x, b1 = b0 ^ 0x33333333u;
  int c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + i), "r"(b1));
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 0x7fffffff) sink[0] = s;
}

template <int NACC>
void run(int ctas_per_sm, int sms, double mhz) {
  int* sink; cudaMalloc(&sink, 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  imma_kernel<NACC><<<sms * ctas_per_sm, 256>>>(100, sink);
  cudaEventRecord(e0);
  imma_kernel<NACC><<<sms * ctas_per_sm, 256>>>(iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)sms * ctas_per_sm * 8 * iters * NACC;
  const double macs = mmas * 16 * 8 * 32;
  printf("NACC=%d ctas/sm=%d: %.3f ms, %.1f TMAC/s = %.2f POPS, %.0f MAC/clk/SM at %.0f MHz, pairs(256b)/clk/SM = %.1f\n", NACC,
         ctas_per_sm, ms, macs / ms / 1e9, 2 * macs / ms / 1e12, macs / (ms * 1e-3) / sms / (mhz * 1e6), mhz,
         macs / 256 / (ms * 1e-3) / sms / (mhz * 1e6));
  cudaFree(sink);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, khz);
  run<1>(2, p.multiProcessorCount, khz / 1e3);
  run<2>(2, p.multiProcessorCount, khz / 1e3);
  run<4>(2, p.multiProcessorCount, khz / 1e3);
  run<8>(2, p.multiProcessorCount, khz / 1e3);
  run<4>(4, p.multiProcessorCount, khz / 1e3);
  run<8>(1, p.multiProcessorCount, khz / 1e3);
  return 0;
}
