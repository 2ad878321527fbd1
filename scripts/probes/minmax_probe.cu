// Throughput of the candidate min/max instructions for the tensor-core matcher's top-2 network
// (build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _bin/minmax_probe minmax_probe.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, uint32_t* sink) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 2654435761u + i * 40503u; b[i] = a[i] * 0x9e3779b9u + sink[i & 3]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("max.u16x2 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 1) { asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("max.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 2) { asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("max.f16x2 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 3) { asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&b[i])); asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&b[i]) : "f"(*(float*)&a[i])); }
        if (MODE == 4) { asm volatile("min.s16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("max.s16x2 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 5) { asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 6) { asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(a[i]) : "r"(b[i])); asm volatile("lop3.b32 %0, %0, %1, %1, 0xe8;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 8) { asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(*(float*)&a[i]) : "f"(*(float*)&b[i]), "f"(*(float*)&a[(i + 1) & 7])); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(*(float*)&b[i]) : "f"(*(float*)&a[i]), "f"(*(float*)&b[(i + 1) & 7])); }
        if (MODE == 9) { a[i] = (uint32_t)min(min((int)a[i], (int)b[i]), (int)a[(i + 1) & 7]); b[i] = (uint32_t)max(max((int)b[i], (int)a[i]), (int)b[(i + 1) & 7]); asm volatile("" : "+r"(a[i]), "+r"(b[i])); }
        if (MODE == 10) { asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a[i]) : "r"(b[i])); asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 11) { asm volatile("min.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("max.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i])); }
        if (MODE == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(b[i])); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[i]) : "r"(a[i])); }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i];
  sink[8 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, uint32_t* sink, int sms) {
  const int iters = 2000, ctas = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<ctas, 256>>>(10, sink);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double ops = (double)iters * 8 * 8 * 2 * 256 * ctas;
  printf("%-12s %8.3f ms  %7.2f Tops/s  = %6.1f thread-ops / clk / SM at 1.965 GHz\n", name, best, ops / best / 1e9,
         ops / (best * 1e-3) / sms / 1.965e9);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* sink; cudaMalloc(&sink, (8 + sms * 8 * 256) * 4 + 64); cudaMemset(sink, 0, 64);
  run<0>("minmax.u16x2", sink, sms);
  run<4>("minmax.s16x2", sink, sms);
  run<1>("minmax.u32", sink, sms);
  run<2>("minmax.f16x2", sink, sms);
  run<5>("minmax.bf16x2", sink, sms);
  run<3>("minmax.f32", sink, sms);
  run<6>("lop3", sink, sms);
  run<7>("imad", sink, sms);
  run<8>("minmax3.f32", sink, sms);
  run<9>("minmax3.s32", sink, sms);
  run<10>("prmt", sink, sms);
  run<11>("minmax.s32", sink, sms);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
