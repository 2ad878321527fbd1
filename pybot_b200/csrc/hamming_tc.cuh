// Shared pieces of the two tensor-core Hamming scan kernels (hamming_tc.cu: tcgen05 kind::i8,
// hamming_tc4.cu: tcgen05 kind::mxf4): tile geometry, TMEM / mbarrier helpers, launch parameters.
#pragma once
#include "pbk_common.cuh"

namespace pbk {

constexpr int kTcM = 128;                         // queries per UMMA (TMEM lanes)
constexpr int kTcQBlocks = 4;                     // query blocks resident per CTA
constexpr int kTcQGroup = kTcM * kTcQBlocks;      // 512 queries per CTA
constexpr int kTcN = 128;                         // train rows per tile = UMMA N = accumulator columns
constexpr int kTcPlaneBytes = 128 * 16;           // 2048: one 16-byte K chunk of the 128 rows of a tile
constexpr int kTcEpiWarps = 8;                    // warps 0-7: warp w may only touch TMEM lanes 32 (w % 4) .. +31
constexpr int kTcPad = 512;                       // expanded buffers are padded to whole query groups / tiles
constexpr uint32_t kTcKeyMul = (1u << 23) + 1u;   // 32-bit key = distance * kTcKeyMul + row in split (as hamming.cu)
constexpr uint32_t kTcNone = 0xFFFFFFFFu;
constexpr long long kTcMaxSplitRows = 1ll << 23;
#ifndef PBK_TC_PARK_NS
#define PBK_TC_PARK_NS 1000
#endif
#ifndef PBK_TC_PREFETCH
#define PBK_TC_PREFETCH 4
#endif
constexpr int kTcPrefetch = PBK_TC_PREFETCH;      // tiles requested into L2 ahead of the shared-memory ring

// ------------------------------------------------------------------ tcgen05 / TMEM helpers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, no swizzle, LBO = 2048 B (K chunks), SBO = 128 B (8-row groups)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kTcPlaneBytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
         (1ull << 46);                               // bits [46,48): descriptor version 1 (Blackwell)
}
// instruction descriptor, kind::i8: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kTcIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kTcIdesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z)
      : "memory");
}
// the mbarrier receives one arrival when every tcgen05 operation this thread issued so far is done
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 64 consecutive accumulator columns of this thread's TMEM lane, their low 16 bits packed two per
// register (column 2 i in the low half of r[i], column 2 i + 1 in the high half)
__device__ __forceinline__ void tc_ld64_packed(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// one lane of a converged warp (the tcgen05 issue instructions take uniform operands: keep the code
// around them warp-uniform and predicate only the instruction)
__device__ __forceinline__ bool tc_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mbarrier wait for the warps that may sleep: try_wait with a suspend-time hint, so a waiting warp
// parks inside the instruction instead of polling shared memory - the tensor pipe fetches its
// operands from the same shared memory
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITP_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONEP_%=;\n\t"
      "bra WAITP_%=;\n\t"
      "DONEP_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(PBK_TC_PARK_NS)
      : "memory");
}

__device__ __forceinline__ void tc_top2_u16x2(uint32_t& k1, uint32_t& k2, uint32_t key) {
  uint32_t mx;
  asm("max.u16x2 %0, %1, %2;" : "=r"(mx) : "r"(k1), "r"(key));
  asm("min.u16x2 %0, %1, %2;" : "=r"(k1) : "r"(k1), "r"(key));
  asm("min.u16x2 %0, %1, %2;" : "=r"(k2) : "r"(k2), "r"(mx));
}
__device__ __forceinline__ void tc_top2_s16x2(uint32_t& k1, uint32_t& k2, uint32_t key) {
  uint32_t mx;
  asm("max.s16x2 %0, %1, %2;" : "=r"(mx) : "r"(k1), "r"(key));
  asm("min.s16x2 %0, %1, %2;" : "=r"(k1) : "r"(k1), "r"(key));
  asm("min.s16x2 %0, %1, %2;" : "=r"(k2) : "r"(k2), "r"(mx));
}
__device__ __forceinline__ void tc_top2(uint32_t& k1, uint32_t& k2, uint32_t key) {
  const uint32_t mx = max(k1, key);
  k1 = min(k1, key);
  k2 = min(k2, mx);
}

struct HammingTcParams {
  const int8_t* q_signs;          // expanded queries, padded to whole groups of 512 rows
  const int8_t* t_signs;          // expanded train rows of this call, padded to whole tiles
  unsigned long long* partial;    // (2 * splits, nq, 2)
  uint32_t* gd2;                  // mxf4 kernel: per (padded) query, a 16-byte record of distance bounds shared by ALL CTAs (tc4_bound)
  long long nq, nt;
  long long rows_per_split;       // multiple of kTcN
  unsigned long long train_base;
  int splits;
  int dbg;                        // PBK_TC_DEBUG timing probes, bits (results are then wrong): 1 = no selection, 2 = no MMA, 4 = no TMEM load, 8 = no tile loads (mxf4 kernel)
};


static inline long long tc_pad(long long n) { return (n + kTcPad - 1) / kTcPad * kTcPad; }

// the two scan engines (same contract): expand q into `q_signs`, scan t_signs, partial keys
// (2 * splits, nq, 2) into `partial`
int hamming_tc8_scan(const uint8_t* q, long long nq, const int8_t* t_signs, long long nt, unsigned long long train_base,
                     int8_t* q_signs, unsigned long long* partial, int splits, long long rows_per_split,
                     const DeviceInfo* di, cudaStream_t s);
int hamming_tc4_scan(const uint8_t* q, long long nq, const int8_t* t_signs, long long nt, unsigned long long train_base,
                     int8_t* q_signs, unsigned long long* partial, int splits, long long rows_per_split,
                     const DeviceInfo* di, cudaStream_t s);
int hamming_tc4_expand(const uint8_t* d, long long n, long long first_row, long long n_padded, int8_t* signs, cudaStream_t s);
int hamming_tc4_peak(int iters, uint32_t* sink, const DeviceInfo* di, cudaStream_t s);
constexpr int kTc8RowBytes = 256;                 // one expanded descriptor, kind::i8: 256 sign bytes
constexpr int kTc4RowBytes = 128;                 // kind::mxf4: 256 e2m1 nibbles

}  // namespace pbk
