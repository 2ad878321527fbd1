// Brute-force Hamming kNN (k = 2) on the 5th-generation tensor cores: tcgen05.mma kind::i8 with
// the accumulators in tensor memory (VERDICT r1 "next round" item 9).
//
// Same contract as hamming_knn2_kernel (hamming.cu) - per query the two smallest
// (distance, global train row) keys, bit-identical to cv2.BFMatcher - by a different route.
// A 256-bit descriptor is expanded ONCE into signed bytes: a query into a_k = bit_k ? +1 : -1,
// a train row into b_k = bit_k ? -64 : +64, so that the exact int32 dot product is
//     sum_k a_k b_k = -64 (256 - 2 d) = 128 d - 16384 .
// The accumulator of (query, train row) is a multiple of 128 that orders like the distance and
// fits 16 bits, and tcgen05.ld.pack::16b delivers two adjacent columns per register.  The epilogue
// first takes the SMALLEST value of an item (64 columns of one query row) with a packed signed
// min tree - 0.5 ALU instructions per pair; the POPC / LOP3 route issues 23.6 per pair and is
// bound by the 16-lane XU pipe at 4 pairs per clock per SM (profiles/r02_hamming_inner_loop.md) -
// and only when that value could still enter some query's top-2 (a few per cent of the items of
// a long scan) builds the full keys
//     key = acc + c = 128 d - 16384 + c     (c = column of the tile, 0 .. 127; signed 16 bit)
// and runs the packed signed top-2 network over them (1.5 instructions per pair).
// (Until round 2's last step a 33rd 32-byte K step of constants made the accumulator the finished
// key 128 d + c: 12.5 % more tensor work and HBM bytes to save ALU work that the pre-filter had
// already removed.)  Rows past the end of the database are all-zero (distance code 128) and are
// dropped by row number when keys are built.
//
// Data layout.  Expanded rows (256 B) live in HBM in exactly the shared-memory image the UMMA
// descriptors describe (K-major, no swizzle), tile by tile: a tile is 128 rows = 16 planes (one
// per 16-byte K chunk) of 128 x 16 bytes, i.e. of sixteen 8-row "core matrices" (128 B each):
//     byte address of element k of row r  =  (r / 128) * 32768 + (k / 16) * 2048 + (r % 128) * 16 + k % 16
// => stride (8-row group) byte offset 128, leading (K-chunk) byte offset 2048; a tile is one
// contiguous 32 KB span and arrives by ONE plain 1-D bulk copy (cp.async.bulk), no tensor map.
// (With the chunks of a row group adjacent instead - LBO 128, SBO 2304 - the same MMAs ran 6 %
// slower in isolation: scripts/probes/utcimma_probe.cu.)  A database keeps its expanded copy resident next to the
// compact one (8 x the bytes: 5.1 GB for the 20 M-row C3 database; the compact rows are still
// what the mutual check gathers).
//
// Kernel.  One CTA per SM (224 KB of shared memory, all 512 TMEM columns), grid = query groups
// of 512 x database splits.  A CTA keeps its 512 expanded queries (4 blocks of M = 128) in
// shared memory for its whole life and streams its split in tiles of N = 128 rows (32 KB,
// 3-stage ring).  Work item = (tile, query block): 8 UTCIMMA (K = 32 each) into one of FOUR
// 128-column accumulator slots, handed to the epilogue through mbarriers
// (tcgen05.commit -> acc_full, epilogue arrive -> acc_empty), so the tensor pipe runs up to
// three items ahead of the selection.
//   warp 9      producer: bulk copies of the query blocks and of the train tiles
//   warp 8      TMEM allocation; one lane issues every tcgen05.mma / tcgen05.commit
//   warps 0-7   epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 (its hardware lane quarter)
//               and columns 64 h .. 64 h + 63 of a slot, h = w / 4: one packed tcgen05.ld
//               (32 registers = 64 keys), requested one item ahead; the slot is released as soon
//               as the values are in registers; top-2 network; the two survivors of the item are
//               folded into 32-bit keys.  A lane is one query row of each of the four blocks.
// Every (split, column half) emits (distance << 32 | global row) keys; the existing merge
// kernels reduce them - bit-identical to a single ascending scan (unsigned min).
#include <atomic>
#include <cstdlib>
#include "hamming_tc.cuh"

namespace pbk {

constexpr int kTcStages = 3;
constexpr int kTcRowBytes = 256;                  // one expanded descriptor: 256 sign bytes
constexpr int kTcChunks = kTcRowBytes / 16;       // 16 K chunks of 16 bytes
constexpr int kTcKSteps = kTcRowBytes / 32;       // 8 UTCIMMA per item
constexpr int kTcTileBytes = kTcN * kTcRowBytes;  // 32 KB
constexpr int kTcBlockBytes = kTcM * kTcRowBytes; // 32 KB
constexpr int kTcSlots = 4;                       // accumulator slots of 128 columns
constexpr int kTcThreads = 352;                   // 11 warps
// The scheduler of an SM sub-partition favours its HIGHEST warp id.  The one lane that issues
// every tcgen05.mma sits in warp 8 (sub-partition 0, above selecting warps 0 and 4) and the
// producer in warp 9: as warps 1 / 0 under the selecting warps the issuer was starved for issue
// slots and the tensor pipe ran at 62 % (950 instead of 640 clocks per item).
// TWO issuing warps (8 and 9: query blocks 0, 2 and 1, 3 of every tile): the tensor pipe accepts
// little work ahead of what it executes, so the ~400 clocks of serial bookkeeping one thread spends
// per item (barrier round trips, descriptor arithmetic, commits) were not hidden behind the 640
// clocks of its nine MMAs - the pipe measured 55 % busy with one issuer.
constexpr int kTcMmaWarp = 8, kTcMmaWarps = 2, kTcProducerWarp = 10;

constexpr int kTcSmemA = kTcQGroup * kTcRowBytes;                 // 131072
constexpr int kTcSmemB = kTcStages * kTcTileBytes;                // 98304
constexpr int kTcSmemBars = (2 * kTcStages + 2 * kTcSlots + 1) * 8;
constexpr int kTcSmemBytes = kTcSmemA + kTcSmemB + kTcSmemBars + 16;

// ------------------------------------------------------------ expansion: bits -> signed bytes
// thread = (row, 16-byte K chunk): 2 descriptor bytes in, one 16-byte core-matrix row out.
// QUERY: +-1; train: -+64; rows past the end: zeros.
template <bool QUERY>
__global__ void __launch_bounds__(256)
hamming_tc_expand_kernel(const uint8_t* __restrict__ d, long long n, long long first_row, long long n_padded,
                         uint4* __restrict__ signs) {
  // 16-byte piece index; rows [first_row, n_padded) are (re)written (first_row: a multiple of 512)
  const long long i = first_row * kTcChunks + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded * kTcChunks) return;
  const long long tile = i / (kTcN * kTcChunks);      // 2048 pieces per tile: [K chunk][row of the tile]
  const int in_t = (int)(i - tile * (kTcN * kTcChunks));
  const int kc = in_t >> 7, rt = in_t & 127;
  const long long r = tile * kTcN + rt;
  const bool real = r < n;
  uint4 o = make_uint4(0, 0, 0, 0);
  if (real) {
    const unsigned bits = (unsigned)d[r * 32 + kc * 2] | ((unsigned)d[r * 32 + kc * 2 + 1] << 8);
    // nibble -> four bytes of 0 / 1 (bit i -> byte i); query: 1 -> 0x01, 0 -> 0xFF (+-1);
    // train: 1 -> 0xC0 (-64), 0 -> 0x40 (+64)
    auto four = [](unsigned nib) -> unsigned {
      const unsigned t = (nib * 0x00204081u) & 0x01010101u;
      return QUERY ? (0xFFFFFFFFu ^ (t * 0xFEu)) : (0x40404040u + t * 0x80u);
    };
    o = make_uint4(four(bits & 15u), four((bits >> 4) & 15u), four((bits >> 8) & 15u), four((bits >> 12) & 15u));
  }
  signs[i] = o;
}

// top-2 of one accumulator half-slot (64 columns = 32 packed registers, acc = 128 d - 16384 in
// each signed 16-bit half) of this thread's query row, folded into (k1, k2).  ONE copy of this
// code serves every query block and both column halves (the kernel is instruction-cache
// sensitive: an earlier build with eight unrolled copies spent 90 % of its stall samples on
// instruction fetch).  tile_base = row in the split of the tile's column 0; rows = rows in the
// split; h = column half of the slot this warp reads (columns 64 h ..).
__device__ __forceinline__ void tc_select(const uint32_t (&v)[32], uint32_t tile_base, uint32_t rows, uint32_t h,
                                          uint32_t& k1, uint32_t& k2) {
  // Most items cannot change a query's top-2 once a few thousand rows have been seen, so the
  // item is first reduced to its SMALLEST accumulator only (a signed min tree: 31 packed
  // operations) and compared with the current second-best distance k2 >> 23 (k2 = d (2^23 + 1) +
  // row): an element can only enter with a distance <= that.  The exact selection below runs
  // only when some lane of the warp needs it; the result is the same either way.
  {
    uint32_t m[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) asm("min.s16x2 %0, %1, %2;" : "=r"(m[j]) : "r"(v[2 * j]), "r"(v[2 * j + 1]));
#pragma unroll
    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
      for (int j = 0; j < w; ++j) asm("min.s16x2 %0, %1, %2;" : "=r"(m[j]) : "r"(m[j]), "r"(m[j + w]));
    const int best = min((int)(short)(m[0] & 0xFFFFu), (int)m[0] >> 16);
    const int bound = (int)((k2 >> 23) << 7) - 16384;
    if (!__any_sync(0xffffffffu, best <= bound)) return;
  }
  // keys: acc + column of the tile (acc is a multiple of 128: the packed add never carries from
  // one half into the other), ordered as SIGNED 16-bit values; two independent trackers (even /
  // odd registers) halve the dependent min / max chain, then merge
  uint32_t p[2][2] = {{0x7FFF7FFFu, 0x7FFF7FFFu}, {0x7FFF7FFFu, 0x7FFF7FFFu}};
  const uint32_t col0 = (h * 64u) * 0x00010001u + 0x00010000u;       // columns (64 h, 64 h + 1) of register 0
#pragma unroll
  for (int j = 0; j < 32; ++j) tc_top2_s16x2(p[j & 1][0], p[j & 1][1], v[j] + col0 + (uint32_t)(2 * j) * 0x00010001u);
  uint32_t lo, hi, t;
  asm("min.s16x2 %0, %1, %2;" : "=r"(lo) : "r"(p[0][0]), "r"(p[1][0]));
  asm("max.s16x2 %0, %1, %2;" : "=r"(t) : "r"(p[0][0]), "r"(p[1][0]));
  asm("min.s16x2 %0, %1, %2;" : "=r"(hi) : "r"(p[0][1]), "r"(p[1][1]));
  asm("min.s16x2 %0, %1, %2;" : "=r"(hi) : "r"(hi), "r"(t));
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t w = (c >> 1) ? hi : lo;
    const int key16 = (c & 1) ? ((int)w >> 16) : (int)(short)(w & 0xFFFFu);
    // distance = (key + 16384) >> 7, column of the tile = key & 127 -> 32-bit key = distance *
    // kTcKeyMul + row in the split; rows past the end of the database (zero rows: distance code
    // 128) and unused tracker slots (0x7FFF) are dropped
    const uint32_t row = tile_base + ((uint32_t)key16 & 127u);
    const uint32_t key = (row >= rows || key16 == 0x7FFF) ? kTcNone
                                                          : ((uint32_t)(key16 + 16384) >> 7) * kTcKeyMul + row;
    tc_top2(k1, k2, key);
  }
}

__global__ void __launch_bounds__(kTcThreads, 1)
hamming_tc_kernel(const __grid_constant__ HammingTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTcSmemA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcSmemA + kTcSmemB);
  uint64_t* b_full = bars;                           // [kTcStages]  train tile landed
  uint64_t* b_empty = bars + kTcStages;              // [kTcStages]  every MMA that read the tile is done
  uint64_t* acc_full = bars + 2 * kTcStages;         // [kTcSlots]   accumulator slot written
  uint64_t* acc_empty = acc_full + kTcSlots;         // [kTcSlots]   all 8 epilogue warps have read it
  uint64_t* a_full = acc_empty + kTcSlots;           // query blocks landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_full + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qg = blockIdx.x, split = blockIdx.y;
  const long long row0 = (long long)split * P.rows_per_split;
  long long rows = P.nt - row0;
  if (rows > P.rows_per_split) rows = P.rows_per_split;
  if (rows < 0) rows = 0;
  const int ntiles = (int)((rows + kTcN - 1) / kTcN);

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], kTcMmaWarps); }
    for (int s = 0; s < kTcSlots; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kTcEpiWarps); }
    mbar_init(a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kTcMmaWarp) {                          // all 512 TMEM columns: one CTA per SM (shared memory bound)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kTcProducerWarp) {
    // ================================================================= producer
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(a_full, kTcSmemA);
      for (int mb = 0; mb < kTcQBlocks; ++mb)
        tma_load_1d(sA + mb * kTcBlockBytes, P.q_signs + ((size_t)qg * kTcQGroup + (size_t)mb * kTcM) * kTcRowBytes,
                    kTcBlockBytes, a_full);
      const int8_t* src = P.t_signs + (size_t)row0 * kTcRowBytes;
      // the two-stage ring leaves one tile time (~1.3 us) for a 36 KB load: tiles are pulled into
      // L2 kTcPrefetch tiles ahead so that the shared-memory copy is an L2 hit
      for (int t = 0; t < kTcPrefetch && t < ntiles; ++t)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)t * kTcTileBytes), "r"(kTcTileBytes) : "memory");
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % kTcStages;
        if (t + kTcPrefetch < ntiles)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)(t + kTcPrefetch) * kTcTileBytes), "r"(kTcTileBytes) : "memory");
        if (t >= kTcStages) mbar_wait_parked(&b_empty[s], (uint32_t)(((t / kTcStages) - 1) & 1));
        mbar_expect_tx(&b_full[s], kTcTileBytes);
        tma_load_1d(sB + s * kTcTileBytes, src + (size_t)t * kTcTileBytes, kTcTileBytes, &b_full[s]);
      }
    }
  } else if (warp >= kTcMmaWarp && warp < kTcMmaWarp + kTcMmaWarps) {
    // =============================================================== MMA issuers
    // the whole warp runs the loop with warp-uniform values, one elected lane issues (see
    // hamming_tc4.cu: under `if (lane == 0)` every UTCIMMA sits in an ELECT / R2UR waterfall loop)
    if (ntiles > 0) {
      mbar_wait(a_full, 0);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      const int mine = __shfl_sync(0xffffffffu, warp - kTcMmaWarp, 0);   // this issuer's query blocks: mine, mine + kTcMmaWarps
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % kTcStages;
        mbar_wait(&b_full[s], (uint32_t)((t / kTcStages) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int mb = mine; mb < kTcQBlocks; mb += kTcMmaWarps) {
          const int item = t * kTcQBlocks + mb;
          const int slot = item & (kTcSlots - 1), use = item / kTcSlots;
          if (use > 0) {
            mbar_wait(&acc_empty[slot], (uint32_t)((use - 1) & 1));
            tc_fence_after();
          }
          const uint32_t d = tmem_base + (uint32_t)(slot * kTcN);
          const uint64_t ad = tc_smem_desc(a_addr + mb * kTcBlockBytes), bd = tc_smem_desc(b_addr + s * kTcTileBytes);
          if (tc_elect_one()) {
            if (!(P.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < kTcKSteps; ++k)               // K = 32 bytes per instruction = two planes: +4096 B = +256 in the address field
                tc_mma_i8(d, ad + (uint64_t)(k * 256), bd + (uint64_t)(k * 256), k > 0 ? 1u : 0u);
            }
            tc_commit(&acc_full[slot]);
          }
          __syncwarp();
        }
        if (tc_elect_one()) tc_commit(&b_empty[s]);   // the tile's shared-memory stage may be refilled
        __syncwarp();
      }
    }
  } else {
    // ================================================================= epilogue
    const int lg = warp & 3, h = warp >> 2;          // TMEM lane quarter of this warp; column half of a slot
    // k1 / k2 of the thread's four queries (one per query block), rotated once per item so that
    // the current block's pair is always element 0 (register-resident without unrolling the loop)
    uint32_t k1[kTcQBlocks], k2[kTcQBlocks];
#pragma unroll
    for (int mb = 0; mb < kTcQBlocks; ++mb) k1[mb] = k2[mb] = kTcNone;
    const int nitems = ntiles * kTcQBlocks;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(h * 64);
    // Software pipeline: the accumulators of item i + 1 are requested (tcgen05.ld, asynchronous)
    // before item i is selected, so the TMEM read latency and the wait for the tensor pipe hide
    // behind ~250 instructions of selection; the slot goes back to the MMA issuer as soon as its
    // values are in registers.  Two register sets, the loop unrolled by two (no copies).
    auto fetch = [&](int item, uint32_t (&v)[32]) {
      const int slot = item & (kTcSlots - 1);
      mbar_wait_parked(&acc_full[slot], (uint32_t)((item / kTcSlots) & 1));
      tc_fence_after();
      tc_ld64_packed(lane_addr + (uint32_t)(slot * kTcN), v);
    };
    auto landed = [&](int item) {            // the values of `item` are in registers: release its slot
      tc_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[item & (kTcSlots - 1)]);
    };
    auto select = [&](int item, const uint32_t (&v)[32]) {
      if (!(P.dbg & 1)) tc_select(v, (uint32_t)(item >> 2) * kTcN, (uint32_t)rows, (uint32_t)h, k1[0], k2[0]);
      else k1[0] ^= v[0] ^ v[31] ^ v[17];
      // rotate: the next item belongs to the next query block
      const uint32_t a1 = k1[0], a2 = k2[0];
#pragma unroll
      for (int m = 0; m < kTcQBlocks - 1; ++m) { k1[m] = k1[m + 1]; k2[m] = k2[m + 1]; }
      k1[kTcQBlocks - 1] = a1;
      k2[kTcQBlocks - 1] = a2;
    };
    uint32_t v0[32], v1[32];
    if (nitems > 0) {
      fetch(0, v0);
      landed(0);
    }
#pragma unroll 1
    for (int item = 0; item < nitems; item += 2) {          // nitems is a multiple of 4
      fetch(item + 1, v1);
      select(item, v0);
      landed(item + 1);
      if (item + 2 < nitems) fetch(item + 2, v0);
      select(item + 1, v1);
      if (item + 2 < nitems) landed(item + 2);
    }
    // nitems is a multiple of 4: after the loop element m is query block m again
    const unsigned long long gbase = P.train_base + (unsigned long long)row0;
#pragma unroll
    for (int mb = 0; mb < kTcQBlocks; ++mb) {
      const long long qi = (long long)qg * kTcQGroup + mb * kTcM + lg * 32 + lane;
      if (qi < P.nq) {
        unsigned long long o[2];
        const uint32_t kk[2] = {k1[mb], k2[mb]};
#pragma unroll
        for (int c = 0; c < 2; ++c)
          o[c] = (kk[c] == kTcNone) ? PBK_KEY_NONE
                                    : (((unsigned long long)(kk[c] / kTcKeyMul) << 32) | (gbase + (kk[c] % kTcKeyMul)));
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.partial + ((long long)(split * 2 + h) * P.nq + qi) * 2);
        *dst = make_ulonglong2(o[0], o[1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTcMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ int8 tensor peak probe
// Measurement only: one CTA per SM issues `iters` x 8 UTCIMMA (M 128, N 128, K 32) on whatever its
// shared memory holds - no data movement, no epilogue: the rate the tensor pipe sustains for the
// matcher's instruction shape.  bench.py measures the roofline denominator with it in the same run.
__global__ void __launch_bounds__(128, 1) hamming_tc_peak_kernel(int iters, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * kTcM * 256 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 0 && lane == 0) {
    const uint32_t a = smem_u32(smem), b = a + kTcM * 256;
    // 256-byte rows here (SBO 2048): the plain K-major layout
    auto desc = [](uint32_t addr) -> uint64_t {
      return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46);
    };
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) tc_mma_i8(tm + (uint32_t)((it & 3) * kTcN), desc(a + k * 256), desc(b + k * 256), k > 0 ? 1u : 0u);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
  if (threadIdx.x == 0 && sink != nullptr) sink[blockIdx.x] = tm;
}


// ------------------------------------------------------------------ format selection
// Two encodings of the expanded descriptors, same keys bit for bit:
//   4  (default) e2m1 nibbles, tcgen05.mma kind::mxf4 (hamming_tc4.cu): 128 B per row, twice the MMA rate
//   8  signed bytes, tcgen05.mma kind::i8 (this file): 256 B per row
// The format is a property of the expanded buffers: choose it (PBK_TC_FORMAT in the environment, or
// pbk_hamming_tc_set_format) before expanding a database and keep it for the scans of that copy.
static std::atomic<int> g_tc_format{0};
static int tc_format() {
  int f = g_tc_format.load(std::memory_order_acquire);
  if (f == 0) {
    f = 4;
    if (const char* e = getenv("PBK_TC_FORMAT")) f = (atoi(e) == 8) ? 8 : 4;
    g_tc_format.store(f, std::memory_order_release);
  }
  return f;
}
static int tc_row_bytes() { return tc_format() == 8 ? kTc8RowBytes : kTc4RowBytes; }

// grid: query groups x splits ~ one CTA per SM; every split a whole number of tiles
static int tc_plan(const DeviceInfo* di, long long nq, long long nt, int* splits, long long* rows_per_split) {
  const long long qgroups = (nq + kTcQGroup - 1) / kTcQGroup;
  long long s = qgroups > 0 ? di->sm_count / qgroups : 1;
  const long long tiles = (nt + kTcN - 1) / kTcN;
  if (s > tiles) s = tiles;
  if (s < 1) s = 1;
  long long rps = (nt + s - 1) / s;
  rps = (rps + kTcN - 1) / kTcN * kTcN;
  if (rps > kTcMaxSplitRows) rps = kTcMaxSplitRows;
  if (rps < kTcN) rps = kTcN;
  s = (nt + rps - 1) / rps;
  if (s < 1) s = 1;
  PBK_REQUIRE(s <= 32767, PBK_EINVAL, "train set too large for one call (%lld splits)", s);
  *splits = (int)s;
  *rows_per_split = rps;
  return PBK_OK;
}

int hamming_tc8_scan(const uint8_t* q, long long nq, const int8_t* t_signs, long long nt, unsigned long long train_base,
                     int8_t* q_signs, unsigned long long* partial, int splits, long long rps, const DeviceInfo* di,
                     cudaStream_t s) {
  const long long nqp = tc_pad(nq);
  hamming_tc_expand_kernel<true><<<(unsigned)((nqp * kTcChunks + 255) / 256), 256, 0, s>>>(q, nq, 0, nqp, reinterpret_cast<uint4*>(q_signs));
  PBK_CHECK_LAUNCH("hamming_tc_expand_kernel");
  static KernelCache kc;
  int rc = kc.setup(di, hamming_tc_kernel, kTcThreads, kTcSmemBytes, nullptr);
  if (rc) return rc;
  HammingTcParams P;
  P.q_signs = q_signs; P.t_signs = t_signs; P.partial = partial; P.gd2 = nullptr;
  P.nq = nq; P.nt = nt; P.rows_per_split = rps; P.train_base = train_base; P.splits = splits;
  P.dbg = 0;
  if (const char* e = getenv("PBK_TC_DEBUG")) P.dbg = atoi(e);
  const dim3 grid((unsigned)((nq + kTcQGroup - 1) / kTcQGroup), (unsigned)splits);
  hamming_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, s>>>(P);
  PBK_CHECK_LAUNCH("hamming_tc_kernel");
  return PBK_OK;
}

// expands q into `q_signs` and scans t_signs: partial keys (2 * splits, nq, 2) into `partial`
int hamming_tc_scan(const uint8_t* q, long long nq, const int8_t* t_signs, long long nt, unsigned long long train_base,
                    int8_t* q_signs, unsigned long long* partial, int* parts, const DeviceInfo* di, cudaStream_t s) {
  int splits;
  long long rps;
  int rc = tc_plan(di, nq, nt, &splits, &rps);
  if (rc) return rc;
  rc = tc_format() == 8 ? hamming_tc8_scan(q, nq, t_signs, nt, train_base, q_signs, partial, splits, rps, di, s)
                        : hamming_tc4_scan(q, nq, t_signs, nt, train_base, q_signs, partial, splits, rps, di, s);
  if (rc) return rc;
  *parts = 2 * splits;
  return PBK_OK;
}

}  // namespace pbk

extern "C" {

int pbk_hamming_tc_format(void) { return pbk::tc_format(); }

int pbk_hamming_tc_set_format(int format) {
  using namespace pbk;
  PBK_REQUIRE(format == 4 || format == 8, PBK_EINVAL, "format must be 4 (e2m1, kind::mxf4) or 8 (s8, kind::i8)");
  g_tc_format.store(format, std::memory_order_release);
  return PBK_OK;
}

size_t pbk_hamming_tc_signs_bytes(int64_t n) {
  if (n < 0) n = 0;
  return (size_t)pbk::tc_pad(n) * pbk::tc_row_bytes();
}

int pbk_hamming_tc_expand(const uint8_t* d, int64_t n, int64_t first_row, int8_t* signs, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && first_row >= 0 && first_row % kTcPad == 0, PBK_EINVAL,
              "first_row must be a non-negative multiple of %d", kTcPad);
  if (n == 0 || first_row >= tc_pad(n)) return PBK_OK;
  PBK_REQUIRE(d != nullptr && signs != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(signs), PBK_EALIGN, "device pointers must be 16-byte aligned");
  const long long np = tc_pad(n);
  if (tc_format() == 4) return hamming_tc4_expand(d, n, first_row, np, signs, as_stream(stream));
  const long long blocks = ((np - first_row) * kTcChunks + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "too many rows for one call");
  hamming_tc_expand_kernel<false><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d, n, first_row, np,
                                                                                  reinterpret_cast<uint4*>(signs));
  PBK_CHECK_LAUNCH("hamming_tc_expand_kernel");
  return PBK_OK;
}

int pbk_hamming_tc_peak_probe(int iters, uint32_t* sink, double* macs_per_launch, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(iters > 0, PBK_EINVAL, "bad arguments");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  // one iteration = all 256 bits of 128 x 128 descriptor pairs in either format
  if (macs_per_launch) *macs_per_launch = (double)iters * 8.0 * kTcM * kTcN * 32.0 * di->sm_count;
  if (tc_format() == 4) return hamming_tc4_peak(iters, sink, di, as_stream(stream));
  constexpr int smem = 2 * kTcM * 256;
  static KernelCache kc;
  rc = kc.setup(di, hamming_tc_peak_kernel, 128, smem, nullptr);
  if (rc) return rc;
  hamming_tc_peak_kernel<<<di->sm_count, 128, smem, as_stream(stream)>>>(iters, sink);
  PBK_CHECK_LAUNCH("hamming_tc_peak_kernel");
  return PBK_OK;
}

int pbk_hamming_tc_plan(int64_t nq, int64_t nt, int* parts, size_t* scratch_bytes) {
  using namespace pbk;
  PBK_REQUIRE(nq >= 0 && nt >= 0 && parts && scratch_bytes, PBK_EINVAL, "bad arguments");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  int splits;
  long long rps;
  rc = tc_plan(di, nq, nt, &splits, &rps);
  if (rc) return rc;
  *parts = 2 * splits;
  // partial keys, then (256 B aligned) the expanded queries, then one 16-byte record per (padded)
  // query: the distance bounds all CTAs of a format-4 scan share (hamming_tc4.cu)
  *scratch_bytes = (((size_t)(*parts) * (size_t)(nq > 0 ? nq : 1) * 16 + 255) & ~(size_t)255) + pbk_hamming_tc_signs_bytes(nq) +
                   (size_t)tc_pad(nq > 0 ? nq : 1) * 16;
  return PBK_OK;
}

}  // extern "C"
