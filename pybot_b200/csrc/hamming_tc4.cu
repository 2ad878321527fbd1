// Brute-force Hamming kNN (k = 2) on the FP4 path of the 5th-generation tensor cores:
// tcgen05.mma kind::mxf4 (block-scaled e2m1, K = 64 per instruction, twice the int8 rate) with
// every block scale 2^0 - the second engine behind pbk_hamming_tc_* (format 4, the default; the
// kind::i8 engine of hamming_tc.cu is format 8).  Same contract as hamming_knn2_kernel
// (hamming.cu): per query the two smallest (distance, global train row) keys, bit-identical to
// cv2.BFMatcher.
//
// A 256-bit descriptor is expanded ONCE into 256 e2m1 nibbles: a query into a_k = bit_k ? +1 : -1
// (0x2 / 0xA), a train row into b_k = bit_k ? -1 : +1, so that the fp32 accumulator is EXACTLY
//     sum_k a_k b_k = 2 d - 256        (|sum| <= 256: every partial sum is an exact fp32 integer;
// measured bit for bit on random descriptors, scripts/probes/mxf4_probe.cu).  Rows past the end
// of the database are zero nibbles (+0.0: accumulator 0, distance code 128) and are dropped by
// row number when keys are built.  The block scale factors (UE8M0, one per 32 elements of A and
// of B, read by the MMA from tensor memory) are the constant 0x7F = 2^0 in every byte of 64 TMEM
// columns next to the accumulators, so no scale-factor layout is involved.
//
// Epilogue.  Accumulators are fp32, one column per register (no packed form), so the item
// pre-filter is a tree of 3-INPUT minima (FMNMX3 issues at the 2-input rate on sm_100a: 32
// instructions for 64 columns = 0.5 per pair, scripts/probes/minmax_probe.cu); only when the
// smallest value of an item is below the query's current second-best distance - rows are scanned
// in ascending order, so a tie can never enter - are the 64 keys of the item built and inserted.
//
// Data layout: as hamming_tc.cu with 128-byte rows - a tile is 128 rows = 8 planes (one per
// 16-byte K chunk = 32 elements) of 128 x 16 bytes:
//     byte address of element k of row r  =  (r / 128) * 16384 + (k / 32) * 2048 + (r % 128) * 16 + (k % 32) / 2
// (low nibble = even element; any fixed order works as long as queries and rows share it),
// K-major, no swizzle, LBO 2048 B, SBO 128 B; a tile is one contiguous 16 KB span (one 1-D bulk
// copy).  The expanded database is 4 x the compact one (2.56 GB for the 20 M rows of C3).
//
// Kernel.  One CTA per SM, grid = query groups of 512 x database splits.  64 KB of expanded
// queries stay in shared memory; the split streams through an 8-stage ring of 16 KB tiles.
// Work item = (tile, query block): 4 UTCOMMA (M 128, N 128, K 64) into one of THREE 128-column
// accumulator slots (columns 0-383; 384-447 hold the scale bytes).
//   warp 19      producer: bulk copies of the query blocks and of the train tiles
//   warps 16-18  one lane each issues the MMAs of ONE accumulator slot (items slot, slot + 3, ...;
//                item = 4 tile + query block) - an item lasts ~275 clocks on the tensor pipe and one
//                thread needs longer than that for its barrier round trips, descriptors and commits
//   warps 0-15   epilogue in two sets of 8 (even / odd items): warp w owns TMEM lanes
//                32 (w % 4) .. +31 and columns 64 ((w / 4) % 2) .. +63 of a slot; a slot is
//                released as soon as its values are in registers, then the item is selected
#include "hamming_tc.cuh"

namespace pbk {

constexpr int kT4Chunks = kTc4RowBytes / 16;        // 8 K chunks of 32 elements
constexpr int kT4KSteps = kTc4RowBytes / 32;        // 4 UTCOMMA per item (K = 64 elements = 32 bytes)
constexpr int kT4TileBytes = kTcN * kTc4RowBytes;   // 16 KB
constexpr int kT4BlockBytes = kTcM * kTc4RowBytes;  // 16 KB
#ifndef PBK_T4_STAGES
#define PBK_T4_STAGES 8
#endif
constexpr int kT4Stages = PBK_T4_STAGES;
#ifndef PBK_T4_SPIN_EPI
#define PBK_T4_SPIN_EPI 0
#endif
constexpr int kT4Slots = 3;
constexpr uint32_t kT4NoBound = 511u;              // "no second-best yet": above every distance
#ifndef PBK_T4_SHARE
#define PBK_T4_SHARE 1
#endif
#ifndef PBK_T4_REFRESH
#define PBK_T4_REFRESH 0
#endif
// a stale bound corrects itself: it lets a candidate through, and the exact path re-reads it; an
// additional periodic re-read (every PBK_T4_REFRESH tiles) costs more hot-path instructions than
// the entries it saves (measured with 4 and 16)
constexpr int kT4Refresh = PBK_T4_REFRESH;
constexpr int kT4SfCol = 384;                       // first of 64 TMEM columns of 0x7F scale bytes
constexpr int kT4EpiWarps = 16;                     // two sets of 8 (even / odd items), see the file header
constexpr int kT4MmaWarp = kT4EpiWarps, kT4MmaWarps = kT4Slots, kT4ProducerWarp = kT4MmaWarp + kT4MmaWarps;   // one issuing warp per accumulator slot
constexpr int kT4Threads = (kT4ProducerWarp + 1) * 32;       // 640
constexpr int kT4SmemA = kTcQGroup * kTc4RowBytes;  // 65536
constexpr int kT4SmemB = kT4Stages * kT4TileBytes;  // 131072
constexpr int kT4SmemBars = (2 * kT4Stages + 3 * kT4Slots + 1) * 8;
constexpr int kT4SmemBytes = kT4SmemA + kT4SmemB + kT4SmemBars + 16;

// ------------------------------------------------------------ expansion: bits -> e2m1 nibbles
// thread = (row, 16-byte K chunk): 4 descriptor bytes in, 32 nibbles out.
// QUERY: bit 1 -> +1 (0x2), 0 -> -1 (0xA); train: 1 -> -1, 0 -> +1; rows past the end: zeros.
template <bool QUERY>
__global__ void __launch_bounds__(256)
hamming_tc4_expand_kernel(const uint8_t* __restrict__ d, long long n, long long first_row, long long n_padded,
                          uint4* __restrict__ signs, uint32_t* __restrict__ gd2) {
  const long long i = first_row * kT4Chunks + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_padded * kT4Chunks) return;
  const long long tile = i / (kTcN * kT4Chunks);      // 1024 pieces per tile: [K chunk][row of the tile]
  const int in_t = (int)(i - tile * (kTcN * kT4Chunks));
  const int kc = in_t >> 7, rt = in_t & 127;
  const long long r = tile * kTcN + rt;
  if (QUERY && kc == 0)                               // the scan's shared bounds (tc4_bound) start unset
    reinterpret_cast<uint4*>(gd2)[r] = make_uint4(kT4NoBound, kT4NoBound, kT4NoBound, kT4NoBound);
  uint4 o = make_uint4(0, 0, 0, 0);
  if (r < n) {
    // bit j of a byte -> bit 4 j + 3 (the sign of nibble j) over the constant magnitude 1.0 (0x2)
    auto eight = [](unsigned b) -> unsigned {
      unsigned x = (b | (b << 12)) & 0x000F000Fu;
      x = (x | (x << 6)) & 0x03030303u;
      x = (x | (x << 3)) & 0x11111111u;
      return 0x22222222u | ((QUERY ? (x ^ 0x11111111u) : x) << 3);
    };
    const uint8_t* src = d + r * 32 + kc * 4;
    o = make_uint4(eight(src[0]), eight(src[1]), eight(src[2]), eight(src[3]));
  }
  signs[i] = o;
}

// instruction descriptor, kind::mxf4: A = B = e2m1 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17,
// scale format UE8M0 (bit 23), M >> 4 at bit 24, scale-factor ids 0, K = 64 (bit 31 = 0)
constexpr uint32_t kT4Idesc = (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | (1u << 23) | ((uint32_t)(kTcM >> 4) << 24);

__device__ __forceinline__ void tc_mma_mxf4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                            uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kT4Idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// 16 TMEM columns of this warp's lane quarter <- one word
__device__ __forceinline__ void tc_st16(uint32_t taddr, uint32_t w) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(w)
               : "memory");
}
#define PBK_F8(b) "=f"(r[b]), "=f"(r[b + 1]), "=f"(r[b + 2]), "=f"(r[b + 3]), "=f"(r[b + 4]), "=f"(r[b + 5]), "=f"(r[b + 6]), "=f"(r[b + 7])
// 64 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld64_f32(uint32_t taddr, float (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : PBK_F8(0), PBK_F8(8), PBK_F8(16), PBK_F8(24), PBK_F8(32), PBK_F8(40), PBK_F8(48), PBK_F8(56)
      : "r"(taddr)
      : "memory");
}
#undef PBK_F8
__device__ __forceinline__ float tc_min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// top-2 of one accumulator half-slot (64 columns, acc = 2 d - 256 as fp32) of this thread's query
// row, folded into (k1, k2).  ONE copy of this code serves every query block and both column
// halves (the kernel is instruction-cache sensitive, see hamming_tc.cu).  tile_base = row in the
// split of the tile's column 0; rows = rows in the split; h = column half of the slot.
__device__ __forceinline__ void tc4_insert(float acc, uint32_t row, uint32_t rows, uint32_t& k1, uint32_t& k2) {
  // 2 d = acc + 256 in [0, 512]: read as the mantissa of acc + (2^23 + 256)
  const uint32_t d2x = __float_as_uint(acc + 8388864.0f) & 0x3FFu;
  const uint32_t key = row >= rows ? kTcNone : (d2x >> 1) * kTcKeyMul + row;
  tc_top2(k1, k2, key);
}

// acc threshold of a query: an element can only matter with acc = 2 d - 256 BELOW it.
//   own scan: rows arrive in ascending order, so only a distance below the current second best
//             d2 = k2 >> 23 (>= the true one; 511 while unset) can enter: acc < 2 d2 - 256;
//   all scans: `bound` is a distance two DISTINCT rows of the database are known to reach, found
//             by any CTA of the launch (tc4_bound): a row beyond it cannot be in the merged top-2
//             whatever its row number - acc <= 2 bound - 256, i.e. acc < 2 bound - 255.
// The partial keys of one CTA are then no longer ITS top-2, but the merge over all of them is
// exactly the top-2 of the database (both global winners pass both tests in their own scans).
// Without the shared bound a warp of a 34 k-row scan (one 8-GPU shard of C3) built keys for a
// quarter of its items; with the second-best word alone for 6 %.
__device__ __forceinline__ float tc4_threshold(uint32_t k2, uint32_t bound = kT4NoBound) {
  return fminf((float)(2 * (int)(k2 >> 23) - 256), (float)(2 * (int)bound - 255));
}
// One 16-byte record per (padded) query, every word an atomicMin target:
//   .x  the smallest SECOND-best distance any warp of the launch has (two rows of one scan)
//   .y  the smallest BEST distance over the scans of column half 0,  .z  the same for column half 1
// The two halves scan disjoint rows, so max(.y, .z) is reached by two distinct rows as well - and
// early in a scan it is far below .x (the best of ~2400 T rows against the best second-best of 74
// scans of 64 T rows each after T tiles).
__device__ __forceinline__ uint32_t tc4_bound(uint4 g) { return min(g.x, max(g.y, g.z)); }

__device__ __forceinline__ void tc4_select(const float (&v)[64], uint32_t tile_base, uint32_t rows, uint32_t h,
                                           uint32_t& k1, uint32_t& k2, float& thr, uint32_t* gd2_q) {
  // minima of seven groups of nine columns (+ column 63), then of the item: 32 FMNMX3 / FMNMX
  float g[7];
#pragma unroll
  for (int i = 0; i < 7; ++i)
    g[i] = tc_min3(tc_min3(v[9 * i], v[9 * i + 1], v[9 * i + 2]), tc_min3(v[9 * i + 3], v[9 * i + 4], v[9 * i + 5]),
                   tc_min3(v[9 * i + 6], v[9 * i + 7], v[9 * i + 8]));
  const float best = fminf(tc_min3(tc_min3(g[0], g[1], g[2]), tc_min3(g[3], g[4], g[5]), g[6]), v[63]);
  if (!__any_sync(0xffffffffu, best < thr)) return;
  // Rare path (some lane of the warp has a candidate; ~a fifth of the items of a 500 k-row split
  // for the 128 queries of a warp): only the groups that hold one are turned into keys - a warp
  // that inserted all 64 columns spent three times the pre-filter's instructions here.
  const uint32_t col0 = tile_base + h * 64u;
#if PBK_T4_SHARE
  // what the other CTAs found meanwhile: requested now, needed after the inserts (an L2 round trip)
  uint4 seen;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(seen.x), "=r"(seen.y), "=r"(seen.z), "=r"(seen.w) : "l"(gd2_q));
  const uint32_t before1 = k1, before2 = k2;
#endif
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    if (__any_sync(0xffffffffu, g[i] < thr)) {
#pragma unroll
      for (int e = 0; e < 9; ++e) tc4_insert(v[9 * i + e], col0 + (uint32_t)(9 * i + e), rows, k1, k2);
    }
  }
  if (__any_sync(0xffffffffu, v[63] < thr)) tc4_insert(v[63], col0 + 63u, rows, k1, k2);
#if PBK_T4_SHARE
  if (k2 != before2) atomicMin(gd2_q, k2 >> 23);    // results unused: fire-and-forget RED.MIN
  if (k1 != before1) atomicMin(gd2_q + 1 + h, k1 >> 23);
  seen.x = min(seen.x, k2 >> 23);
  if (h == 0) seen.y = min(seen.y, k1 >> 23); else seen.z = min(seen.z, k1 >> 23);
  thr = tc4_threshold(k2, tc4_bound(seen));
#else
  thr = tc4_threshold(k2);
#endif
}

__global__ void __launch_bounds__(kT4Threads, 1)
hamming_tc4_kernel(const __grid_constant__ HammingTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kT4SmemA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kT4SmemA + kT4SmemB);
  uint64_t* b_full = bars;                           // [kT4Stages]  train tile landed
  uint64_t* b_empty = bars + kT4Stages;              // [kT4Stages]  the MMAs of all four query blocks that read the tile are done
  // [2][kT4Slots] accumulator slot written, one barrier per (item parity, slot): an epilogue set
  // sees only every other use of a slot, and a parity wait is only valid on CONSECUTIVE phases
  uint64_t* acc_full = bars + 2 * kT4Stages;
  uint64_t* acc_empty = acc_full + 2 * kT4Slots;     // [kT4Slots]   the 8 warps of the item's set have read it
  uint64_t* a_full = acc_empty + kT4Slots;           // query blocks landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_full + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qg = blockIdx.x, split = blockIdx.y;
  const long long row0 = (long long)split * P.rows_per_split;
  long long rows = P.nt - row0;
  if (rows > P.rows_per_split) rows = P.rows_per_split;
  if (rows < 0) rows = 0;
  const int ntiles = (int)((rows + kTcN - 1) / kTcN);

  if (tid == 0) {
    for (int s = 0; s < kT4Stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], kTcQBlocks); }
    for (int s = 0; s < kT4Slots; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_full[kT4Slots + s], 1);
      mbar_init(&acc_empty[s], kTcEpiWarps);
    }
    mbar_init(a_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kT4MmaWarp) {                          // all 512 TMEM columns: one CTA per SM (shared memory bound)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp < 4) {                                    // unit block scales: 0x7F in every byte of columns 384 .. 447
#pragma unroll
    for (int c = 0; c < 4; ++c) tc_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kT4SfCol + 16 * c), 0x7F7F7F7Fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == kT4ProducerWarp) {
    // ================================================================= producer
    if (lane == 0 && ntiles > 0) {
      mbar_expect_tx(a_full, kT4SmemA);
      for (int b = 0; b < kTcQBlocks; ++b)
        tma_load_1d(sA + b * kT4BlockBytes, P.q_signs + ((size_t)qg * kTcQGroup + (size_t)b * kTcM) * kTc4RowBytes,
                    kT4BlockBytes, a_full);
      const int8_t* src = P.t_signs + (size_t)row0 * kTc4RowBytes;
      for (int t = 0; t < kTcPrefetch && t < ntiles; ++t)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)t * kT4TileBytes), "r"(kT4TileBytes) : "memory");
      for (int t = 0; t < ntiles && !(P.dbg & 8); ++t) {
        const int s = t % kT4Stages;
        if (t + kTcPrefetch < ntiles)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)(t + kTcPrefetch) * kT4TileBytes), "r"(kT4TileBytes) : "memory");
        if (t >= kT4Stages) mbar_wait_parked(&b_empty[s], (uint32_t)(((t / kT4Stages) - 1) & 1));
        mbar_expect_tx(&b_full[s], kT4TileBytes);
        tma_load_1d(sB + s * kT4TileBytes, src + (size_t)t * kT4TileBytes, kT4TileBytes, &b_full[s]);
      }
    }
  } else if (warp >= kT4MmaWarp) {
    // =============================================================== MMA issuers
    // Issuer j owns accumulator slot j: items j, j + 3, j + 6, ... (item = 4 tile + query block).
    // One thread per slot waits on CONSECUTIVE phases of the slot's acc_empty barrier - with the
    // items of a slot spread over several threads a thread could wait for a phase two completions
    // ahead of the barrier, which a parity wait cannot express (it returns at once).
    // The WHOLE warp runs this loop with warp-uniform values (waits included) and one elected lane
    // issues the tcgen05 instructions: under `if (lane == 0)` the compiler cannot prove the
    // descriptors uniform and wraps every UTCOMMA / UTCBAR in an ELECT + 5 x R2UR.BROADCAST
    // waterfall loop - ~170 instructions and ~300 clocks from "slot free" to the last MMA of an item.
    if (ntiles > 0) {
      const int slot = __shfl_sync(0xffffffffu, warp - kT4MmaWarp, 0);
      mbar_wait(a_full, 0);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      const uint32_t d = tmem_base + (uint32_t)(slot * kTcN);
      const uint32_t sfa = tmem_base + kT4SfCol, sfb = tmem_base + kT4SfCol + 16;
      const int nitems = ntiles * kTcQBlocks;
      int use = 0;
#pragma unroll 1
      for (int item = slot; item < nitems; item += kT4Slots, ++use) {
        const int t = item >> 2, mb = item & 3;
        const int s = t % kT4Stages;
        const uint64_t ad = tc_smem_desc(a_addr + mb * kT4BlockBytes), bd = tc_smem_desc(b_addr + s * kT4TileBytes);
        uint64_t* const full_bar = &acc_full[(item & 1) * kT4Slots + slot];
        if (!(P.dbg & 8)) mbar_wait_parked(&b_full[s], (uint32_t)((t / kT4Stages) & 1));
        if (use > 0) mbar_wait_parked(&acc_empty[slot], (uint32_t)((use - 1) & 1));
        tc_fence_after();
        if (tc_elect_one()) {
          if (!(P.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < kT4KSteps; ++k)            // K = 64 elements = two planes: +4096 B = +256 in the address field
              tc_mma_mxf4(d, ad + (uint64_t)(k * 256), bd + (uint64_t)(k * 256), k > 0 ? 1u : 0u, sfa, sfb);
          }
          tc_commit(full_bar);
          tc_commit(&b_empty[s]);                     // one of the four arrivals (one per query block) that free the stage
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= epilogue
    // Warp w: TMEM lane quarter w & 3, column half (w >> 2) & 1 of a slot, item set w >> 3: set 0
    // takes the even items (query blocks 0 and 2 of every tile), set 1 the odd ones (blocks 1, 3).
    // While one set selects item i the other is parked on item i + 1's barrier, so a finished
    // accumulator is read (and its slot handed back) at once: with THREE slots and ~275 clocks of
    // tensor time per item the slot turn-around is what the pipe waits for, and one set of 8
    // warps doing fetch / select in turn answered a finished item up to a whole selection late.
    const int lg = warp & 3, h = (warp >> 2) & 1, set = warp >> 3;
    uint32_t ka1 = kTcNone, ka2 = kTcNone, kb1 = kTcNone, kb2 = kTcNone;   // query blocks set, set + 2
    float tha = tc4_threshold(kTcNone), thb = tha;    // their acc thresholds (kept: the hot path is one compare)
    // this lane's query of block `set` in the shared bound array; block set + 2 is 256 further
    const uint32_t gq = (uint32_t)(qg * kTcQGroup + set * kTcM + lg * 32 + lane);
    const int nitems = ntiles * kTcQBlocks;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(h * 64);
    float v[64];
    int slot = set, use = 0;                          // item % 3, item / 3 of the current item (item = set, set + 2, ...)
    auto step = [&](int item, uint32_t& k1, uint32_t& k2, float& thr, uint32_t* gd2_q) {
      const uint32_t taddr = lane_addr + (uint32_t)(slot * kTcN);
      uint64_t* const empty_bar = &acc_empty[slot];
#if PBK_T4_SPIN_EPI
      mbar_wait(&acc_full[set * kT4Slots + slot], (uint32_t)((use >> 1) & 1));
#else
      mbar_wait_parked(&acc_full[set * kT4Slots + slot], (uint32_t)((use >> 1) & 1));   // the (set, slot) pair recurs every 6 items
#endif
      tc_fence_after();
      if (!(P.dbg & 4)) tc_ld64_f32(taddr, v);
      tc_ld_wait();                                   // the values are in registers: the slot goes back to its issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar);
      if (!(P.dbg & 1)) tc4_select(v, (uint32_t)(item >> 2) * kTcN, (uint32_t)rows, (uint32_t)h, k1, k2, thr, gd2_q);
      else k1 ^= __float_as_uint(v[0]) ^ __float_as_uint(v[63]) ^ __float_as_uint(v[17]);
      slot += 2;                                      // the set's next item is item + 2
      if (slot >= kT4Slots) { slot -= kT4Slots; ++use; }
    };
#pragma unroll 1
    for (int item = set; item < nitems; item += 4) {  // nitems is a multiple of 4
      step(item, ka1, ka2, tha, P.gd2 + 4 * gq);
      step(item + 2, kb1, kb2, thb, P.gd2 + 4 * (gq + 2 * kTcM));
#if PBK_T4_SHARE && PBK_T4_REFRESH
      if (((item >> 2) & (kT4Refresh - 1)) == kT4Refresh - 1) {      // bounds the other CTAs published since
        tha = tc4_threshold(ka2, *reinterpret_cast<volatile uint32_t*>(P.gd2 + 4 * gq));
        thb = tc4_threshold(kb2, *reinterpret_cast<volatile uint32_t*>(P.gd2 + 4 * (gq + 2 * kTcM)));
      }
#endif
    }
    const unsigned long long gbase = P.train_base + (unsigned long long)row0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int mb = set + 2 * c;
      const long long qi = (long long)qg * kTcQGroup + mb * kTcM + lg * 32 + lane;
      if (qi < P.nq) {
        unsigned long long o[2];
        const uint32_t kk[2] = {c ? kb1 : ka1, c ? kb2 : ka2};
#pragma unroll
        for (int e = 0; e < 2; ++e)
          o[e] = (kk[e] == kTcNone) ? PBK_KEY_NONE
                                    : (((unsigned long long)(kk[e] / kTcKeyMul) << 32) | (gbase + (kk[e] % kTcKeyMul)));
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.partial + ((long long)(split * 2 + h) * P.nq + qi) * 2);
        *dst = make_ulonglong2(o[0], o[1]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kT4MmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------ FP4 tensor peak probe
// Measurement only: one CTA per SM issues `iters` x 4 UTCOMMA (M 128, N 128, K 64) on whatever its
// shared memory holds - no data movement, no epilogue (pbk_hamming_tc_peak_probe, format 4).
__global__ void __launch_bounds__(128, 1) hamming_tc4_peak_kernel(int iters, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * kT4BlockBytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x22222222u;
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
#pragma unroll
  for (int c = 0; c < 4; ++c) tc_st16(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kT4SfCol + 16 * c), 0x7F7F7F7Fu);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && lane == 0) {
    const uint64_t ad = tc_smem_desc(smem_u32(smem)), bd = tc_smem_desc(smem_u32(smem) + kT4BlockBytes);
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tm + (uint32_t)((it % kT4Slots) * kTcN);
#pragma unroll
      for (int k = 0; k < kT4KSteps; ++k)
        tc_mma_mxf4(d, ad + (uint64_t)(k * 256), bd + (uint64_t)(k * 256), k > 0 ? 1u : 0u, tm + kT4SfCol, tm + kT4SfCol + 16);
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

int hamming_tc4_expand(const uint8_t* d, long long n, long long first_row, long long n_padded, int8_t* signs, cudaStream_t s) {
  const long long blocks = ((n_padded - first_row) * kT4Chunks + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "too many rows for one call");
  hamming_tc4_expand_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(d, n, first_row, n_padded, reinterpret_cast<uint4*>(signs), nullptr);
  PBK_CHECK_LAUNCH("hamming_tc4_expand_kernel");
  return PBK_OK;
}

int hamming_tc4_peak(int iters, uint32_t* sink, const DeviceInfo* di, cudaStream_t s) {
  constexpr int smem = 2 * kT4BlockBytes;
  static KernelCache kc;
  int rc = kc.setup(di, hamming_tc4_peak_kernel, 128, smem, nullptr);
  if (rc) return rc;
  hamming_tc4_peak_kernel<<<di->sm_count, 128, smem, s>>>(iters, sink);
  PBK_CHECK_LAUNCH("hamming_tc4_peak_kernel");
  return PBK_OK;
}

int hamming_tc4_scan(const uint8_t* q, long long nq, const int8_t* t_signs, long long nt, unsigned long long train_base,
                     int8_t* q_signs, unsigned long long* partial, int splits, long long rps, const DeviceInfo* di,
                     cudaStream_t s) {
  const long long nqp = tc_pad(nq);
  uint32_t* gd2 = reinterpret_cast<uint32_t*>(q_signs + (size_t)nqp * kTc4RowBytes);     // pbk_hamming_tc_plan sizes the scratch for it
  hamming_tc4_expand_kernel<true><<<(unsigned)((nqp * kT4Chunks + 255) / 256), 256, 0, s>>>(q, nq, 0, nqp, reinterpret_cast<uint4*>(q_signs), gd2);
  PBK_CHECK_LAUNCH("hamming_tc4_expand_kernel");
  static KernelCache kc;
  int rc = kc.setup(di, hamming_tc4_kernel, kT4Threads, kT4SmemBytes, nullptr);
  if (rc) return rc;
  HammingTcParams P;
  P.q_signs = q_signs; P.t_signs = t_signs; P.partial = partial; P.gd2 = gd2;
  P.nq = nq; P.nt = nt; P.rows_per_split = rps; P.train_base = train_base; P.splits = splits;
  P.dbg = 0;
  if (const char* e = getenv("PBK_TC_DEBUG")) P.dbg = atoi(e);
  const dim3 grid((unsigned)((nq + kTcQGroup - 1) / kTcQGroup), (unsigned)splits);
  hamming_tc4_kernel<<<grid, kT4Threads, kT4SmemBytes, s>>>(P);
  PBK_CHECK_LAUNCH("hamming_tc4_kernel");
  return PBK_OK;
}

}  // namespace pbk
