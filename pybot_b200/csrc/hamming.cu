// Brute-force Hamming kNN (k=2) over 256-bit ORB descriptors with
// cv2.BFMatcher semantics.
//
// Replaces cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, k=2) + Lowe ratio +
// mutual check.  The reference has no call site for it; the
// semantics are pinned against OpenCV 4.13: per query the two smallest
// distances, ties broken by the LOWEST train index (imgIdx, trainIdx order
// for collections == global row order).
//
// Design (integer pipe; tensor cores are not used - there is no b1 tcgen05
// kind):
//  * grid = (query tiles of 512) x (train splits).  A CTA of 128 threads keeps
//    4 queries per thread in registers (8 x u32 each).
//  * the CTA's train split streams through a 4-stage shared-memory ring of
//    256-row (8 KB) tiles filled by TMA bulk copies (cp.async.bulk ->
//    UBLKCP) that complete on mbarriers; the query tile is staged the same way.
//  * every lane reads the same train row (broadcast LDS.128 x2) and scores
//    it against its 4 queries: XOR (LOP3) + POPC + integer adds.
//    VARIANT 0 is the textbook 8 XOR + 8 POPC.  VARIANT 1 (default) compresses
//    the eight words with carry-save adders (LOP3 full adders) so only 4 POPC
//    per pair are issued - the popc pipe is the narrow one (16 lanes/clk/SM vs
//    64 for LOP3).  To keep the LOP3 count at 13 per pair, each staged train
//    row is rewritten in shared memory (once per CTA, amortised over 512
//    queries) to [t0 t1 t3 t4 t7 | t0^t1^t2 | t3^t4^t5 | t0^..^t6] and every
//    query is held in the same derived form: a full adder's sum bit is then one
//    XOR of precomputed words and its carry is maj(x, y, x^y^sum) - one LOP3.
//  * top-2 selection is a 3-instruction min/max network on a packed 32-bit key
//    (distance * (2^23+1) + local row): rows are scanned in ascending order and the
//    key order is lexicographic, so ties resolve to the lowest train index.
//    VARIANT 2 (default for the 4-queries-per-thread shape) runs that network on
//    16-bit keys, two queries per register (VIMNMX.U16x2: 1.5 instead of 3 ALU
//    instructions per pair), over 128-row blocks whose winners are folded into the
//    32-bit keys; 966 -> 996 Gpairs/s on a 2.5 M-row shard.  With selection removed
//    altogether the loop runs only 3 % faster again: 4 POPC per pair on the 16-lane
//    XU pipe is then the bound.
//  * each CTA emits (distance << 32 | global row) u64 keys per query; a merge
//    kernel reduces the splits (and, multi-GPU, the all-gathered shards) with
//    the same unsigned min - bit-identical to a single ascending scan.
#include "pbk_common.cuh"

namespace pbk {

constexpr int kHThreads = 128;
constexpr int kRQ = 4;                          // queries per thread (throughput shape)
constexpr int kQTile = kHThreads * kRQ;         // 512 queries per CTA
// Per-frame problems (C1: 2000 x 2000) give the throughput shape 4 query tiles x 8
// stage-sized splits = 32 CTAs on 148 SMs, each scoring 256 rows x 4 queries per
// thread (26 us for 4 M pairs).  The latency shape keeps ONE query per thread:
// 4x the CTAs, a quarter of the work per thread.
constexpr int kRQSmall = 1;
constexpr int kTR = 256;                        // train rows per stage
constexpr int kStages = 4;
constexpr int kStageBytes = kTR * 32;
constexpr int kKeyShift = 23;                   // local row index < 2^23
// packed key = distance * kKeyMul + local_row.  An odd multiplier (not a power of
// two) keeps the key build on the FMA pipe as IMAD; a shift would be an ALU LEA
// and the ALU pipe is the binding one.
constexpr uint32_t kKeyMul = (1u << kKeyShift) + 1u;
constexpr uint32_t kLocalNone = 0xFFFFFFFFu;
constexpr long long kMaxSplitRows = 1ll << kKeyShift;
constexpr int kHMaxWorld = 16;                  // ranks of the fused multi-GPU exchange

struct HammingParams {
  const uint8_t* q;
  const uint8_t* t;
  unsigned long long* partial;   // (splits, nq, 2)
  long long nq, nt;
  long long rows_per_split;      // multiple of kTR
  unsigned long long train_base;
  int splits;
};

// ------------------------------------------------------------------ scoring
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// carry of a full adder given two inputs and the SUM bit: maj(a, b, a^b^s)
__device__ __forceinline__ uint32_t carry_from_sum(uint32_t a, uint32_t b, uint32_t s) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
  return r;
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;       // keeps the weighted popc sum on the FMA pipe (IMAD), off the ALU pipe
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// derived 8-word form of a descriptor (VARIANT 1): [w0 w1 w3 w4 | w7 Sa Sb Sabc]
__device__ __forceinline__ void derive_row(uint4& a, uint4& b) {
  const uint32_t sa = xor3(a.x, a.y, a.z), sb = xor3(a.w, b.x, b.y);
  const uint32_t sabc = xor3(sa, sb, b.z);
  a = make_uint4(a.x, a.y, a.w, b.x);
  b = make_uint4(b.w, sa, sb, sabc);
}

// packed key (distance << 23) + rowkey of one query against one train row
template <int VARIANT>
__device__ __forceinline__ uint32_t pair_key(const uint32_t* q, const uint4& a, const uint4& b,
                                             uint32_t rowkey) {
  if constexpr (VARIANT == 0) {
    const uint32_t d = __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
                       __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
    return mad_u32(d, kKeyMul, rowkey);
  } else {
    // q and (a,b) in derived form.  ones: sc, x7 ; twos: sd ; fours: cd
    const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x3 = q[2] ^ a.z, x4 = q[3] ^ a.w;
    const uint32_t x7 = q[4] ^ b.x;
    const uint32_t sa = q[5] ^ b.y, sb = q[6] ^ b.z, sc = q[7] ^ b.w;
    const uint32_t ca = carry_from_sum(x0, x1, sa);       // maj(x0, x1, x2)
    const uint32_t cb = carry_from_sum(x3, x4, sb);       // maj(x3, x4, x5)
    const uint32_t cc = carry_from_sum(sa, sb, sc);       // maj(sa, sb, x6)
    const uint32_t sd = xor3(ca, cb, cc), cd = maj3(ca, cb, cc);
    uint32_t key = mad_u32(__popc(x7), kKeyMul, rowkey);
    key = mad_u32(__popc(sc), kKeyMul, key);
    key = mad_u32(__popc(sd), 2u * kKeyMul, key);
    return mad_u32(__popc(cd), 4u * kKeyMul, key);
  }
}

// VARIANT 2: two queries' keys ride in one register as u16 halves so that the top-2 network is
// 3 VIMNMX.U16x2 per TWO pairs.  A 16-bit key = distance * 129 + row-in-block (row < 128) keeps
// the (distance, row) order for distances up to 256 (max 33151; 0xFFFF = none), so the tile is
// scored in 128-row blocks whose packed top-2 are folded into the 32-bit keys once per block.
constexpr int kBlk = 128;
constexpr uint32_t kMul16 = kBlk + 1;
// popc weights of one derived-form pair, summed with IMADs (FMA pipe) onto `acc`
__device__ __forceinline__ uint32_t pair_acc(const uint32_t* q, const uint4& a, const uint4& b, uint32_t mul,
                                             uint32_t acc) {
  const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x3 = q[2] ^ a.z, x4 = q[3] ^ a.w;
  const uint32_t x7 = q[4] ^ b.x;
  const uint32_t sa = q[5] ^ b.y, sb = q[6] ^ b.z, sc = q[7] ^ b.w;
  const uint32_t ca = carry_from_sum(x0, x1, sa);
  const uint32_t cb = carry_from_sum(x3, x4, sb);
  const uint32_t cc = carry_from_sum(sa, sb, sc);
  const uint32_t sd = xor3(ca, cb, cc), cd = maj3(ca, cb, cc);
  acc = mad_u32(__popc(x7), mul, acc);
  acc = mad_u32(__popc(sc), mul, acc);
  acc = mad_u32(__popc(sd), 2u * mul, acc);
  return mad_u32(__popc(cd), 4u * mul, acc);
}
__device__ __forceinline__ void top2_insert_u16x2(uint32_t& k1, uint32_t& k2, uint32_t key) {
  uint32_t mx;
  asm("max.u16x2 %0, %1, %2;" : "=r"(mx) : "r"(k1), "r"(key));
  asm("min.u16x2 %0, %1, %2;" : "=r"(k1) : "r"(k1), "r"(key));
  asm("min.u16x2 %0, %1, %2;" : "=r"(k2) : "r"(k2), "r"(mx));
}

__device__ __forceinline__ void top2_insert(uint32_t& k1, uint32_t& k2, uint32_t key) {
  const uint32_t mx = max(k1, key);
  k1 = min(k1, key);
  k2 = min(k2, mx);
}

// Scores `nrows` staged rows (FULL: exactly kTR) against the thread's 4 queries.
template <int VARIANT, bool FULL, int RQ>
__device__ __forceinline__ void score_tile(const uint4* __restrict__ tile, int nrows, uint32_t rowkey,
                                           const uint32_t (&qr)[RQ][8], uint32_t (&k1)[RQ],
                                           uint32_t (&k2)[RQ]) {
  if constexpr (VARIANT == 3) {
    // probe only: packed keys reduced with ONE min per two pairs (no second-best) - the ceiling
    // of the XOR/CSA/POPC/IMAD part of the loop
    uint32_t rk = rowkey;
#pragma unroll 4
    for (int j = 0; j < kTR; ++j) {
      const uint4 a = tile[2 * j], b = tile[2 * j + 1];
#pragma unroll
      for (int h = 0; h < RQ / 2; ++h) {
        const uint32_t lo = pair_acc(qr[2 * h], a, b, kMul16, rk);
        const uint32_t key = pair_acc(qr[2 * h + 1], a, b, kMul16 << 16, lo);
        asm("min.u16x2 %0, %1, %2;" : "=r"(k1[2 * h]) : "r"(k1[2 * h]), "r"(key));
      }
      rk += 0x10001u;
    }
    return;
  }
  if constexpr (VARIANT == 2 && RQ % 2 == 0) {
    for (int b0 = 0; b0 < kTR; b0 += kBlk) {
      const int n = FULL ? kBlk : (nrows - b0 < kBlk ? nrows - b0 : kBlk);
      if (!FULL && n <= 0) break;
      uint32_t p1[RQ / 2], p2[RQ / 2];
#pragma unroll
      for (int h = 0; h < RQ / 2; ++h) p1[h] = p2[h] = 0xFFFFFFFFu;
      const uint4* blk = tile + 2 * b0;
      uint32_t rk = 0;                       // row-in-block in both halves
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const uint4 a = blk[2 * j], b = blk[2 * j + 1];
#pragma unroll
        for (int h = 0; h < RQ / 2; ++h) {
          const uint32_t lo = pair_acc(qr[2 * h], a, b, kMul16, rk);
          top2_insert_u16x2(p1[h], p2[h], pair_acc(qr[2 * h + 1], a, b, kMul16 << 16, lo));
        }
        rk += 0x10001u;
      }
      // fold the block's packed top-2 into the 32-bit keys (0xFFFF -> none, which never wins)
#pragma unroll
      for (int r = 0; r < RQ; ++r) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t pk = c == 0 ? p1[r / 2] : p2[r / 2];
          const uint32_t v = (r & 1) ? (pk >> 16) : (pk & 0xFFFFu);
          const uint32_t d = v / kMul16;
          const uint32_t key = v == 0xFFFFu ? kLocalNone : d * (kKeyMul - kMul16) + v + rowkey + (uint32_t)b0;
          top2_insert(k1[r], k2[r], key);
        }
      }
    }
    return;
  }
  const int n = FULL ? kTR : nrows;
#pragma unroll 4
  for (int j = 0; j < n; ++j) {
    const uint4 a = tile[2 * j], b = tile[2 * j + 1];
#pragma unroll
    for (int r = 0; r < RQ; ++r) top2_insert(k1[r], k2[r], pair_key<VARIANT>(qr[r], a, b, rowkey));
    ++rowkey;
  }
}

template <int VARIANT, int RQ = kRQ>
__global__ void __launch_bounds__(kHThreads, 4)
hamming_knn2_kernel(const __grid_constant__ HammingParams P) {
  constexpr int kRQ = RQ, kQTile = kHThreads * RQ;      // shadow the throughput-shape constants
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;                                   // kStages * kStageBytes
  uint8_t* qstage = smem + kStages * kStageBytes;         // kQTile * 32
  uint64_t* full = reinterpret_cast<uint64_t*>(qstage + kQTile * 32);   // kStages + 1 barriers

  const int tid = threadIdx.x;
  const long long q0 = (long long)blockIdx.x * kQTile;
  const int split = blockIdx.y;
  const long long row0 = (long long)split * P.rows_per_split;
  long long rows = P.nt - row0;
  if (rows > P.rows_per_split) rows = P.rows_per_split;
  if (rows < 0) rows = 0;
  const int nstage_total = (int)((rows + kTR - 1) / kTR);
  const int nq_tile = (int)((P.nq - q0) < kQTile ? (P.nq - q0) : kQTile);

  if (tid == 0) {
    for (int s = 0; s <= kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint8_t* tsplit = P.t + row0 * 32;
  auto issue = [&](int it) {
    const int s = it % kStages;
    const long long r = (long long)it * kTR;
    const uint32_t bytes = (uint32_t)(((rows - r) < kTR ? (rows - r) : kTR) * 32);
    mbar_expect_tx(&full[s], bytes);
    tma_load_1d(ring + s * kStageBytes, tsplit + r * 32, bytes, &full[s]);
  };
  if (tid == 0) {
    mbar_expect_tx(&full[kStages], (uint32_t)nq_tile * 32u);
    tma_load_1d(qstage, P.q + q0 * 32, (uint32_t)nq_tile * 32u, &full[kStages]);
    for (int it = 0; it < kStages - 1 && it < nstage_total; ++it) issue(it);
  }

  // this thread's queries: q0 + r*128 + tid
  uint32_t qr[kRQ][8];
  mbar_wait(&full[kStages], 0);
#pragma unroll
  for (int r = 0; r < kRQ; ++r) {
    const uint4* src = reinterpret_cast<const uint4*>(qstage + (r * kHThreads + tid) * 32);
    const uint4 a = src[0], b = src[1];
    uint4 da = a, db = b;
    if (VARIANT >= 1) derive_row(da, db);
    qr[r][0] = da.x; qr[r][1] = da.y; qr[r][2] = da.z; qr[r][3] = da.w;
    qr[r][4] = db.x; qr[r][5] = db.y; qr[r][6] = db.z; qr[r][7] = db.w;
  }
  uint32_t k1[kRQ], k2[kRQ];
#pragma unroll
  for (int r = 0; r < kRQ; ++r) k1[r] = k2[r] = kLocalNone;

  for (int it = 0; it < nstage_total; ++it) {
    const int s = it % kStages;
    // refill the slot freed in the previous iteration (all threads passed the
    // __syncthreads at its end)
    if (tid == 0 && it + kStages - 1 < nstage_total) issue(it + kStages - 1);
    mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
    const long long r_base = (long long)it * kTR;
    const int nrows = (int)((rows - r_base) < kTR ? (rows - r_base) : kTR);
    uint4* tile = reinterpret_cast<uint4*>(ring + s * kStageBytes);
    if (VARIANT >= 1) {
      // rewrite the staged rows to derived form, two rows per thread, in place
#pragma unroll
      for (int h = 0; h < kTR / kHThreads; ++h) {
        const int row = h * kHThreads + tid;
        if (row < nrows) {
          uint4 a = tile[2 * row], b = tile[2 * row + 1];
          derive_row(a, b);
          tile[2 * row] = a;
          tile[2 * row + 1] = b;
        }
      }
      // these generic-proxy stores are later overwritten by the TMA refill of the slot (async
      // proxy): the writer orders them across the proxies before the barriers that precede it
      fence_proxy_async_smem();
      __syncthreads();
    }
    if (nrows == kTR) score_tile<VARIANT, true, kRQ>(tile, kTR, (uint32_t)r_base, qr, k1, k2);
    else score_tile<VARIANT, false, kRQ>(tile, nrows, (uint32_t)r_base, qr, k1, k2);
    __syncthreads();   // everyone is done with slot s before it is refilled
  }

  // emit (distance << 32 | global row) keys
  const unsigned long long gbase = P.train_base + (unsigned long long)row0;
#pragma unroll
  for (int r = 0; r < kRQ; ++r) {
    const long long qi = q0 + r * kHThreads + tid;
    if (qi < P.nq) {
      unsigned long long o[2];
      const uint32_t kk[2] = {k1[r], k2[r]};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        o[c] = (kk[c] == kLocalNone)
                   ? PBK_KEY_NONE
                   : (((unsigned long long)(kk[c] / kKeyMul) << 32) | (gbase + (kk[c] % kKeyMul)));
      }
      ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.partial + ((long long)split * P.nq + qi) * 2);
      *dst = make_ulonglong2(o[0], o[1]);
    }
  }
}

// Top-2 (unsigned min) of the `parts` partial key pairs of query i, by the kMergeLanes
// consecutive lanes that share the query: lane `sub` takes the parts sub, sub + 8, ... (its
// loads are independent and in flight together), then three shuffle steps combine the lanes.
// A tensor-core scan leaves 74 partial pairs per query; one thread walking them one dependent
// 16-byte load after the other took 22 us for 2000 queries.  Every lane of the warp must call it
// (i >= nq: no loads); the result is valid in all kMergeLanes lanes of the query.
constexpr int kMergeLanes = 8;
__device__ __forceinline__ void top2_insert(unsigned long long& b1, unsigned long long& b2, unsigned long long v) {
  const unsigned long long mx = max(b1, v);
  b1 = min(b1, v);
  b2 = min(b2, mx);
}
__device__ __forceinline__ void merge_parts(const unsigned long long* __restrict__ partial, int parts, long long nq,
                                            long long i, int sub, unsigned long long& b1, unsigned long long& b2) {
  b1 = PBK_KEY_NONE;
  b2 = PBK_KEY_NONE;
  if (i < nq) {
    for (int p0 = sub; p0 < parts; p0 += 4 * kMergeLanes) {
      ulonglong2 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int p = p0 + k * kMergeLanes;
        v[k] = p < parts ? *reinterpret_cast<const ulonglong2*>(partial + ((long long)p * nq + i) * 2)
                         : make_ulonglong2(PBK_KEY_NONE, PBK_KEY_NONE);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        top2_insert(b1, b2, v[k].x);
        top2_insert(b1, b2, v[k].y);
      }
    }
  }
#pragma unroll
  for (int o = 1; o < kMergeLanes; o <<= 1) {
    const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, b1, o), o2 = __shfl_xor_sync(0xffffffffu, b2, o);
    top2_insert(b1, b2, o1);
    top2_insert(b1, b2, o2);
  }
}
static inline unsigned merge_blocks(long long nq) { return (unsigned)((nq * kMergeLanes + 255) / 256); }

// partial (parts, nq, 2) -> keys (nq, 2): global top-2 by unsigned min.
__global__ void __launch_bounds__(256)
hamming_merge_kernel(const unsigned long long* __restrict__ partial, int parts, long long nq,
                     unsigned long long* __restrict__ keys) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / kMergeLanes;
  const int sub = (int)(t % kMergeLanes);
  unsigned long long b1, b2;
  merge_parts(partial, parts, nq, i, sub, b1, b2);
  if (i < nq && sub == 0) *reinterpret_cast<ulonglong2*>(keys + i * 2) = make_ulonglong2(b1, b2);
}

__global__ void __launch_bounds__(256)
hamming_gather_kernel(const uint8_t* __restrict__ t, long long nt, unsigned long long train_base,
                      const unsigned long long* __restrict__ keys, long long nq,
                      uint8_t* __restrict__ rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long k = keys[2 * i];
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (k != PBK_KEY_NONE) {
    const unsigned long long g = k & 0xffffffffull;
    if (g >= train_base && g - train_base < (unsigned long long)nt) {
      const uint4* src = reinterpret_cast<const uint4*>(t + (g - train_base) * 32);
      a = src[0];
      b = src[1];
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(rows + i * 32);
  dst[0] = a;
  dst[1] = b;
}

__global__ void __launch_bounds__(256)
hamming_ratio_mutual_kernel(const unsigned long long* __restrict__ keys,
                            const unsigned long long* __restrict__ rev, long long nq, double ratio,
                            long long* __restrict__ idx, int* __restrict__ dist,
                            uint8_t* __restrict__ ratio_ok, uint8_t* __restrict__ mutual_ok,
                            uint8_t* __restrict__ valid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const unsigned long long a = keys[2 * i], b = keys[2 * i + 1];
  const bool h1 = a != PBK_KEY_NONE, h2 = b != PBK_KEY_NONE;
  const int d1 = h1 ? (int)(a >> 32) : -1, d2 = h2 ? (int)(b >> 32) : -1;
  if (idx != nullptr) {
    idx[2 * i] = h1 ? (long long)(a & 0xffffffffull) : -1;
    idx[2 * i + 1] = h2 ? (long long)(b & 0xffffffffull) : -1;
  }
  if (dist != nullptr) {
    dist[2 * i] = d1;
    dist[2 * i + 1] = d2;
  }
  // Python: m.distance < ratio * n.distance on doubles holding exact integers
  const bool r_ok = h2 && ((double)d1 < __dmul_rn(ratio, (double)d2));
  bool m_ok = false;
  if (h1 && rev != nullptr) {
    const unsigned long long rk = rev[2 * i];
    m_ok = rk != PBK_KEY_NONE && (long long)(rk & 0xffffffffull) == i;
  }
  if (ratio_ok != nullptr) ratio_ok[i] = r_ok;
  if (mutual_ok != nullptr) mutual_ok[i] = m_ok;
  if (valid != nullptr) valid[i] = r_ok && m_ok;
}

// ------------------------------------------ multi-GPU exchange over peer memory
// The sharded matcher needs two exchanges per query frame: every rank's local
// top-2 keys (nq x 16 B) to all ranks, and the matched train rows (32 B each,
// owned by exactly one rank) to all ranks for the mutual check.  Both are fused
// into the kernels that produce the data: they store straight into every peer's
// exchange buffer over NVLink (peer-mapped symmetric memory) and publish a
// sequence number; the consuming kernel waits for the `world` sequence numbers.
// Exchange buffer layout (u64 words), two copies selected by the call parity:
//   keys  [2][world][nq][2]   local top-2 keys of every rank
//   rows  [2][nq][4]          matched 256-bit train rows
//   flags [2][2][world]       sequence number per (phase, producing rank)
// A rank reuses a copy only after every peer has finished the call in between.
struct HammingPeers {
  unsigned long long* buf[kHMaxWorld];
  long long nq;
  int rank, world;
  unsigned long long seq;
  long long timeout_clk;        // bound of the spin waits (SM clocks)
  uint32_t* status;             // set to kStatusPeerTimeout when a wait ran into the bound
};
__device__ __forceinline__ size_t hx_keys(const HammingPeers& X, int par, int r) {
  return ((size_t)par * X.world + r) * (size_t)X.nq * 2;
}
__device__ __forceinline__ size_t hx_rows(const HammingPeers& X, int par) {
  return (size_t)2 * X.world * X.nq * 2 + (size_t)par * X.nq * 4;
}
__device__ __forceinline__ size_t hx_flags(const HammingPeers& X, int par, int phase) {
  return (size_t)2 * X.world * X.nq * 2 + (size_t)2 * X.nq * 4 + ((size_t)par * 2 + phase) * X.world;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// the last CTA of a producing kernel publishes the sequence number to every peer
__device__ __forceinline__ void hx_publish(const HammingPeers& X, int phase, unsigned int* done_counter) {
  __shared__ int s_last;
  __threadfence_system();                      // this thread's peer stores are visible system-wide
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < X.world) {
      unsigned long long* f = X.buf[threadIdx.x] + hx_flags(X, (int)(X.seq & 1ull), phase) + X.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(X.seq) : "memory");
    }
    if (threadIdx.x == 0) *done_counter = 0u;  // ready for the next call
  }
}
// every CTA of a consuming kernel waits until all ranks have published `phase`.  The wait is
// bounded (X.timeout_clk, ~2 s by default): a peer that never arrives makes the call FAIL -
// *X.status is set (the host wrapper raises on it) and the caller poisons its outputs
// (no-match keys / zero rows) instead of consuming stale exchange words.  Returns true when
// every rank's data has arrived.
__device__ __forceinline__ bool hx_wait(const HammingPeers& X, int phase) {
  bool ok = true;
  if ((int)threadIdx.x < X.world) {
    const unsigned long long* f = X.buf[X.rank] + hx_flags(X, (int)(X.seq & 1ull), phase) + threadIdx.x;
    unsigned long long got = 0;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(f) : "memory");
    } while (got != X.seq && clock64() - t0 < X.timeout_clk);
    ok = got == X.seq;
    if (!ok) raise_status(X.status, kStatusPeerTimeout);
  }
  return __syncthreads_and(ok) != 0;
}

// split merge of this rank's partial keys + push of the result to every peer
__global__ void __launch_bounds__(256)
hamming_merge_push_kernel(const unsigned long long* __restrict__ partial, int parts,
                          const __grid_constant__ HammingPeers X, unsigned int* done_counter) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / kMergeLanes;
  const int sub = (int)(t % kMergeLanes);
  unsigned long long b1, b2;
  merge_parts(partial, parts, X.nq, i, sub, b1, b2);
  if (i < X.nq) {
    // the query's lanes share the pushes: lane `sub` serves the peers sub, sub + 8, ...
    const size_t off = hx_keys(X, (int)(X.seq & 1ull), X.rank) + (size_t)i * 2;
    for (int p = sub; p < X.world; p += kMergeLanes) {
      st_sys_u64(X.buf[p] + off, b1);
      st_sys_u64(X.buf[p] + off + 1, b2);
    }
  }
  hx_publish(X, 0, done_counter);
}

// wait for every rank's keys, then global top-2 by unsigned min (same order on every rank)
__global__ void __launch_bounds__(256)
hamming_merge_world_kernel(const __grid_constant__ HammingPeers X, unsigned long long* __restrict__ keys) {
  const bool ok = hx_wait(X, 0);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= X.nq) return;
  const unsigned long long* mine = X.buf[X.rank];
  unsigned long long b1 = PBK_KEY_NONE, b2 = PBK_KEY_NONE;
  for (int r = 0; ok && r < X.world; ++r) {
    const unsigned long long* src = mine + hx_keys(X, (int)(X.seq & 1ull), r) + (size_t)i * 2;
    const unsigned long long vx = __ldcv(src), vy = __ldcv(src + 1);
    unsigned long long mx = max(b1, vx);
    b1 = min(b1, vx);
    b2 = min(b2, mx);
    mx = max(b1, vy);
    b1 = min(b1, vy);
    b2 = min(b2, mx);
  }
  *reinterpret_cast<ulonglong2*>(keys + i * 2) = make_ulonglong2(b1, b2);
}

// the owner of each matched train row pushes it to every peer (rank 0 pushes zeros for "no match")
__global__ void __launch_bounds__(256)
hamming_gather_push_kernel(const uint8_t* __restrict__ t, long long nt, unsigned long long train_base,
                           const unsigned long long* __restrict__ keys, const __grid_constant__ HammingPeers X,
                           unsigned int* done_counter) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < X.nq) {
    const unsigned long long k = keys[2 * i];
    const unsigned long long g = k & 0xffffffffull;
    const bool none = k == PBK_KEY_NONE;
    const bool own = !none && g >= train_base && g - train_base < (unsigned long long)nt;
    if (own || (none && X.rank == 0)) {
      ulonglong2 a = make_ulonglong2(0, 0), b = a;
      if (own) {
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(t + (g - train_base) * 32);
        a = src[0];
        b = src[1];
      }
      const size_t off = hx_rows(X, (int)(X.seq & 1ull)) + (size_t)i * 4;
      for (int p = 0; p < X.world; ++p) {
        unsigned long long* dst = X.buf[p] + off;
        st_sys_u64(dst, a.x); st_sys_u64(dst + 1, a.y); st_sys_u64(dst + 2, b.x); st_sys_u64(dst + 3, b.y);
      }
    }
  }
  hx_publish(X, 1, done_counter);
}

// wait for every rank's rows, then copy the matched rows out of the exchange buffer
__global__ void __launch_bounds__(256)
hamming_rows_world_kernel(const __grid_constant__ HammingPeers X, uint8_t* __restrict__ rows) {
  const bool ok = hx_wait(X, 1);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= X.nq) return;
  const unsigned long long* src = X.buf[X.rank] + hx_rows(X, (int)(X.seq & 1ull)) + (size_t)i * 4;
  ulonglong2* dst = reinterpret_cast<ulonglong2*>(rows + i * 32);
  dst[0] = ok ? make_ulonglong2(__ldcv(src), __ldcv(src + 1)) : make_ulonglong2(0, 0);
  dst[1] = ok ? make_ulonglong2(__ldcv(src + 2), __ldcv(src + 3)) : make_ulonglong2(0, 0);
}

// ------------------------------------------------------- popc pipe probe
// mode 0: pure POPC chains (8 independent per thread) -> popc-pipe peak
// mode 1: POPC + one LOP3 per POPC (checks POPC/ALU overlap)
// mode 2 / 3: the matcher's own inner loop (VARIANT 0 / 1) over one resident
//             shared-memory tile, no global traffic, no barriers: the
//             instruction-mix ceiling of the kernel on this chip.
template <int MODE>
__global__ void __launch_bounds__(MODE >= 2 ? kHThreads : 256) popc_probe_kernel(int iters, uint32_t* sink) {
  if (MODE >= 2) {
    __shared__ __align__(16) uint4 tile[kTR * 2];
    for (int i = threadIdx.x; i < kTR * 2; i += blockDim.x) {
      const uint32_t h = (i + 1) * 2654435761u;
      tile[i] = make_uint4(h, h * 40503u, h ^ (h >> 7), h * 0x9e3779b9u);
    }
    __syncthreads();
    uint32_t qr[kRQ][8], k1[kRQ], k2[kRQ];
#pragma unroll
    for (int r = 0; r < kRQ; ++r) {
      k1[r] = k2[r] = kLocalNone;
#pragma unroll
      for (int w = 0; w < 8; ++w) qr[r][w] = (threadIdx.x * 8 + w + r * 977) * 2246822519u + sink[w];
    }
    for (int it = 0; it < iters; ++it)
      score_tile<MODE - 2, true>(tile, kTR, (uint32_t)it * kTR, qr, k1, k2);
    uint32_t acc = 0;
#pragma unroll
    for (int r = 0; r < kRQ; ++r) acc ^= k1[r] ^ k2[r];
    sink[8 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
    return;
  }
  uint32_t v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __popc(v[i]);
      }
    }
  } else {
    const uint32_t c = sink[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __popc(v[i] ^ c);
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= v[i];
  sink[8 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// hamming_tc.cu: expands q into q_signs, scans t_signs, writes partial keys (*parts, nq, 2)
int hamming_tc_scan(const uint8_t* q, long long nq, const int8_t* t_signs, long long nt, unsigned long long train_base,
                    int8_t* q_signs, unsigned long long* partial, int* parts, const DeviceInfo* di, cudaStream_t s);

static std::atomic<int> g_variant{-1};
static int hamming_variant() {
  int v = g_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("PBK_HAMMING_VARIANT");
    v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
    g_variant.store(v, std::memory_order_relaxed);      // idempotent: every thread computes the same value
  }
  return v;
}

static int plan(const DeviceInfo* di, long long nq, long long nt, int* splits, long long* rows_per_split,
                int* rq = nullptr) {
  const long long target = (long long)di->sm_count * 4;            // 4 CTAs (16 warps) per SM
  const long long stages = (nt + kTR - 1) / kTR;
  // queries per thread: 4, unless even one-stage splits leave SMs without a CTA
  const long long qt4 = (nq + kQTile - 1) / kQTile;
  const int rqv = (qt4 * (stages > 0 ? stages : 1) < di->sm_count) ? kRQSmall : kRQ;
  if (rq) *rq = rqv;
  const long long qtile = (long long)kHThreads * rqv;
  const long long qtiles = (nq + qtile - 1) / qtile;
  long long s = qtiles > 0 ? (target + qtiles - 1) / qtiles : 1;
  const long long max_by_rows = (nt + kTR - 1) / kTR;              // >= one stage per split
  if (s > max_by_rows) s = max_by_rows;
  if (s < 1) s = 1;
  long long rps = (nt + s - 1) / s;
  rps = ((rps + kTR - 1) / kTR) * kTR;
  if (rps > kMaxSplitRows) rps = kMaxSplitRows;
  if (rps < kTR) rps = kTR;
  s = (nt + rps - 1) / rps;
  if (s < 1) s = 1;
  PBK_REQUIRE(s <= 65535, PBK_EINVAL, "train set too large for one call (%lld splits)", s);
  *splits = (int)s;
  *rows_per_split = rps;
  return PBK_OK;
}

}  // namespace pbk

namespace pbk {
// launches the scoring kernel in the shape plan() chose (shared-memory opt-in cached per device)
template <int V, int R>
static int launch_knn2_as(const HammingParams& P, const DeviceInfo* di, cudaStream_t s) {
  constexpr int qtile = kHThreads * R;
  constexpr int smem = kStages * kStageBytes + qtile * 32 + (kStages + 1) * 8;
  static KernelCache kc;
  const int rc = kc.setup(di, hamming_knn2_kernel<V, R>, kHThreads, smem, nullptr);
  if (rc) return rc;
  const dim3 grid((unsigned)((P.nq + qtile - 1) / qtile), (unsigned)P.splits);
  hamming_knn2_kernel<V, R><<<grid, kHThreads, smem, s>>>(P);
  PBK_CHECK_LAUNCH("hamming_knn2_kernel");
  return PBK_OK;
}
static int launch_knn2(const HammingParams& P, int rq, const DeviceInfo* di, cudaStream_t s) {
  const int v = hamming_variant();
  if (v == 0) return rq == kRQ ? launch_knn2_as<0, kRQ>(P, di, s) : launch_knn2_as<0, kRQSmall>(P, di, s);
  if (v == 2 && rq == kRQ) return launch_knn2_as<2, kRQ>(P, di, s);
  return rq == kRQ ? launch_knn2_as<1, kRQ>(P, di, s) : launch_knn2_as<1, kRQSmall>(P, di, s);
}
}  // namespace pbk

extern "C" {

int pbk_hamming_plan(int64_t nq, int64_t nt, int* splits, size_t* scratch_bytes) {
  using namespace pbk;
  PBK_REQUIRE(nq >= 0 && nt >= 0 && splits && scratch_bytes, PBK_EINVAL, "bad arguments");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  long long rps;
  rc = plan(di, nq, nt, splits, &rps);
  if (rc) return rc;
  *scratch_bytes = (size_t)(*splits) * (size_t)(nq > 0 ? nq : 1) * 16;
  return PBK_OK;
}

int pbk_hamming_knn2(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, uint64_t train_base,
                     uint64_t* keys, void* scratch, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nq >= 0 && nt >= 0, PBK_EINVAL, "negative size");
  if (nq == 0) return PBK_OK;
  PBK_REQUIRE(q != nullptr && keys != nullptr && scratch != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(nt == 0 || t != nullptr, PBK_EINVAL, "NULL train pointer");
  PBK_REQUIRE(aligned16(q) && aligned16(t) && aligned16(keys) && aligned16(scratch), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  PBK_REQUIRE(train_base + (uint64_t)nt <= (1ull << 32), PBK_EINVAL,
              "global train row index must fit 32 bits");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  int splits, rq;
  long long rps;
  rc = plan(di, nq, nt, &splits, &rps, &rq);
  if (rc) return rc;
  HammingParams P;
  P.q = q; P.t = t; P.partial = static_cast<unsigned long long*>(scratch);
  P.nq = nq; P.nt = nt; P.rows_per_split = rps; P.train_base = train_base; P.splits = splits;
  rc = launch_knn2(P, rq, di, s);
  if (rc) return rc;
  hamming_merge_kernel<<<merge_blocks(nq), 256, 0, s>>>(P.partial, splits, nq,
                                                                    reinterpret_cast<unsigned long long*>(keys));
  PBK_CHECK_LAUNCH("hamming_merge_kernel");
  return PBK_OK;
}

int pbk_hamming_merge_top2(const uint64_t* partial, int parts, int64_t nq, uint64_t* keys,
                           pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(parts >= 1 && nq >= 0, PBK_EINVAL, "bad arguments");
  if (nq == 0) return PBK_OK;
  PBK_REQUIRE(partial != nullptr && keys != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(partial) && aligned16(keys), PBK_EALIGN, "device pointers must be 16-byte aligned");
  hamming_merge_kernel<<<merge_blocks(nq), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const unsigned long long*>(partial), parts, nq,
      reinterpret_cast<unsigned long long*>(keys));
  PBK_CHECK_LAUNCH("hamming_merge_kernel");
  return PBK_OK;
}

int pbk_hamming_gather_best(const uint8_t* t, int64_t nt, uint64_t train_base, const uint64_t* keys,
                            int64_t nq, uint8_t* rows, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nt >= 0 && nq >= 0, PBK_EINVAL, "negative size");
  if (nq == 0) return PBK_OK;
  PBK_REQUIRE(keys != nullptr && rows != nullptr && (nt == 0 || t != nullptr), PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(t) && aligned16(keys) && aligned16(rows), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  hamming_gather_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, as_stream(stream)>>>(
      t, nt, train_base, reinterpret_cast<const unsigned long long*>(keys), nq, rows);
  PBK_CHECK_LAUNCH("hamming_gather_kernel");
  return PBK_OK;
}

int pbk_hamming_ratio_mutual(const uint64_t* keys, const uint64_t* rev_keys, int64_t nq, double ratio,
                             int64_t* idx, int32_t* dist, uint8_t* ratio_ok, uint8_t* mutual_ok,
                             uint8_t* valid, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nq >= 0, PBK_EINVAL, "negative size");
  if (nq == 0) return PBK_OK;
  PBK_REQUIRE(keys != nullptr, PBK_EINVAL, "NULL device pointer");
  hamming_ratio_mutual_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const unsigned long long*>(keys), reinterpret_cast<const unsigned long long*>(rev_keys),
      nq, ratio, reinterpret_cast<long long*>(idx), dist, ratio_ok, mutual_ok, valid);
  PBK_CHECK_LAUNCH("hamming_ratio_mutual_kernel");
  return PBK_OK;
}

size_t pbk_hamming_exchange_words(int64_t nq, int world) {
  if (nq < 0 || world < 1) return 0;
  return (size_t)2 * world * nq * 2 + (size_t)2 * nq * 4 + (size_t)4 * world;
}

static int fill_peers(pbk::HammingPeers* X, const uint64_t* peer_bufs, int64_t nq, int rank, int world,
                      uint64_t seq, uint32_t* status) {
  using namespace pbk;
  PBK_REQUIRE(world >= 2 && world <= kHMaxWorld && rank >= 0 && rank < world, PBK_EINVAL, "bad rank/world");
  PBK_REQUIRE(peer_bufs != nullptr && seq >= 1, PBK_EINVAL, "the fused exchange needs peer buffers and seq >= 1");
  for (int p = 0; p < kHMaxWorld; ++p)
    X->buf[p] = p < world ? reinterpret_cast<unsigned long long*>(peer_bufs[p]) : nullptr;
  X->nq = nq; X->rank = rank; X->world = world; X->seq = seq;
  X->timeout_clk = exchange_timeout_clk();
  X->status = status;
  return PBK_OK;
}

int pbk_hamming_knn2_dist(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, uint64_t train_base,
                          uint64_t* keys, void* scratch, const uint64_t* peer_bufs, int rank, int world,
                          uint64_t seq, uint32_t* done_counter, uint32_t* status, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nq > 0 && nt >= 0, PBK_EINVAL, "bad sizes");
  PBK_REQUIRE(q != nullptr && keys != nullptr && scratch != nullptr && done_counter != nullptr, PBK_EINVAL,
              "NULL device pointer");
  PBK_REQUIRE(nt == 0 || t != nullptr, PBK_EINVAL, "NULL train pointer");
  PBK_REQUIRE(aligned16(q) && aligned16(t) && aligned16(keys) && aligned16(scratch), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  PBK_REQUIRE(train_base + (uint64_t)nt <= (1ull << 32), PBK_EINVAL, "global train row index must fit 32 bits");
  HammingPeers X;
  int rc = fill_peers(&X, peer_bufs, nq, rank, world, seq, status);
  if (rc) return rc;
  const DeviceInfo* di;
  rc = current_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  int splits, rq;
  long long rps;
  rc = plan(di, nq, nt, &splits, &rps, &rq);
  if (rc) return rc;
  HammingParams P;
  P.q = q; P.t = t; P.partial = static_cast<unsigned long long*>(scratch);
  P.nq = nq; P.nt = nt; P.rows_per_split = rps; P.train_base = train_base; P.splits = splits;
  rc = launch_knn2(P, rq, di, s);
  if (rc) return rc;
  const unsigned blocks = (unsigned)((nq + 255) / 256);
  hamming_merge_push_kernel<<<merge_blocks(nq), 256, 0, s>>>(P.partial, splits, X, done_counter);
  PBK_CHECK_LAUNCH("hamming_merge_push_kernel");
  hamming_merge_world_kernel<<<blocks, 256, 0, s>>>(X, reinterpret_cast<unsigned long long*>(keys));
  PBK_CHECK_LAUNCH("hamming_merge_world_kernel");
  return PBK_OK;
}

int pbk_hamming_gather_best_dist(const uint8_t* t, int64_t nt, uint64_t train_base, const uint64_t* keys,
                                 int64_t nq, uint8_t* rows, const uint64_t* peer_bufs, int rank, int world,
                                 uint64_t seq, uint32_t* done_counter, uint32_t* status, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nt >= 0 && nq > 0, PBK_EINVAL, "bad sizes");
  PBK_REQUIRE(keys != nullptr && rows != nullptr && done_counter != nullptr && (nt == 0 || t != nullptr),
              PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(t) && aligned16(keys) && aligned16(rows), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  HammingPeers X;
  int rc = fill_peers(&X, peer_bufs, nq, rank, world, seq, status);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const unsigned blocks = (unsigned)((nq + 255) / 256);
  hamming_gather_push_kernel<<<blocks, 256, 0, s>>>(t, nt, train_base,
                                                    reinterpret_cast<const unsigned long long*>(keys), X, done_counter);
  PBK_CHECK_LAUNCH("hamming_gather_push_kernel");
  hamming_rows_world_kernel<<<blocks, 256, 0, s>>>(X, rows);
  PBK_CHECK_LAUNCH("hamming_rows_world_kernel");
  return PBK_OK;
}

// ---------------------------------------------------------------- tensor-core route
// (hamming_tc.cu: the scan; the split / world merges are the kernels above)
int pbk_hamming_tc_knn2(const uint8_t* q, int64_t nq, const int8_t* t_signs, int64_t nt, uint64_t train_base,
                        uint64_t* keys, void* scratch, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nq >= 0 && nt >= 0, PBK_EINVAL, "negative size");
  if (nq == 0) return PBK_OK;
  PBK_REQUIRE(q != nullptr && keys != nullptr && scratch != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(nt == 0 || t_signs != nullptr, PBK_EINVAL, "NULL train pointer");
  PBK_REQUIRE(aligned16(q) && aligned16(t_signs) && aligned16(keys) && aligned16(scratch), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  PBK_REQUIRE(train_base + (uint64_t)nt <= (1ull << 32), PBK_EINVAL, "global train row index must fit 32 bits");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  int parts = 0;
  size_t bytes = 0;
  rc = pbk_hamming_tc_plan(nq, nt, &parts, &bytes);
  if (rc) return rc;
  unsigned long long* partial = static_cast<unsigned long long*>(scratch);
  int8_t* q_signs = static_cast<int8_t*>(scratch) + (((size_t)parts * (size_t)nq * 16 + 255) & ~(size_t)255);
  rc = hamming_tc_scan(q, nq, t_signs, nt, train_base, q_signs, partial, &parts, di, s);
  if (rc) return rc;
  hamming_merge_kernel<<<merge_blocks(nq), 256, 0, s>>>(partial, parts, nq,
                                                                    reinterpret_cast<unsigned long long*>(keys));
  PBK_CHECK_LAUNCH("hamming_merge_kernel");
  return PBK_OK;
}

int pbk_hamming_tc_knn2_dist(const uint8_t* q, int64_t nq, const int8_t* t_signs, int64_t nt, uint64_t train_base,
                             uint64_t* keys, void* scratch, const uint64_t* peer_bufs, int rank, int world,
                             uint64_t seq, uint32_t* done_counter, uint32_t* status, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nq > 0 && nt >= 0, PBK_EINVAL, "bad sizes");
  PBK_REQUIRE(q != nullptr && keys != nullptr && scratch != nullptr && done_counter != nullptr, PBK_EINVAL,
              "NULL device pointer");
  PBK_REQUIRE(nt == 0 || t_signs != nullptr, PBK_EINVAL, "NULL train pointer");
  PBK_REQUIRE(aligned16(q) && aligned16(t_signs) && aligned16(keys) && aligned16(scratch), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  PBK_REQUIRE(train_base + (uint64_t)nt <= (1ull << 32), PBK_EINVAL, "global train row index must fit 32 bits");
  HammingPeers X;
  int rc = fill_peers(&X, peer_bufs, nq, rank, world, seq, status);
  if (rc) return rc;
  const DeviceInfo* di;
  rc = current_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  int parts = 0;
  size_t bytes = 0;
  rc = pbk_hamming_tc_plan(nq, nt, &parts, &bytes);
  if (rc) return rc;
  unsigned long long* partial = static_cast<unsigned long long*>(scratch);
  int8_t* q_signs = static_cast<int8_t*>(scratch) + (((size_t)parts * (size_t)nq * 16 + 255) & ~(size_t)255);
  rc = hamming_tc_scan(q, nq, t_signs, nt, train_base, q_signs, partial, &parts, di, s);
  if (rc) return rc;
  const unsigned blocks = (unsigned)((nq + 255) / 256);
  hamming_merge_push_kernel<<<merge_blocks(nq), 256, 0, s>>>(partial, parts, X, done_counter);
  PBK_CHECK_LAUNCH("hamming_merge_push_kernel");
  hamming_merge_world_kernel<<<blocks, 256, 0, s>>>(X, reinterpret_cast<unsigned long long*>(keys));
  PBK_CHECK_LAUNCH("hamming_merge_world_kernel");
  return PBK_OK;
}

int pbk_popc_peak_probe(int mode, int ctas, int iters, uint32_t* sink, double* popc_per_launch,
                        pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(ctas > 0 && iters > 0 && sink != nullptr, PBK_EINVAL, "bad arguments");
  cudaStream_t s = as_stream(stream);
  double per_thread;   // POPC per thread per iteration (modes 2/3: 4 queries x 256 rows x POPC/pair)
  int threads = 256;
  switch (mode) {
    case 0: popc_probe_kernel<0><<<ctas, 256, 0, s>>>(iters, sink); per_thread = 16.0 * 8; break;
    case 1: popc_probe_kernel<1><<<ctas, 256, 0, s>>>(iters, sink); per_thread = 16.0 * 8; break;
    case 2: popc_probe_kernel<2><<<ctas, kHThreads, 0, s>>>(iters, sink); per_thread = kRQ * kTR * 8.0; threads = kHThreads; break;
    case 3: popc_probe_kernel<3><<<ctas, kHThreads, 0, s>>>(iters, sink); per_thread = kRQ * kTR * 4.0; threads = kHThreads; break;
    case 4: popc_probe_kernel<4><<<ctas, kHThreads, 0, s>>>(iters, sink); per_thread = kRQ * kTR * 4.0; threads = kHThreads; break;
    case 5: popc_probe_kernel<5><<<ctas, kHThreads, 0, s>>>(iters, sink); per_thread = kRQ * kTR * 4.0; threads = kHThreads; break;
    default: PBK_REQUIRE(false, PBK_EINVAL, "bad probe mode %d", mode);
  }
  PBK_CHECK_LAUNCH("popc_probe_kernel");
  if (popc_per_launch) *popc_per_launch = per_thread * (double)iters * (double)threads * (double)ctas;
  return PBK_OK;
}

}  // extern "C"
