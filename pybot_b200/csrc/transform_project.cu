// Batched SE(3)/Sim(3) point transform and Camera.project (SURVEY.md 8a rows
// a1, a3, a4, a5).
//
// Replaces
//   np.hstack + np.dot(self.matrix, X)      pybot/geometry/rigid_transform.py:139-141
//   cv2.projectPoints(X, rvec, t, K, D)     pybot/vision/camera_utils.py:697
//   depth_from_projection                   camera_utils.py:736-743
//   depth / bounds masks and x[valid]       camera_utils.py:708-732
//
// Both kernels stream AoS (n,3) points.  A warp owns a tile of 128 points: the
// tile is loaded with coalesced 16 B loads into a warp-private shared-memory
// stage, lane l then owns points {l, l+32, l+64, l+96} of the tile (3-word
// stride: conflict-free), computes in fp64 registers, and results leave
// through the same stage as coalesced 16 B stores (transform) or directly as
// lane-contiguous 8/16 B stores (projection).  Algorithmic traffic: 12 B in +
// 8 B out per point for a float32 projection.
//
// Arithmetic
//   transform: acc = R_i0*X; acc = fma(R_i1,Y,acc); acc = fma(R_i2,Z,acc);
//              acc = fma(t_i,1,acc)   - the k=0..3 FMA chain OpenBLAS's dgemm
//              runs for np.dot(T4x4, X4xN) (bit-exact with numpy 2.3.5 here).
//   project:   OpenCV's projectPoints loop, fp64, no FMA (oracle/model.py),
//              one rounding to the output dtype.
//
// Order-preserving compaction (x[valid]) is fused: tiles of 1024 points chain
// their valid counts with a decoupled look-back (project_compact_kernel), so
// the compacted outputs are written in the same pass that reads X.
#include <type_traits>

#include "pbk_common.cuh"
#include "project_math.cuh"

namespace pbk {

constexpr int kWarps = 8;
constexpr int kTilePts = 128;                 // per warp
constexpr int kCtaPts = kWarps * kTilePts;    // 1024

struct XformParams {
  double R[9];
  double t[3];
  double scale;
  int has_scale;
};

// ragged tile: plain loads of `valid_bytes` (multiple of 4), zero fill to `bytes`
__device__ __forceinline__ void warp_load_ragged(const char* __restrict__ g, char* s, int lane,
                                                 int bytes, int valid_bytes) {
  for (int off = lane * 4; off < bytes; off += 128)
    *reinterpret_cast<uint32_t*>(s + off) = off < valid_bytes ? ld_stream_u1(g + off) : 0u;
}

template <typename T>
__device__ __forceinline__ void stage_read_point(const char* s, int p, double X[3]) {
  const T* r = reinterpret_cast<const T*>(s) + 3 * p;
  X[0] = (double)r[0]; X[1] = (double)r[1]; X[2] = (double)r[2];
}

template <typename T> struct RingCfg { static constexpr int kStages = 4; };
template <> struct RingCfg<double> { static constexpr int kStages = 3; };

// =============================================================== transform
template <typename TI, typename TO>
__global__ void __launch_bounds__(kWarps * 32)
transform_points_kernel(const TI* __restrict__ X, long long n, TO* __restrict__ Y,
                        const __grid_constant__ XformParams P) {
  constexpr int NST = RingCfg<TI>::kStages;
  constexpr int kInBytes = kTilePts * 3 * sizeof(TI);
  constexpr int kOutBytes = kTilePts * 3 * sizeof(TO);
  extern __shared__ __align__(128) char dyn_smem[];
  // layout: [warp][NST][kInBytes] | [warp][kOutBytes] | barriers
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* out_stage = dyn_smem + kWarps * NST * kInBytes + warp * kOutBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + kWarps * (NST * kInBytes + kOutBytes)) + warp * NST;
  WarpRing<kInBytes, NST> ring;
  ring.init(dyn_smem + warp * NST * kInBytes, bars, lane);

  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  const long long nfull = n / kTilePts;
  const long long stride = (long long)gridDim.x * kWarps;
  const long long t0 = (long long)blockIdx.x * kWarps + warp;
  const char* Xb = reinterpret_cast<const char*>(X);
  if (lane == 0)
    for (int k = 0; k < NST - 1; ++k)
      if (t0 + k * stride < nfull) ring.issue(k, Xb + (t0 + k * stride) * kInBytes);

  int k = 0;
  for (long long tile = t0; tile < ntiles; tile += stride, ++k) {
    const long long nxt = tile + (NST - 1) * stride;
    if (lane == 0 && nxt < nfull) ring.issue(k + NST - 1, Xb + nxt * kInBytes);
    const long long first = tile * kTilePts;
    const int npts = (int)((n - first) < kTilePts ? (n - first) : kTilePts);
    char* st = ring.slot(k);
    if (tile < nfull) ring.wait(k);
    else { warp_load_ragged(Xb + first * 3 * sizeof(TI), st, lane, kInBytes, npts * 3 * (int)sizeof(TI)); __syncwarp(); }
    double Xp[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) stage_read_point<TI>(st, j * 32 + lane, Xp[j]);
    __syncwarp();                       // slot may be refilled from here on
    TO* so = reinterpret_cast<TO*>(out_stage);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double acc = __dmul_rn(P.R[3 * i], Xp[j][0]);
        acc = __fma_rn(P.R[3 * i + 1], Xp[j][1], acc);
        acc = __fma_rn(P.R[3 * i + 2], Xp[j][2], acc);
        if (P.has_scale) acc = __fma_rn(P.scale, acc, P.t[i]);
        else acc = __fma_rn(P.t[i], 1.0, acc);
        so[3 * (j * 32 + lane) + i] = (TO)acc;
      }
    }
    __syncwarp();
    char* dst = reinterpret_cast<char*>(Y + first * 3);
    const int valid = npts * 3 * (int)sizeof(TO);
    if (valid == kOutBytes) {
#pragma unroll
      for (int off = 0; off < kOutBytes; off += 512)
        st_stream_u4(dst + off + lane * 16, *reinterpret_cast<const uint4*>(out_stage + off + lane * 16));
    } else {
      for (int off = lane * 4; off < valid; off += 128)
        *reinterpret_cast<uint32_t*>(dst + off) = *reinterpret_cast<const uint32_t*>(out_stage + off);
    }
    __syncwarp();
  }
}

// ================================================================= project
struct Projected {
  double u, v, depth;
  bool valid;
};

template <typename T> struct OutPix;
template <> struct OutPix<float> {
  using vec = float2;
  static __device__ __forceinline__ float2 make(float u, float v) { return make_float2(u, v); }
  static __device__ __forceinline__ void store_vec(float* uv, long long i, float2 p) { st_stream_f2(uv + 2 * i, p); }
  static __device__ __forceinline__ bool in_bounds(float u, float v, float W, float H) {
    return u >= 0.f && u < W && v >= 0.f && v < H;
  }
  static __device__ __forceinline__ void store(float* uv, long long i, float u, float v) {
    st_stream_f2(uv + 2 * i, make_float2(u, v));
  }
  // predicated (branch-free) store of one pixel at p
  static __device__ __forceinline__ void store_if(unsigned pred, float* p, float u, float v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.cs.v2.f32 [%1], {%2,%3};\n\t}"
                 ::"r"(pred), "l"(p), "f"(u), "f"(v) : "memory");
  }
};
template <> struct OutPix<double> {
  using vec = double2;
  static __device__ __forceinline__ double2 make(double u, double v) { return make_double2(u, v); }
  static __device__ __forceinline__ void store_vec(double* uv, long long i, double2 p) { store(uv, i, p.x, p.y); }
  static __device__ __forceinline__ bool in_bounds(double u, double v, double W, double H) {
    return u >= 0.0 && u < W && v >= 0.0 && v < H;
  }
  static __device__ __forceinline__ void store(double* uv, long long i, double u, double v) {
    double2 o = make_double2(u, v);
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(uv + 2 * i), "d"(o.x), "d"(o.y) : "memory");
  }
  static __device__ __forceinline__ void store_if(unsigned pred, double* p, double u, double v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.cs.v2.f64 [%1], {%2,%3};\n\t}"
                 ::"r"(pred), "l"(p), "d"(u), "d"(v) : "memory");
  }
};
__device__ __forceinline__ void st_stream_d1_if(unsigned pred, double* p, double v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.cs.f64 [%1], %2;\n\t}"
               ::"r"(pred), "l"(p), "d"(v) : "memory");
}

// status word of the decoupled look-back: [63:62] state, [61:0] value
constexpr unsigned long long kStateAgg = 1ull << 62;
constexpr unsigned long long kStatePrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// sum of the published 64-bit status words st[0 .. n) (values only): the eight loads
// of a lane are issued before the first one is examined
__device__ __forceinline__ long long sum_aggregates(const unsigned long long* st, long long n, int lane) {
  long long acc = 0;
  for (long long c0 = 0; c0 < n; c0 += 256) {
    unsigned long long sw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long idx = c0 + lane + 32 * j;
      sw[j] = idx < n ? ld_status(st + idx) : kStateAgg;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long idx = c0 + lane + 32 * j;
      while ((sw[j] >> 62) == 0) sw[j] = ld_status(st + idx);
      acc += (long long)(sw[j] & kValueMask);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// Valid flags of four 32-point rows (ballots m[0..3], row r = points base + 32 r ..)
// as 128 bytes of 0/1: lane l writes the four flags (l % 8) * 4 .. + 3 of row l / 8 as
// one 32-bit word - one coalesced 128 B store per warp instead of four rows of
// single-byte stores.  `row_stride` = distance between the rows in points.
__device__ __forceinline__ void store_valid_rows(uint8_t* dst, const unsigned (&m)[4], int row_stride, int lane) {
  const int r = lane >> 3;
  const unsigned mr = r == 0 ? m[0] : (r == 1 ? m[1] : (r == 2 ? m[2] : m[3]));
  const unsigned nib = (mr >> ((lane & 7) * 4)) & 0xFu;
  const unsigned word = (nib * 0x00204081u) & 0x01010101u;      // bit i -> byte i
  *reinterpret_cast<unsigned*>(dst + (size_t)r * row_stride + (lane & 7) * 4) = word;
}

// MODE 0: plain projection (no depth, masks, count).
// MODE 1: depth / masks / valid flags per runtime switches, outputs uncompacted.
// (the compacting variant is project_compact_kernel below)
template <typename T, bool DIST, int MODE>
__global__ void __launch_bounds__(kWarps * 32)
project_points_kernel(const T* __restrict__ X, long long n, T* __restrict__ uv,
                      double* __restrict__ depth, uint8_t* __restrict__ valid_out,
                      long long* __restrict__ count, const __grid_constant__ pbk_project_params P) {
  constexpr int NST = RingCfg<T>::kStages;
  constexpr int kInBytes = kTilePts * 3 * sizeof(T);
  extern __shared__ __align__(128) char dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + kWarps * NST * kInBytes) + warp * NST;
  WarpRing<kInBytes, NST> ring;
  ring.init(dyn_smem + warp * NST * kInBytes, bars, lane);

  const long long nfull = n / kTilePts;                    // full 128-point tiles
  const long long nwt = (n + kTilePts - 1) / kTilePts;     // warp tiles
  const char* Xb = reinterpret_cast<const char*>(X);
  const bool need_depth = MODE != 0 && ((depth != nullptr) || P.check_depth);
  const T Wf = (T)P.width, Hf = (T)P.height;      // image bounds in the pixel dtype (exact: < 2^24)
  const bool valid_words = (reinterpret_cast<uintptr_t>(valid_out) & 3) == 0;

  const long long stride = (long long)gridDim.x * kWarps;
  const long long t0 = (long long)blockIdx.x * kWarps + warp;
  if (lane == 0) {
    for (int k = 0; k < NST - 1; ++k) {
      const long long t = t0 + (long long)k * stride;
      if (t < nfull) ring.issue(k, Xb + t * kInBytes);
    }
  }

  long long local_count = 0;
  int k = 0;
  for (long long tile = t0; tile < nwt; tile += stride, ++k) {
    const long long nxt = tile + (NST - 1) * stride;
    if (lane == 0 && nxt < nfull) ring.issue(k + NST - 1, Xb + nxt * kInBytes);
    const long long first = tile * kTilePts;
    const long long rem = n - first;
    const int npts = rem < kTilePts ? (int)rem : kTilePts;
    char* st = ring.slot(k);
    if (npts == kTilePts) ring.wait(k);
    else {
      warp_load_ragged(Xb + first * 3 * sizeof(T), st, lane, kInBytes, npts * 3 * (int)sizeof(T));
      __syncwarp();
    }
    double Xp[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) stage_read_point<T>(st, j * 32 + lane, Xp[j]);
    __syncwarp();                       // slot may be refilled from here on

    T uo[4], vo[4];
    double dep[4], ud[4], vd[4];
    bool ok[4];
    project_many<DIST, 4>(P, Xp, ud, vd);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uo[j] = (T)ud[j];
      vo[j] = (T)vd[j];
      if (MODE != 0) {
        bool good = (j * 32 + lane) < npts;
        if (need_depth) {
          dep[j] = depth_of(P, Xp[j]);
          if (P.check_depth) good = good && (dep[j] >= P.min_depth);
        }
        if (P.check_bounds) good = good && OutPix<T>::in_bounds(uo[j], vo[j], Wf, Hf);
        ok[j] = good;
      }
    }

    if (npts == kTilePts) {
#pragma unroll
      for (int j = 0; j < 4; ++j) OutPix<T>::store(uv, first + j * 32 + lane, uo[j], vo[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j * 32 + lane < npts) OutPix<T>::store(uv, first + j * 32 + lane, uo[j], vo[j]);
    }
    if (MODE == 0) continue;

    unsigned mv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = j * 32 + lane;
      mv[j] = __ballot_sync(0xffffffffu, ok[j]);
      local_count += __popc(mv[j]);
      if (p < npts && depth != nullptr) st_stream_d1(depth + first + p, dep[j]);
    }
    if (valid_out != nullptr) {
      if (npts == kTilePts && valid_words) {
        store_valid_rows(valid_out + first, mv, 32, lane);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j * 32 + lane < npts) valid_out[first + j * 32 + lane] = ok[j] ? 1 : 0;
      }
    }
  }
  if (MODE == 1 && count != nullptr) {
    if (lane == 0 && local_count)
      atomicAdd(reinterpret_cast<unsigned long long*>(count), (unsigned long long)local_count);
  }
  // no masks: every point is valid (written from the device: a host-sourced copy would be
  // illegal under stream capture)
  if (MODE == 0 && count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *count = n;
}

// Phase clocks of the compaction kernel (diagnostic; compiled in with -DPBK_COMPACT_TIMING)
#ifdef PBK_COMPACT_TIMING
__device__ unsigned long long g_ct_clk[16];
#endif
#ifdef PBK_EA_TIMELINE
__device__ unsigned long long g_ea_ts[64][8];
__device__ unsigned long long g_ea_cta[4][512];      // per CTA: publish time of generations 20 / 21, offset-seen time of 20, SM id
__device__ __forceinline__ unsigned long long ea_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define EA_TS(gen, col) do { if ((gen) < 64) g_ea_ts[(gen)][(col)] = ea_now(); } while (0)
#define EA_TS_MAX(gen, col) do { if ((gen) < 64) atomicMax(&g_ea_ts[(gen)][(col)], ea_now()); } while (0)
#else
#define EA_TS(gen, col) do {} while (0)
#define EA_TS_MAX(gen, col) do {} while (0)
#endif
#ifdef PBK_COMPACT_TIMING
#define CT_T0() long long _t0 = clock64()
#define CT_ACC(slot)                                                        \
  do {                                                                      \
    const long long _t1 = clock64();                                        \
    if (lane == 0) atomicAdd(&g_ct_clk[slot], (unsigned long long)(_t1 - _t0)); \
    _t0 = _t1;                                                              \
  } while (0)
#else
#define CT_T0() do {} while (0)
#define CT_ACC(slot) do {} while (0)
#endif

// --------------------------------- projection + x[valid], single pass (look-back)
// Order-preserving compaction fused into the projection pass (no second read
// of X, no library scan): a two-level scan over 1024-point tiles, software-
// pipelined two tiles deep.
//
// Persistent, co-resident CTAs (cooperative launch) of 8 compute warps + 2 scan
// warps + 1 chain warp.  CTA b owns tiles b, b+G, b+2G ...: all CTAs work on one
// generation of G consecutive tiles at a time.
//   compute warps read tile i from a 2-stage TMA input ring (the last warp to
//                 consume a stage refills it), project it (thread t owns points
//                 t + 256 k), park uv / depth in one of three shared-memory
//                 buffers, publish the tile aggregate (the last warp to finish
//                 does it), then scatter tile i-2 whose base offset has been
//                 resolved meanwhile;
//   scan warp w   (tiles of parity w) scans the 32 (row, warp) valid counts of a
//                 finished tile; its exclusive offset is the prefix of its
//                 GENERATION plus the aggregates of the tiles of the same
//                 generation before it (<= G status words, all loads in flight
//                 together with the generation prefix): no tile waits for another
//                 tile's look-back;
//   chain warp    (CTA 0 only) is the one serial dependency: it sums the G
//                 aggregates of generation g as they appear and publishes the
//                 prefix of generation g+1 - one L2 round trip per generation.
// History: a per-tile decoupled look-back (every tile resolving against its
// predecessors' prefixes) made the prefix frontier advance one generation per ~3
// dependent L2 round trips, 0.100 ms; this form 0.095 ms.  Versions that waited for
// the prefix inside the tile, took tickets ahead of time, or published the aggregate
// from the scan warp ran 1.5-6x slower.
// Valid points are written at base + rank; lanes of a warp hold consecutive
// ranks, so the stores are contiguous runs.
// scratch layout (zeroed per call): [0] unused, [1 .. ntiles] tile aggregate words, then
// one prefix word per generation.
constexpr int kCtWarps = 8;                          // compute warps
constexpr int kCtThreads = kCtWarps * 32;            // 256 compute threads
// Shape knobs (scripts/build_variants.py builds A/B libraries from them).  Measured on C2
// (10 M points, flushed): 3 CTAs/SM x 3 buffers, one point at a time 0.0952 ms; 2 CTAs/SM x 5
// buffers with the projection batched in pairs (80 registers, no spills) 0.0932 ms - shipped;
// batched by four (spills) 0.0973, four scan warps 0.1096, 6 buffers + 3 scan warps 0.1239.
#ifndef PBK_CT_SCAN
#define PBK_CT_SCAN 2
#endif
#ifndef PBK_CT_BUFS
#define PBK_CT_BUFS 5
#endif
#ifndef PBK_CT_MINB
#define PBK_CT_MINB 2
#endif
constexpr int kCtScan = PBK_CT_SCAN;                           // scan warps (one look-back each in flight); must be <= buffers
constexpr int kCtBlock = kCtThreads + 32 * kCtScan + 32;     // + the chain warp (works in CTA 0 only)
constexpr int kCtPpt = 4;                            // points per thread
constexpr int kCtTile = kCtThreads * kCtPpt;         // 1024 points per tile
constexpr int kCtSegs = kCtPpt * kCtWarps;           // 32 (row, warp) segments
// parked-output buffers (pipeline depth = buffers - 1 tiles of slack between a
// tile's projection and its scatter); see the shape knobs above for what was measured.
template <typename T> struct CtBufs { static constexpr int value = PBK_CT_BUFS; };
static_assert(kCtScan <= CtBufs<float>::value, "a scan warp may only wait one mbarrier phase ahead");

template <typename T>
struct CompactSmem {
  static constexpr int kIn = kCtTile * 3 * (int)sizeof(T);             // one input stage
  static constexpr int kOut = kCtTile * (2 * (int)sizeof(T) + 8);      // parked uv + depth of one tile
  static constexpr int kBytes = 2 * kIn + CtBufs<T>::value * kOut;
};

template <typename T, bool DIST>
__global__ void __launch_bounds__(kCtBlock, sizeof(T) == 4 ? PBK_CT_MINB : 1)
project_compact_kernel(const T* __restrict__ X, long long n, T* __restrict__ uv,
                       double* __restrict__ depth, uint8_t* __restrict__ valid_out,
                       long long* __restrict__ count, unsigned long long* __restrict__ scratch,
                       const __grid_constant__ pbk_project_params P) {
  using SM = CompactSmem<T>;
  using UV = typename OutPix<T>::vec;
  constexpr int kCtBufs = CtBufs<T>::value;
  extern __shared__ __align__(128) char dyn_smem[];
  // barriers and aggregate slots are per parked buffer: tile i + kCtBufs cannot be
  // projected before tile i has been scattered, so a slot is never shared by two
  // live tiles however far the compute warps drift apart
  __shared__ __align__(8) uint64_t in_full[2], cnt_ready[kCtBufs], base_ready[kCtBufs];
  __shared__ long long s_base[kCtBufs];
  __shared__ int s_cnt[kCtBufs][kCtSegs], s_excl[kCtBufs][kCtSegs], s_aggsum[kCtBufs], s_arrived[kCtBufs], s_consumed[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ntiles = (n + kCtTile - 1) / kCtTile;
  unsigned long long* status = scratch + 1;                // one aggregate word per tile
  unsigned long long* gen_prefix = scratch + 1 + ntiles;   // one prefix word per generation of gridDim.x tiles
  // generation i of this CTA is tile i * gridDim.x + blockIdx.x
  const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  auto tile_of = [&](int i) -> long long { return (long long)i * gridDim.x + blockIdx.x; };

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&in_full[s], 1);
      s_consumed[s] = 0;
    }
    for (int q = 0; q < kCtBufs; ++q) {
      mbar_init(&cnt_ready[q], kCtWarps);
      mbar_init(&base_ready[q], 1);
      s_aggsum[q] = 0;
      s_arrived[q] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // input stage (i & 1) <- tile of generation i: one TMA bulk copy (plus a < 16 B
  // tail by hand), issued by a single thread
  auto feed = [&](int i) {
    if (i >= my_tiles) return;
    const int st = i & 1;
    const long long first = tile_of(i) * kCtTile;
    const int npts = (n - first) < kCtTile ? (int)(n - first) : kCtTile;
    const int bytes = npts * 3 * (int)sizeof(T), bulk = bytes & ~15;
    char* dst = dyn_smem + st * SM::kIn;
    const char* src = reinterpret_cast<const char*>(X + first * 3);
    for (int o = bulk; o < bytes; o += 4)
      *reinterpret_cast<uint32_t*>(dst + o) = *reinterpret_cast<const uint32_t*>(src + o);
    if (bulk > 0) {
      mbar_expect_tx(&in_full[st], (uint32_t)bulk);
      tma_load_1d(dst, src, (uint32_t)bulk, &in_full[st]);
    } else {
      mbar_arrive(&in_full[st]);
    }
  };
  if (tid == 0) {
    feed(0);
    feed(1);
  }

  if (warp == kCtWarps + kCtScan) {
    // ============================================================== chain warp
    // CTA 0 only: publishes the exclusive prefix of every generation of gridDim.x
    // tiles.  A generation's total needs only its tiles' aggregates (published by
    // the compute warps, no look-back involved), so this loop is the one serial
    // dependency of the kernel and it advances one L2 round trip per generation.
    if (blockIdx.x != 0) return;
    long long running = 0;
    for (long long g0 = 0, g = 0; g0 < ntiles; g0 += gridDim.x, ++g) {
      if (lane == 0) st_status(gen_prefix + g, kStatePrefix | (unsigned long long)running);
      const long long g1 = (g0 + gridDim.x) < ntiles ? (g0 + gridDim.x) : ntiles;
      const long long acc = sum_aggregates(status + g0, g1 - g0, lane);
      running += acc;
    }
    if (lane == 0 && count != nullptr) *count = running;
    return;
  }
  if (warp >= kCtWarps) {
    // ============================================================== scan warps
    const int w = warp - kCtWarps;                   // this warp resolves tiles w, w + kCtScan, ...
    for (int i = w; i < my_tiles; i += kCtScan) {
      const int q = i % kCtBufs;
      const uint32_t ph = (uint32_t)((i / kCtBufs) & 1);
      // tile i's segment counts -> exclusive offsets, aggregate
      CT_T0();
      mbar_wait(&cnt_ready[q], ph);
      CT_ACC(2);
      const int c = s_cnt[q][lane];
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
      }
      s_excl[q][lane] = inc - c;
      const long long agg = __shfl_sync(0xffffffffu, inc, 31);
      const long long tile = tile_of(i);              // its aggregate was published by the compute warps
      // exclusive offset = prefix of generation i (from the chain warp) + the
      // aggregates of the tiles of the SAME generation before this one.  No tile
      // waits on another tile's look-back: the only serial chain is the chain warp's
      // running sum of generation totals.
      const long long g0 = (long long)i * gridDim.x;  // first tile of generation i
      unsigned long long gp = ld_status(gen_prefix + i);          // in flight together with the aggregates
      const long long acc = sum_aggregates(status + g0, tile - g0, lane);
      while ((gp >> 62) == 0) gp = ld_status(gen_prefix + i);
      const long long excl = (long long)(gp & kValueMask) + acc;
      CT_ACC(3);
      __syncwarp();                         // every lane's s_excl store precedes the arrive
      if (lane == 0) {
        s_base[q] = excl;
        mbar_arrive(&base_ready[q]);
      }
    }
    return;
  }

  // ============================================================ compute warps
  const bool need_depth = (depth != nullptr) || P.check_depth;
  const T Wf = (T)P.width, Hf = (T)P.height;      // image bounds in the pixel dtype (exact: < 2^24)
  const unsigned lt = (1u << lane) - 1u;
  constexpr int kLag = kCtBufs - 1;         // a tile is scattered kLag iterations after its projection
  unsigned mh[kLag][kCtPpt];                // ballots of tiles i-1 .. i-kLag
  for (int i = 0; i < my_tiles + kLag; ++i) {
    unsigned m[kCtPpt];
    CT_T0();
    if (i < my_tiles) {
      const int s = i & 1, q = i % kCtBufs;
      mbar_wait(&in_full[s], (uint32_t)((i >> 1) & 1));
      if (warp == 0) CT_ACC(4);
      const long long tile = tile_of(i);
      const long long first = tile * kCtTile;
      const int npts = (n - first) < kCtTile ? (int)(n - first) : kCtTile;
      const T* sin = reinterpret_cast<const T*>(dyn_smem + s * SM::kIn);
      T xin[kCtPpt][3];
#pragma unroll
      for (int k = 0; k < kCtPpt; ++k) {
        const T* r = sin + 3 * (k * kCtThreads + tid);
        xin[k][0] = r[0]; xin[k][1] = r[1]; xin[k][2] = r[2];
      }
      __syncwarp();
      if (lane == 0 && atomicAdd(&s_consumed[s], 1) == kCtWarps - 1) {
        s_consumed[s] = 0;                             // last warp to read the stage refills it
        feed(i + 2);
      }
      UV* suv = reinterpret_cast<UV*>(dyn_smem + 2 * SM::kIn + q * SM::kOut);
      double* sdep = reinterpret_cast<double*>(dyn_smem + 2 * SM::kIn + q * SM::kOut + kCtTile * 2 * sizeof(T));
      const bool valid_words_tile = npts == kCtTile && (reinterpret_cast<uintptr_t>(valid_out) & 3) == 0;
#ifndef PBK_CT_BATCH
#define PBK_CT_BATCH 2
#endif
#if PBK_CT_BATCH > 0
      constexpr int kB = PBK_CT_BATCH;          // points projected side by side (shared reciprocal slow path)
      double ub[kCtPpt], vb[kCtPpt];
#pragma unroll
      for (int k0 = 0; k0 < kCtPpt; k0 += kB) {
        double Xb[kB][3], u2[kB], v2[kB];
#pragma unroll
        for (int j = 0; j < kB; ++j) {
          Xb[j][0] = (double)xin[k0 + j][0]; Xb[j][1] = (double)xin[k0 + j][1]; Xb[j][2] = (double)xin[k0 + j][2];
        }
        project_many<DIST, kB>(P, Xb, u2, v2);
#pragma unroll
        for (int j = 0; j < kB; ++j) { ub[k0 + j] = u2[j]; vb[k0 + j] = v2[j]; }
      }
#pragma unroll
      for (int k = 0; k < kCtPpt; ++k) {
        const int p = k * kCtThreads + tid;
        const double Xp[3] = {(double)xin[k][0], (double)xin[k][1], (double)xin[k][2]};
        const T uo = (T)ub[k], vo = (T)vb[k];
#else
#pragma unroll
      for (int k = 0; k < kCtPpt; ++k) {
        const int p = k * kCtThreads + tid;
        const double Xp[3] = {(double)xin[k][0], (double)xin[k][1], (double)xin[k][2]};
        double u, v;
        project_one<DIST>(P, Xp, u, v);
        const T uo = (T)u, vo = (T)v;
#endif
        bool good = p < npts;
        if (need_depth) {
          const double dep = depth_of(P, Xp);
          if (depth != nullptr) sdep[p] = dep;
          if (P.check_depth) good = good && (dep >= P.min_depth);
        }
        if (P.check_bounds) good = good && OutPix<T>::in_bounds(uo, vo, Wf, Hf);
        suv[p] = OutPix<T>::make(uo, vo);
        m[k] = __ballot_sync(0xffffffffu, good);
        if (lane == 0) s_cnt[q][k * kCtWarps + warp] = __popc(m[k]);
        if (valid_out != nullptr && !valid_words_tile && p < npts) valid_out[first + p] = good ? 1 : 0;
      }
      if (valid_out != nullptr && valid_words_tile) store_valid_rows(valid_out + first + warp * 32, m, kCtThreads, lane);
      __syncwarp();
      if (lane == 0) {
        // The tile aggregate leaves the moment the last compute warp is done - never
        // behind a look-back, or every successor tile would inherit that delay.
        int wsum = 0;
#pragma unroll
        for (int k = 0; k < kCtPpt; ++k) wsum += __popc(m[k]);
        atomicAdd(&s_aggsum[q], wsum);
        __threadfence_block();
        if (atomicAdd(&s_arrived[q], 1) == kCtWarps - 1) {
          __threadfence_block();
          const int total = atomicExch(&s_aggsum[q], 0);
          s_arrived[q] = 0;
          st_status(status + tile, kStateAgg | (unsigned long long)total);
        }
        mbar_arrive(&cnt_ready[q]);
      }
      if (warp == 0) CT_ACC(5);
    }
    if (i >= kLag) {
      // scatter tile i-kLag (parked by this same thread kLag iterations ago)
      const int j = i - kLag, q = j % kCtBufs;
      mbar_wait(&base_ready[q], (uint32_t)((j / kCtBufs) & 1));
      if (warp == 0) CT_ACC(6);
      const long long base = s_base[q];
      const UV* suv = reinterpret_cast<const UV*>(dyn_smem + 2 * SM::kIn + q * SM::kOut);
      const double* sdep =
          reinterpret_cast<const double*>(dyn_smem + 2 * SM::kIn + q * SM::kOut + kCtTile * 2 * sizeof(T));
#pragma unroll
      for (int k = 0; k < kCtPpt; ++k) {
        if ((mh[kLag - 1][k] >> lane) & 1u) {
          const int p = k * kCtThreads + tid;
          const long long o = base + s_excl[q][k * kCtWarps + warp] + __popc(mh[kLag - 1][k] & lt);
          OutPix<T>::store_vec(uv, o, suv[p]);
          if (depth != nullptr) st_stream_d1(depth + o, sdep[p]);
        }
      }
      if (warp == 0) CT_ACC(7);
    }
#pragma unroll
    for (int k = 0; k < kCtPpt; ++k) {
#pragma unroll
      for (int a = kLag - 1; a > 0; --a) mh[a][k] = mh[a - 1][k];
      mh[0][k] = m[k];
    }
  }
}


// ------------------------------------------------- projection + x[valid]
// Camera.project(check_bounds / check_depth) returns x[valid]: an order-preserving
// compaction (camera_utils.py:727-732).  Two kernels per chunk of X (<= 128 MB, so that what
// the first writes is still in L2 when the second reads it):
//
//  compact_classify_kernel  decides every point and compacts the valid ones.  Validity is a
//      handful of sign tests on LINEAR forms of X - |fx xc + (cx - W/2) zc| < (W/2)|zc| and
//      the v / depth analogues - evaluated in fp32 FMAs (~30 instructions per point instead of
//      ~90 in fp64) with a rigorous error band b = kappa (|X|+|Y|+|Z|+1): outside the band the
//      fp32 sign IS the reference's decision; the rare point inside it (~1e-5 of C2; also NaN,
//      inf, z ~ 0) takes the exact fp64 route of the uncompacted kernel.  A lane owns four
//      CONSECUTIVE points (48 contiguous bytes = three 16 B loads straight into registers: no
//      staging, no barriers), so its 0/1 flags leave as one 32-bit store.  The valid points of
//      a 128-point tile are packed, in order, through a warp-private shared-memory stage into
//      the tile's slot of the scratch region X' (16 B stores), the tile's count into tile_cnt.
//      Tiles are dealt to warps in contiguous RUNS; the last CTA to finish turns the run totals
//      into output offsets (exclusive scan of <= 9.5 k integers) and writes the total count.
//  project_compacted_kernel is the PLAIN dense projection over the X' tiles (L2 hits): fp64 in
//      OpenCV's order, fully coalesced uv / depth stores at run offset + running count.  No
//      ballots, no parking, no look-back, and the fp64 work is proportional to the number of
//      VALID points.
//
// HBM traffic = the algorithmic bytes (X once, flags, compacted outputs); X' lives in L2.
// History: a single-pass decoupled look-back over 1024-point tiles with parked
// outputs ran at 93 us on C2 however it was shaped - every tile's base depends on all earlier
// tiles and the fp64 projection of a tile had to finish before its count existed; a first
// two-phase form inside one cooperative kernel (TMA rings in both phases, 120 registers, 16
// warps per SM) spent 430 warp instructions per 128-point tile in the classification phase
// alone (61 us) - ring bookkeeping, not arithmetic.
constexpr int kCpWarps = 8;
constexpr int kCpThreads = kCpWarps * 32;
#ifndef PBK_CP_MINB_A
#define PBK_CP_MINB_A 4
#endif
#ifndef PBK_CP_CHUNK_MB
#define PBK_CP_CHUNK_MB 128
#endif
constexpr int kCpMaxRuns = 16384;

struct ClassifyParams {
  float z[4];        // zc            = z . (X, Y, Z, 1)
  float cu[4];       // fx xc + (cx - W/2) zc
  float cv[4];       // fy yc + (cy - H/2) zc
  float d[4];        // depth - min_depth
  float hw, hh;      // W/2, H/2
  float kappa;       // band = kappa * (|X| + |Y| + |Z| + 1)
};

// scratch layout (bytes): [0] arrival counter of the classify kernel (u32, zero between calls)
// [64 ..) run totals int[kCpMaxRuns], run offsets i64[kCpMaxRuns], then tile counts u8[tiles of a
// chunk] (256 B aligned), then the X' tiles of one chunk
constexpr size_t kCpOffTot = 64;
constexpr size_t kCpOffBase = kCpOffTot + (size_t)kCpMaxRuns * 4;
constexpr size_t kCpOffCnt = kCpOffBase + (size_t)kCpMaxRuns * 8;
static inline size_t cp_off_xprime(long long tiles) { return kCpOffCnt + (((size_t)tiles + 255) & ~(size_t)255); }

template <typename T> struct CpTile {
  static constexpr int kBytes = kTilePts * 3 * (int)sizeof(T);
  static constexpr int kVecs = kBytes / 16 / 32;        // 16 B vectors per lane: 3 (float) / 6 (double)
};

// exact validity of one point: the arithmetic of project_points_kernel<T, DIST, 1>
template <typename T, bool DIST>
__device__ __noinline__ bool exact_valid(const pbk_project_params& P, T x, T y, T z) {
  const double Xp[3] = {(double)x, (double)y, (double)z};
  bool good = true;
  if (P.check_depth) good = depth_of(P, Xp) >= P.min_depth;
  if (P.check_bounds) {
    double u, v;
    project_one<DIST>(P, Xp, u, v);
    good = good && OutPix<T>::in_bounds((T)u, (T)v, (T)P.width, (T)P.height);
  }
  return good;
}

__device__ __forceinline__ float lin4(const float (&c)[4], float x, float y, float z) {
  return fmaf(c[0], x, fmaf(c[1], y, fmaf(c[2], z, c[3])));
}

// Validity of one point by the fp32 linear forms.  decided = every test is farther than the
// error band from its threshold (then `in` IS the reference's decision); undecided points
// (borderline, NaN - band is then NaN -, inf, z ~ 0) take the exact route.  Branch-free: a
// disabled check has neutral constants (classify_params): depth form = +huge, |zc| form =
// +huge with hw = hh = 1 (pixel forms identically 0: always far inside).
__device__ __forceinline__ void classify_point(const ClassifyParams& C, float x, float y, float z, bool& in,
                                               bool& decided) {
  const float band = fmaf(C.kappa, fabsf(x), fmaf(C.kappa, fabsf(y), fmaf(C.kappa, fabsf(z), C.kappa)));
  const float ld = lin4(C.d, x, y, z);
  const float zc = fabsf(lin4(C.z, x, y, z));
  const float tu = fmaf(C.hw, zc, -fabsf(lin4(C.cu, x, y, z)));     // > 0: inside the image columns
  const float tv = fmaf(C.hh, zc, -fabsf(lin4(C.cv, x, y, z)));     // > 0: inside the image rows
  in = (ld > 0.f) && (tu > 0.f) && (tv > 0.f);
  decided = fminf(fminf(fabsf(ld), zc), fminf(fabsf(tu), fabsf(tv))) > band;
}

// the rare route for one row of a tile: lanes with an undecided point redo it exactly
template <typename T, bool DIST>
__device__ __noinline__ bool resolve_point(const pbk_project_params& P, const ClassifyParams& C, const T* pt) {
  const T x = pt[0], y = pt[1], z = pt[2];
  bool in = false, decided = false;
  if (!DIST) classify_point(C, (float)x, (float)y, (float)z, in, decided);
  if (!decided) in = exact_valid<T, DIST>(P, x, y, z);
  return in;
}

constexpr int kCpRing = 3;      // input tiles in flight per warp (TMA bulk copies: no registers held)

template <typename T, bool DIST>
__global__ void __launch_bounds__(kCpThreads, PBK_CP_MINB_A)
compact_classify_kernel(const T* __restrict__ X, long long n, uint8_t* __restrict__ valid_out,
                        unsigned char* __restrict__ scratch, long long xprime_off, int n_runs,
                        long long* __restrict__ count, const __grid_constant__ pbk_project_params P,
                        const __grid_constant__ ClassifyParams C) {
  constexpr int kInBytes = CpTile<T>::kBytes;
  constexpr int NV = CpTile<T>::kVecs;                     // 16 B vectors per lane and tile
  // dynamic shared memory per warp: kCpRing input slots + one compaction stage; then the ring barriers
  extern __shared__ __align__(128) char dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* const ringw = dyn_smem + warp * (kCpRing + 1) * kInBytes;
  T* const sg = reinterpret_cast<T*>(ringw + kCpRing * kInBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + kCpWarps * (kCpRing + 1) * kInBytes) + warp * kCpRing;
  WarpRing<kInBytes, kCpRing> ring;
  ring.init(ringw, bars, lane);
  unsigned* done_ctr = reinterpret_cast<unsigned*>(scratch);
  int* run_tot = reinterpret_cast<int*>(scratch + kCpOffTot);
  long long* run_base = reinterpret_cast<long long*>(scratch + kCpOffBase);
  uint8_t* tile_cnt = scratch + kCpOffCnt;
  char* xprime = reinterpret_cast<char*>(scratch) + xprime_off;

  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  const long long nfull = n / kTilePts;
  const int NW = (int)gridDim.x * kCpWarps, gw = (int)blockIdx.x * kCpWarps + warp;
  const char* Xb = reinterpret_cast<const char*>(X);
  const unsigned lt = (1u << lane) - 1u;
  const bool valid_words = (reinterpret_cast<uintptr_t>(valid_out) & 3) == 0;
  const uint64_t pol_stream = l2_policy_evict_first();     // X is read once: do not let it push X' out of L2

  int k = 0;                                               // ring sequence number
  for (int run = gw; run < n_runs; run += NW) {
    const long long ta = (long long)run * ntiles / n_runs, tb = ((long long)run + 1) * ntiles / n_runs;
    const int my = (int)(tb - ta);
    int wtot = 0;
    if (lane == 0)
      for (int d = 0; d < kCpRing - 1 && d < my; ++d)
        if (ta + d < nfull) ring.issue_hint(k + d, Xb + (ta + d) * kInBytes, pol_stream);
    for (int i = 0; i < my; ++i, ++k) {
      const long long tile = ta + i;
      const long long first = tile * kTilePts;
      const bool ragged = tile >= nfull;
      if (lane == 0 && i + kCpRing - 1 < my && tile + kCpRing - 1 < nfull)
        ring.issue_hint(k + kCpRing - 1, Xb + (tile + kCpRing - 1) * kInBytes, pol_stream);
      char* const st = ring.slot(k);
      if (ragged) {
        const int npts = (int)(n - first);
        warp_load_ragged(Xb + first * 3 * sizeof(T), st, lane, kInBytes, npts * 3 * (int)sizeof(T));
        __syncwarp();
        if (lane == 0) ring.skip(k);          // keep the slot's barrier phase in step
      }
      ring.wait(k);
      // lane l owns points l, l + 32, l + 64, l + 96 of the tile (3-word stride: conflict-free)
      const T* pts = reinterpret_cast<const T*>(st) + 3 * lane;
      unsigned mv[4];
      bool undecided = DIST;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bool in = false, decided = false;
        if (!DIST) classify_point(C, (float)pts[96 * j], (float)pts[96 * j + 1], (float)pts[96 * j + 2], in, decided);
        mv[j] = __ballot_sync(0xffffffffu, in);
        undecided = undecided || !decided;
      }
      if (__any_sync(0xffffffffu, undecided)) {
#pragma unroll 1
        for (int j = 0; j < 4; ++j) mv[j] = __ballot_sync(0xffffffffu, resolve_point<T, DIST>(P, C, pts + 96 * j));
      }
      if (ragged) {
        const int npts = (int)(n - first);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int live = npts - j * 32;
          mv[j] &= live >= 32 ? 0xffffffffu : (live <= 0 ? 0u : ((1u << live) - 1u));
        }
      }
      if (valid_out != nullptr) {
        if (!ragged && valid_words) {
          store_valid_rows(valid_out + first, mv, 32, lane);
        } else {
          const int npts = ragged ? (int)(n - first) : kTilePts;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j * 32 + lane < npts) valid_out[first + j * 32 + lane] = (uint8_t)((mv[j] >> lane) & 1u);
        }
      }
      // the tile's valid points, in order, to the warp's stage
      __syncwarp();                         // the previous tile's copy-out has finished reading the stage
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((mv[j] >> lane) & 1u) {
          T* dst = sg + 3 * (cnt + __popc(mv[j] & lt));
          dst[0] = pts[96 * j]; dst[1] = pts[96 * j + 1]; dst[2] = pts[96 * j + 2];
        }
        cnt += __popc(mv[j]);
      }
      __syncwarp();                         // (also: every lane is done with the ring slot)
      // stage -> the tile's X' slot: 16-byte vectors, rounded up (the slot is a full tile wide)
      {
        const int nvec = (cnt * 3 * (int)sizeof(T) + 15) >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(sg);
        uint4* d4 = reinterpret_cast<uint4*>(xprime + tile * kInBytes);
#pragma unroll
        for (int m = 0; m < NV; ++m) {
          const int v = m * 32 + lane;
          if (v < nvec) d4[v] = s4[v];
        }
      }
      if (lane == 0) tile_cnt[tile] = (uint8_t)cnt;       // cnt <= 128 (128 -> 0x80)
      wtot += cnt;
    }
    if (lane == 0) run_tot[run] = wtot;
  }

  // ---- the last CTA to finish turns the run totals into output offsets
  __shared__ int s_last;
  __shared__ long long s_scan[kCpThreads];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const long long base0 = (count != nullptr) ? *count : 0;       // valid points of the chunks before this one
  const int per = (n_runs + kCpThreads - 1) / kCpThreads;
  const int r0 = threadIdx.x * per, r1 = (r0 + per) < n_runs ? (r0 + per) : n_runs;
  long long mine = 0;
  for (int r = r0; r < r1; ++r) mine += __ldcg(run_tot + r);
  s_scan[threadIdx.x] = mine;
  __syncthreads();
  for (int o = 1; o < kCpThreads; o <<= 1) {
    const long long v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
    __syncthreads();
    s_scan[threadIdx.x] += v;
    __syncthreads();
  }
  long long run = base0 + s_scan[threadIdx.x] - mine;
  for (int r = r0; r < r1; ++r) {
    run_base[r] = run;
    run += __ldcg(run_tot + r);
  }
  if (threadIdx.x == kCpThreads - 1 && count != nullptr) *count = base0 + s_scan[kCpThreads - 1];
  if (threadIdx.x == 0) *done_ctr = 0u;                  // ready for the next launch
}

// The plain projection over the compacted tiles of one chunk.  B-warp w takes the runs w,
// w + (its grid's warps), ... of the classify kernel's partition.
template <typename T, bool DIST>
__global__ void __launch_bounds__(kWarps * 32)
project_compacted_kernel(long long n, T* __restrict__ uv, double* __restrict__ depth,
                         const unsigned char* __restrict__ scratch, long long xprime_off, int n_runs,
                         const __grid_constant__ pbk_project_params P) {
  constexpr int NST = RingCfg<T>::kStages;
  constexpr int kInBytes = CpTile<T>::kBytes;
  extern __shared__ __align__(128) char dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + kWarps * NST * kInBytes) + warp * NST;
  WarpRing<kInBytes, NST> ring;
  ring.init(dyn_smem + warp * NST * kInBytes, bars, lane);
  const long long* run_base = reinterpret_cast<const long long*>(scratch + kCpOffBase);
  const uint8_t* tile_cnt = scratch + kCpOffCnt;
  const char* xprime = reinterpret_cast<const char*>(scratch) + xprime_off;
  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  const int NW = (int)gridDim.x * kWarps, gw = (int)blockIdx.x * kWarps + warp;

  int k = 0;                      // ring sequence number
  for (int run = gw; run < n_runs; run += NW) {
    const long long ta = (long long)run * ntiles / n_runs, tb = ((long long)run + 1) * ntiles / n_runs;
    const int my = (int)(tb - ta);
    long long out = run_base[run];
    // counts of the run's tiles, 32 at a time in the lanes' registers
    int cnts = 0, cnts_lo = -1;
    auto count_of = [&](int i) -> int {       // all lanes call it with the same i, ascending
      if ((i >> 5) != cnts_lo) {
        cnts_lo = i >> 5;
        const int t = cnts_lo * 32 + lane;
        cnts = t < my ? (int)tile_cnt[ta + t] : 0;
      }
      return __shfl_sync(0xffffffffu, cnts, i & 31);
    };
    auto issue = [&](int i, int kk, int cn) {  // lane 0
      const uint32_t bytes = (uint32_t)((cn * 3 * (int)sizeof(T) + 15) & ~15);
      if (bytes) ring.issue_bytes(kk, xprime + (ta + i) * kInBytes, bytes);
      else ring.skip(kk);
    };
    // software pipeline: tile i + NST - 1 is requested while tile i is projected.  The counts of
    // the tiles in flight are kept in a small register queue so that count_of() is only ever
    // called with ascending i.
    int q[NST];
#pragma unroll
    for (int d = 0; d < NST; ++d) q[d] = 0;
#pragma unroll
    for (int d = 0; d < NST - 1; ++d) {
      if (d < my) {
        q[d] = count_of(d);
        if (lane == 0) issue(d, k + d, q[d]);
      }
    }
    for (int i = 0; i < my; ++i, ++k) {
      if (i + NST - 1 < my) {
        q[NST - 1] = count_of(i + NST - 1);
        if (lane == 0) issue(i + NST - 1, k + NST - 1, q[NST - 1]);
      }
      const int cnt = q[0];
#pragma unroll
      for (int d = 0; d < NST - 1; ++d) q[d] = q[d + 1];
      ring.wait(k);
      const char* st = ring.slot(k);
      const int rows = (cnt + 31) >> 5;
      const long long o0 = out;
      out += cnt;
      auto emit = [&](auto NR) {
        constexpr int R = decltype(NR)::value;
        double Xp[R][3], ud[R], vd[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (j * 32 + lane < cnt) stage_read_point<T>(st, j * 32 + lane, Xp[j]);
          else { Xp[j][0] = 0.0; Xp[j][1] = 0.0; Xp[j][2] = 1.0; }
        }
        project_many<DIST, R>(P, Xp, ud, vd);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (j * 32 + lane < cnt) {
            OutPix<T>::store(uv, o0 + j * 32 + lane, (T)ud[j], (T)vd[j]);
            if (depth != nullptr) st_stream_d1(depth + o0 + j * 32 + lane, depth_of(P, Xp[j]));
          }
        }
      };
      if (rows == 4) emit(std::integral_constant<int, 4>());
      else if (rows == 3) emit(std::integral_constant<int, 3>());
      else if (rows == 2) emit(std::integral_constant<int, 2>());
      else if (rows == 1) emit(std::integral_constant<int, 1>());
      __syncwarp();                       // slot may be refilled from here on
    }
  }
}

// ============================== compaction, single pass, aggregates published EARLY
// The look-back kernel above can publish a tile's count only after the tile's fp64 projection,
// so the prefix chain and the arithmetic are in series and the outputs have to be parked.
// Here the count comes from the fp32 classifier (exact by construction, see classify_point),
// a few hundred clocks after the tile's points land and kEaAhead TILES AHEAD of the projection:
//
//   compute warp (8 per CTA, 128 points of the CTA's 1024-point tile each), iteration i:
//     classify tile i + kEaAhead -> valid flags out, count to shared memory; the last of the
//       eight warps to get there publishes the tile's aggregate                      (early)
//     project tile i in fp64 (results stay in registers)
//     read the tile's output offset (one word, normally long there), scatter
//   chain warp (one: the last CTA of the grid holds nothing else): tile = i G + b, so all CTAs
//     work on generation i together; the chain warp collects the G aggregates of a generation as they appear, scans them and
//     writes every tile's exclusive offset.  It is the only place where generations are chained.
//
// The chain of generation i + kEaAhead therefore runs while generation i is being projected;
// nothing is parked, no CTA-wide barrier is executed after the prologue.
#ifndef PBK_EA_WARPS
#define PBK_EA_WARPS 8
#endif
constexpr int kEaWarps = PBK_EA_WARPS;      // compute warps per CTA (128 points each per tile)
static_assert(kEaWarps >= 1 && kEaWarps <= 16, "the offset sums run over 16-lane groups");
constexpr int kEaCtaPts = kEaWarps * kTilePts;
constexpr int kEaThreads = kEaWarps * 32;
constexpr int kEaSlots = 16;                // per-tile bookkeeping in shared memory (see the static_assert)
#ifndef PBK_EA_AHEAD
#define PBK_EA_AHEAD 2
#endif
constexpr int kEaAheadMax = PBK_EA_AHEAD;
#ifndef PBK_EA_MINB
#define PBK_EA_MINB 2
#endif
#ifndef PBK_EA_BATCH
#define PBK_EA_BATCH 4
#endif
// a warp in iteration k (writing the counts of tile k + kEaAhead) has seen the offset of tile k-1,
// so every warp of the CTA has classified tile k-1 and may still be reading the counts of tile
// k-1-kEaAhead: the slots of tiles k-1-kEaAhead .. k+kEaAhead must be distinct
static_assert(kEaAheadMax >= 1 && 2 * kEaAheadMax + 2 <= kEaSlots, "bookkeeping slots");

template <typename T> struct EarlyCfg {
  // float64 tiles are twice as large: two tiles ahead keep two CTAs per SM
  static constexpr int kAhead = sizeof(T) == 4 ? kEaAheadMax : (kEaAheadMax < 2 ? kEaAheadMax : 2);
  static constexpr int kStages = kAhead + (sizeof(T) == 4 ? 3 : 2);       // input tiles per warp
  static constexpr int kRing = kEaWarps * kStages * CpTile<T>::kBytes;
  static constexpr int kBytes = kRing + kEaWarps * kStages * 8 + kEaSlots * 8 + kEaSlots * kEaWarps * 4;
};
constexpr int kEaPer = 14;                  // tiles of a generation per chain lane: G <= 448
constexpr int kEaCarrySlots = 16;

// arrive on an mbarrier (release, CTA scope); true for the arrival that completes the phase
__device__ __forceinline__ bool mbar_arrive_is_last(uint64_t* bar) {
  uint64_t state;
  uint32_t pending;
  asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(smem_u32(bar)) : "memory");
  asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(state));
  return pending == 1u;
}

template <typename T, bool DIST>
__global__ void __launch_bounds__(kEaThreads, PBK_EA_MINB)
project_compact_early_kernel(const T* __restrict__ X, long long n, T* __restrict__ uv,
                             double* __restrict__ depth, uint8_t* __restrict__ valid_out,
                             long long* __restrict__ count, unsigned long long* __restrict__ status,
                             const __grid_constant__ pbk_project_params P,
                             const __grid_constant__ ClassifyParams C) {
  constexpr int NST = EarlyCfg<T>::kStages;
  constexpr int kEaAhead = EarlyCfg<T>::kAhead;
  constexpr int kInBytes = CpTile<T>::kBytes;
  extern __shared__ __align__(128) char dyn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* ring_bars = reinterpret_cast<uint64_t*>(dyn_smem + EarlyCfg<T>::kRing);
  uint64_t* arr = ring_bars + kEaWarps * NST;                         // [slot] 8 arrivals: the counts are in
  int* cnt_w = reinterpret_cast<int*>(arr + kEaSlots);              // [slot][warp] valid points

  const long long ncta = (n + kEaCtaPts - 1) / kEaCtaPts;
  const int G = (int)gridDim.x - 1, b = (int)blockIdx.x; // G compute CTAs + the chain CTA
  const int mine = b < G ? (int)((ncta - b + G - 1) / G) : 0;       // tiles of this CTA (G <= ncta)
  unsigned long long* const aggw = status;               // [ncta] tile aggregates
  unsigned long long* const prefw = status + ncta;       // [ncta] tile offsets

  if (b == G) {
    // ------------------------------------------------------------ chain CTA
    // Warp w scans the generations w, w + 8, ... (G consecutive tiles each, kEaPer or fewer per
    // lane): the loads, the wait for late aggregates and the scan inside a generation overlap
    // between the warps; only the carry is handed from generation to generation, through one
    // shared-memory word [seq : 24 | carry : 40].
    volatile unsigned long long* carry_s = reinterpret_cast<volatile unsigned long long*>(dyn_smem);
    if (threadIdx.x < kEaCarrySlots) carry_s[threadIdx.x] = 0ull;
    __syncthreads();
    const long long ngen = (ncta + G - 1) / G;
    const int per = (G + 31) / 32;                          // <= kEaPer (launch_compact_early)
    CT_T0();
    for (long long k = warp; k < ngen; k += kEaWarps) {
      const long long t0 = k * G;
      const int Gi = (ncta - t0) < G ? (int)(ncta - t0) : G;
      unsigned long long sw[kEaPer];
#pragma unroll
      for (int j = 0; j < kEaPer; ++j) {                    // lane owns `per` consecutive tiles
        const int idx = lane * per + j;
        sw[j] = (j < per && idx < Gi) ? ld_status(aggw + t0 + idx) : kStateAgg;
      }
      // poll every missing word of the lane together (one L2 round trip per round, not per word);
      // politely while the generation before this one is still open
      for (;;) {
        bool all = true;
#pragma unroll
        for (int j = 0; j < kEaPer; ++j) all = all && ((sw[j] >> 62) != 0);
        if (__all_sync(0xffffffffu, all)) break;
        const bool critical = k == 0 || ((carry_s[k % kEaCarrySlots] >> 40) == (unsigned long long)(k & 0xffffff));
        if (!critical) __nanosleep(200);
#pragma unroll
        for (int j = 0; j < kEaPer; ++j)
          if ((sw[j] >> 62) == 0) sw[j] = ld_status(aggw + t0 + lane * per + j);
      }
      if (lane == 0) EA_TS(k, 1);
      long long mine_sum = 0;
#pragma unroll
      for (int j = 0; j < kEaPer; ++j) mine_sum += (long long)(sw[j] & kValueMask);
      long long incl = mine_sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      CT_ACC(8);
      unsigned long long cw = 0;
      if (lane == 31) {
        if (k > 0) {
          do { cw = carry_s[k % kEaCarrySlots]; } while ((cw >> 40) != (unsigned long long)(k & 0xffffff));
          cw &= (1ull << 40) - 1;
        }
        carry_s[(k + 1) % kEaCarrySlots] = ((unsigned long long)((k + 1) & 0xffffff) << 40) | (cw + (unsigned long long)incl);
        if (k == ngen - 1) *count = (long long)cw + incl;
      }
      const long long carry = (long long)__shfl_sync(0xffffffffu, cw, 31);
      CT_ACC(9);
      long long run = carry + incl - mine_sum;
#pragma unroll
      for (int j = 0; j < kEaPer; ++j) {
        const int idx = lane * per + j;
        if (j < per && idx < Gi) st_status(prefw + t0 + idx, kStatePrefix | (unsigned long long)run);
        run += (long long)(sw[j] & kValueMask);
      }
      if (lane == 0) EA_TS(k, 2);
      CT_ACC(10);
    }
    return;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kEaSlots; ++s) mbar_init(&arr[s], kEaWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  WarpRing<kInBytes, NST> ring;
  ring.init(dyn_smem + warp * NST * kInBytes, ring_bars + warp * NST, lane);
  __syncthreads();

  // ---------------------------------------------------------------- compute warps
  const char* Xb = reinterpret_cast<const char*>(X);
  const unsigned lt = (1u << lane) - 1u;
  const bool valid_words = (reinterpret_cast<uintptr_t>(valid_out) & 3) == 0;
  CT_T0();
  auto first_of = [&](int i) -> long long { return ((long long)i * G + b) * kEaCtaPts + warp * kTilePts; };
  auto issue = [&](int i) {                                // lane 0
    const long long first = first_of(i);
    if (first + kTilePts <= n) ring.issue(i, Xb + first * 3 * sizeof(T));
  };
  // tile i of this warp: validity ballots of its four 32-point rows, flags out, count handed over
  auto classify = [&](int i, unsigned (&m)[4]) {
    const long long first = first_of(i);
    const long long rem = n - first;
    const int npts = rem >= kTilePts ? kTilePts : (rem > 0 ? (int)rem : 0);
    char* const st = ring.slot(i);
    if (npts == kTilePts) ring.wait(i);
    else {                                                 // only ever the last tile of the array
      warp_load_ragged(Xb + first * 3 * sizeof(T), st, lane, kInBytes, npts * 3 * (int)sizeof(T));
      __syncwarp();
    }
    CT_ACC(15);
    const T* pts = reinterpret_cast<const T*>(st) + 3 * lane;
    bool undecided = DIST;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool in = false, decided = false;
      if (!DIST) classify_point(C, (float)pts[96 * j], (float)pts[96 * j + 1], (float)pts[96 * j + 2], in, decided);
      m[j] = __ballot_sync(0xffffffffu, in);
      undecided = undecided || !decided;
    }
    if (__any_sync(0xffffffffu, undecided)) {
#pragma unroll 1
      for (int j = 0; j < 4; ++j) m[j] = __ballot_sync(0xffffffffu, resolve_point<T, DIST>(P, C, pts + 96 * j));
    }
    if (npts != kTilePts) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int live = npts - j * 32;
        m[j] &= live >= 32 ? 0xffffffffu : (live <= 0 ? 0u : ((1u << live) - 1u));
      }
    }
    {
      // lane 0 hands the count over; everything after it is warp-uniform again (a wait loop
      // inside a lane-0-only branch leaves the warp split for the rest of the iteration)
      const int slot = i & (kEaSlots - 1);
      bool last = false;
      if (lane == 0) {
        cnt_w[slot * kEaWarps + warp] = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
        last = mbar_arrive_is_last(&arr[slot]);
      }
      if (__shfl_sync(0xffffffffu, (int)last, 0)) {        // the tile's eight counts are in
        mbar_wait(&arr[slot], (uint32_t)((i / kEaSlots) & 1));      // (acquire; the phase is complete)
        int agg = lane < kEaWarps ? cnt_w[slot * kEaWarps + lane] : 0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) agg += __shfl_xor_sync(0xffffffffu, agg, o);
        if (lane == 0) {
          st_status(aggw + (long long)i * G + b, kStateAgg | (unsigned long long)agg);
          EA_TS_MAX(i, 5);
          if (b == 0) EA_TS(i, 0);
        }
      }
    }
    if (valid_out != nullptr) {
      if (npts == kTilePts && valid_words) {
        store_valid_rows(valid_out + first, m, 32, lane);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j * 32 + lane < npts) valid_out[first + j * 32 + lane] = (uint8_t)((m[j] >> lane) & 1u);
      }
    }
  };

  if (lane == 0)
    for (int d = 0; d < NST - 1 && d < mine; ++d) issue(d);
  unsigned mq[kEaAhead + 1][4];                            // ballots of tiles i .. i + kEaAhead
#pragma unroll
  for (int a = 0; a <= kEaAhead; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) mq[a][j] = 0u;
#pragma unroll
  for (int a = 0; a < kEaAhead; ++a)
    if (a < mine) classify(a, mq[a]);
  for (int i = 0; i < mine; ++i) {
    if (lane == 0 && i + NST - 1 < mine) issue(i + NST - 1);     // into the slot of tile i-1
    if (i + kEaAhead < mine) classify(i + kEaAhead, mq[kEaAhead]);
    CT_ACC(11);
    const unsigned (&mv)[4] = mq[0];
    const char* st = ring.slot(i);
    const bool any = (mv[0] | mv[1] | mv[2] | mv[3]) != 0u;
    T uo[4] = {}, vo[4] = {};
    double dep[4] = {};
    if (any) {
      // PBK_EA_BATCH points side by side: 4 interleaves four fp64 chains (fewest stalls per warp),
      // 2 halves the live registers (three CTAs per SM instead of two)
#pragma unroll
      for (int j0 = 0; j0 < 4; j0 += PBK_EA_BATCH) {
        double Xp[PBK_EA_BATCH][3], ud[PBK_EA_BATCH], vd[PBK_EA_BATCH];
#pragma unroll
        for (int j = 0; j < PBK_EA_BATCH; ++j) stage_read_point<T>(st, (j0 + j) * 32 + lane, Xp[j]);
        project_many<DIST, PBK_EA_BATCH>(P, Xp, ud, vd);
#pragma unroll
        for (int j = 0; j < PBK_EA_BATCH; ++j) {
          uo[j0 + j] = (T)ud[j];
          vo[j0 + j] = (T)vd[j];
          if (depth != nullptr) dep[j0 + j] = depth_of(P, Xp[j]);
        }
      }
    }
    __syncwarp();                                          // slot may be refilled from here on
    CT_ACC(12);
    // the tile's offset: published by the chain warp once every aggregate of the generation was
    // in; the counts of the warps before this one are in shared memory since the aggregate
    const int slot = i & (kEaSlots - 1);
    unsigned long long w;
    {
      if (lane == 0 && b == 0 && warp == 0) EA_TS(i, 4);
      const unsigned long long* pw = prefw + (long long)i * G + b;
      do { w = ld_status(pw); } while ((w >> 62) == 0);   // every lane, one address: the warp stays whole
      if (lane == 0 && b == 0 && warp == 0) EA_TS(i, 3);
      if (lane == 0 && warp == 0) EA_TS_MAX(i, 6);
    }
    mbar_wait(&arr[slot], (uint32_t)((i / kEaSlots) & 1));          // (acquire; completed long ago)
    int bsum = (lane < warp) ? cnt_w[slot * kEaWarps + lane] : 0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
    const long long o0 = (long long)(w & kValueMask) + __shfl_sync(0xffffffffu, bsum, 0);
    CT_ACC(13);
    if (any) {
      T* const uvb = uv + 2 * o0;
      double* const db = depth + o0;                       // (never dereferenced when depth is null)
      const bool has_depth = depth != nullptr;
      int run = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned bit = (mv[j] >> lane) & 1u;
        const int r = run + __popc(mv[j] & lt);
        OutPix<T>::store_if(bit, uvb + 2 * r, uo[j], vo[j]);
        st_stream_d1_if(bit && has_depth, db + r, dep[j]);
        run += __popc(mv[j]);
      }
    }
#pragma unroll
    for (int a = 0; a < kEaAhead; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) mq[a][j] = mq[a + 1][j];
    CT_ACC(14);
  }
}

template <typename TI, typename TO>
static int launch_transform(const void* X, int64_t n, const XformParams& P, void* Y,
                            cudaStream_t s, const DeviceInfo* di) {
  constexpr int NST = RingCfg<TI>::kStages;
  const int smem = kWarps * (NST * kTilePts * 3 * (int)sizeof(TI) + kTilePts * 3 * (int)sizeof(TO)) + kWarps * NST * 8;
  auto kern = transform_points_kernel<TI, TO>;
  static KernelCache kc;
  int per_sm;
  int rc = kc.setup(di, kern, kWarps * 32, smem, &per_sm);
  if (rc) return rc;
  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  const int grid = resident_grid(di, per_sm, (ntiles + kWarps - 1) / kWarps);
  kern<<<grid, kWarps * 32, smem, s>>>(static_cast<const TI*>(X), n, static_cast<TO*>(Y), P);
  PBK_CHECK_LAUNCH("transform_points_kernel");
  return PBK_OK;
}

template <typename T, bool DIST, int MODE>
static int launch_project(const void* X, int64_t n, const pbk_project_params& P, void* uv,
                          double* depth, uint8_t* valid, int64_t* count, cudaStream_t s,
                          const DeviceInfo* di) {
  constexpr int NST = RingCfg<T>::kStages;
  const int smem = kWarps * NST * kTilePts * 3 * (int)sizeof(T) + kWarps * NST * 8;
  auto kern = project_points_kernel<T, DIST, MODE>;
  static KernelCache kc;
  int per_sm;
  int rc = kc.setup(di, kern, kWarps * 32, smem, &per_sm);
  if (rc) return rc;
  const long long ncta = (n + kCtaPts - 1) / kCtaPts;
  const int grid = resident_grid(di, per_sm, ncta);
  kern<<<grid, kWarps * 32, smem, s>>>(static_cast<const T*>(X), n, static_cast<T*>(uv), depth, valid,
                                       reinterpret_cast<long long*>(count), P);
  PBK_CHECK_LAUNCH("project_points_kernel");
  return PBK_OK;
}

template <typename T, bool DIST>
static int launch_compact_lookback(const void* X, int64_t n, const pbk_project_params& P, void* uv,
                          double* depth, uint8_t* valid, int64_t* count, void* scratch, cudaStream_t s,
                          const DeviceInfo* di) {
  constexpr int smem = CompactSmem<T>::kBytes;
  auto kern = project_compact_kernel<T, DIST>;
  static KernelCache kc;
  int per_sm;
  int rc = kc.setup(di, kern, kCtBlock, smem, &per_sm);
  if (rc) return rc;
  const long long ntiles = (n + kCtTile - 1) / kCtTile;
  const int grid = resident_grid(di, per_sm, ntiles);
  PBK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(2 * ntiles + 4) * 8, s));       // status words
  const T* Xp = static_cast<const T*>(X);
  long long nn = n;
  T* uvp = static_cast<T*>(uv);
  long long* cnt = reinterpret_cast<long long*>(count);
  unsigned long long* scr = static_cast<unsigned long long*>(scratch);
  pbk_project_params Pc = P;
  void* args[] = {&Xp, &nn, &uvp, &depth, &valid, &cnt, &scr, &Pc};
  PBK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(kCtBlock), args,
                                       smem, s));
  return PBK_OK;
}


static void classify_params(const pbk_project_params& P, ClassifyParams* C) {
  // linear forms in double, rounded once to float; kappa covers the fp32 evaluation (<= 5 x
  // 2^-24 of max|coefficient| x (|X|+|Y|+|Z|+1), inputs rounded to float included), the fp32
  // rounding of (W/2)|zc|, and the final rounding of the reference's pixel to the output dtype
  // (2^-24 |u| |zc|), with a > 4x margin
  const double W = P.width, H = P.height;
  const double zr[4] = {P.R[6], P.R[7], P.R[8], P.t[2]};
  const double xr[4] = {P.R[0], P.R[1], P.R[2], P.t[0]};
  const double yr[4] = {P.R[3], P.R[4], P.R[5], P.t[1]};
  double mz = 0, mu = 0, mv = 0, md = 0;
  for (int i = 0; i < 4; ++i) {
    const double cu = P.fx * xr[i] + (P.cx - 0.5 * W) * zr[i];
    const double cv = P.fy * yr[i] + (P.cy - 0.5 * H) * zr[i];
    const double d = i < 3 ? P.Rz[i] : (P.tz - P.min_depth);
    C->z[i] = (float)zr[i]; C->cu[i] = (float)cu; C->cv[i] = (float)cv; C->d[i] = (float)d;
    mz = fmax(mz, fabs(zr[i]));
    mu = fmax(mu, fabs(P.fx * xr[i]) + (fabs(P.cx) + W) * fabs(zr[i]));
    mv = fmax(mv, fabs(P.fy * yr[i]) + (fabs(P.cy) + H) * fabs(zr[i]));
    md = fmax(md, i < 3 ? fabs(P.Rz[i]) : fabs(P.tz) + fabs(P.min_depth));
  }
  C->hw = (float)(0.5 * W);
  C->hh = (float)(0.5 * H);
  if (!P.check_depth) {                 // depth form: always provably inside
    C->d[0] = C->d[1] = C->d[2] = 0.f;
    C->d[3] = 1e30f;
    md = 0;
  }
  if (!P.check_bounds) {                // |zc| = 1e30 >> band, pixel forms 0 < 1e30 - band
    for (int i = 0; i < 3; ++i) C->z[i] = C->cu[i] = C->cv[i] = 0.f;
    C->z[3] = 1e30f;
    C->cu[3] = C->cv[3] = 0.f;
    C->hw = C->hh = 1.f;
    mu = mv = mz = 0;
  }
  C->kappa = (float)(4e-6 * fmax(fmax(mu + W * mz, mv + H * mz), fmax(md, mz)));
}

// points per chunk: both kernels of a chunk run back to back so that X' is read from L2
static long long compact_chunk_points(long long n, size_t elem) {
  long long mb = PBK_CP_CHUNK_MB;
  if (const char* e = getenv("PBK_COMPACT_CHUNK_MB")) {
    const long long v = atoll(e);
    if (v >= 1 && v <= 4096) mb = v;
  }
  long long cap = ((mb << 20) / (3 * (long long)elem)) / kTilePts * kTilePts;
  if (cap < kTilePts) cap = kTilePts;
  if (n <= cap) return n > 0 ? n : 1;
  const long long nchunks = (n + cap - 1) / cap;
  const long long per = (n + nchunks - 1) / nchunks;                       // equal chunks ...
  return (per + kTilePts - 1) / kTilePts * kTilePts;                      // ... of whole tiles
}

// which compaction: 1 = single pass with look-back (93 us on C2), 2 = classify + dense
// projection kernels (104 us), 3 = single pass with early aggregates.
// PBK_COMPACT_IMPL overrides.
#ifndef PBK_COMPACT_DEFAULT
#define PBK_COMPACT_DEFAULT 1
#endif
static int compact_impl() {
  const char* e = getenv("PBK_COMPACT_IMPL");           // read per call: tests switch it
  if (e && e[0] >= '1' && e[0] <= '3') return e[0] - '0';
  return PBK_COMPACT_DEFAULT;
}

template <typename T, bool DIST>
static int launch_compact_early(const void* X, int64_t n, const pbk_project_params& P, void* uv,
                                double* depth, uint8_t* valid, int64_t* count, void* scratch, cudaStream_t s,
                                const DeviceInfo* di) {
  constexpr int smem = EarlyCfg<T>::kBytes;
  auto kern = project_compact_early_kernel<T, DIST>;
  static KernelCache kc;
  int per_sm;
  int rc = kc.setup(di, kern, kEaThreads, smem, &per_sm);
  if (rc) return rc;
  const long long ncta = (n + kEaCtaPts - 1) / kEaCtaPts;
  int grid = resident_grid(di, per_sm, ncta + 1);                // compute CTAs + the chain CTA
  if (grid > 32 * kEaPer + 1) grid = 32 * kEaPer + 1;
  PBK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(2 * ncta + 2) * 8, s));          // aggregate + offset words
  ClassifyParams C;
  classify_params(P, &C);
  const T* Xp = static_cast<const T*>(X);
  long long nn = n;
  T* uvp = static_cast<T*>(uv);
  long long* cnt = reinterpret_cast<long long*>(count);
  unsigned long long* scr = static_cast<unsigned long long*>(scratch);
  pbk_project_params Pc = P;
  void* args[] = {&Xp, &nn, &uvp, &depth, &valid, &cnt, &scr, &Pc, &C};
  PBK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(kEaThreads), args,
                                       smem, s));
  return PBK_OK;
}

template <typename T, bool DIST>
static int launch_compact(const void* X, int64_t n, const pbk_project_params& P, void* uv,
                          double* depth, uint8_t* valid, int64_t* count, void* scratch, cudaStream_t s,
                          const DeviceInfo* di) {
  const int impl = compact_impl();
  if (impl == 1) return launch_compact_lookback<T, DIST>(X, n, P, uv, depth, valid, count, scratch, s, di);
  if (impl == 3) return launch_compact_early<T, DIST>(X, n, P, uv, depth, valid, count, scratch, s, di);
  constexpr int NST = RingCfg<T>::kStages;
  constexpr int smem_a = kCpWarps * (kCpRing + 1) * CpTile<T>::kBytes + kCpWarps * kCpRing * 8;
  constexpr int smem_b = kWarps * NST * CpTile<T>::kBytes + kWarps * NST * 8;
  auto kern_a = compact_classify_kernel<T, DIST>;
  auto kern_b = project_compacted_kernel<T, DIST>;
  static KernelCache kc_a, kc_b;
  int per_sm_a, per_sm_b;
  int rc = kc_a.setup(di, kern_a, kCpThreads, smem_a, &per_sm_a);
  if (rc) return rc;
  rc = kc_b.setup(di, kern_b, kWarps * 32, smem_b, &per_sm_b);
  if (rc) return rc;
  ClassifyParams C;
  classify_params(P, &C);
  unsigned char* scr = static_cast<unsigned char*>(scratch);
  long long* cnt = reinterpret_cast<long long*>(count);
  const long long chunk = compact_chunk_points(n, sizeof(T));
  const long long chunk_tiles = (chunk + kTilePts - 1) / kTilePts;
  const long long xoff = (long long)cp_off_xprime(chunk_tiles);
  PBK_CUDA(cudaMemsetAsync(scratch, 0, 64, s));                  // the arrival counter
  PBK_CUDA(cudaMemsetAsync(count, 0, 8, s));                     // running total across chunks
  for (long long c0 = 0; c0 < n; c0 += chunk) {
    const long long nc = (n - c0) < chunk ? (n - c0) : chunk;
    const long long ntiles = (nc + kTilePts - 1) / kTilePts;
    const int grid_a = resident_grid(di, per_sm_a, (ntiles + kCpWarps - 1) / kCpWarps);
    long long runs = (long long)grid_a * kCpWarps;
    if (runs > ntiles) runs = ntiles;
    if (runs > kCpMaxRuns) runs = kCpMaxRuns;
    const int grid_b = resident_grid(di, per_sm_b, (runs + kWarps - 1) / kWarps);
    const T* Xc = static_cast<const T*>(X) + c0 * 3;
    kern_a<<<grid_a, kCpThreads, smem_a, s>>>(Xc, nc, valid ? valid + c0 : nullptr, scr, xoff, (int)runs, cnt, P, C);
    PBK_CHECK_LAUNCH("compact_classify_kernel");
    // the projection writes at absolute offsets (run offsets include the chunks before)
    kern_b<<<grid_b, kWarps * 32, smem_b, s>>>(nc, static_cast<T*>(uv), depth, scr, xoff, (int)runs, P);
    PBK_CHECK_LAUNCH("project_compacted_kernel");
  }
  return PBK_OK;
}

}  // namespace pbk

extern "C" {

int pbk_transform_points(const void* X, int in_dtype, int64_t n, const double* Rt, double scale,
                         void* Y, int out_dtype, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && Rt != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE((in_dtype == PBK_F32 || in_dtype == PBK_F64) && (out_dtype == PBK_F32 || out_dtype == PBK_F64),
              PBK_EINVAL, "points must be float32 or float64");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(X != nullptr && Y != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(X) && aligned16(Y), PBK_EALIGN, "device pointers must be 16-byte aligned");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  XformParams P;
  for (int i = 0; i < 9; ++i) P.R[i] = Rt[i];
  for (int i = 0; i < 3; ++i) P.t[i] = Rt[9 + i];
  P.scale = scale;
  P.has_scale = (scale != 1.0);
  cudaStream_t s = as_stream(stream);
  if (in_dtype == PBK_F32 && out_dtype == PBK_F64) return launch_transform<float, double>(X, n, P, Y, s, di);
  if (in_dtype == PBK_F32 && out_dtype == PBK_F32) return launch_transform<float, float>(X, n, P, Y, s, di);
  if (in_dtype == PBK_F64 && out_dtype == PBK_F64) return launch_transform<double, double>(X, n, P, Y, s, di);
  return launch_transform<double, float>(X, n, P, Y, s, di);
}

#ifdef PBK_COMPACT_TIMING
#endif
#ifdef PBK_EA_TIMELINE
PBK_API int pbk_debug_compact_cta(unsigned long long out[2048]) {
  return cudaMemcpyFromSymbol(out, pbk::g_ea_cta, sizeof(unsigned long long) * 2048) == cudaSuccess ? 0 : -2;
}
PBK_API int pbk_debug_compact_ts(unsigned long long out[512], int reset) {
  if (cudaMemcpyFromSymbol(out, pbk::g_ea_ts, sizeof(unsigned long long) * 512) != cudaSuccess) return -2;
  if (reset) {
    static unsigned long long z[512] = {0};
    if (cudaMemcpyToSymbol(pbk::g_ea_ts, z, sizeof(z)) != cudaSuccess) return -2;
  }
  return 0;
}
#endif
#ifdef PBK_COMPACT_TIMING
PBK_API int pbk_debug_compact_clocks(unsigned long long out[16], int reset) {
  if (cudaMemcpyFromSymbol(out, pbk::g_ct_clk, sizeof(unsigned long long) * 16) != cudaSuccess) return -2;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(pbk::g_ct_clk, z, sizeof(z)) != cudaSuccess) return -2;
  }
  return 0;
}
#endif

size_t pbk_project_scratch_bytes(int64_t n, int in_dtype) {
  if (n < 0) n = 0;
  const size_t elem = in_dtype == PBK_F64 ? 8 : 4;
  const long long chunk = pbk::compact_chunk_points(n, elem);
  const long long tiles = (chunk + pbk::kTilePts - 1) / pbk::kTilePts;
  // two-kernel form: control words + run totals / offsets + per-tile counts + the X' tiles of one chunk
  const size_t two = pbk::cp_off_xprime(tiles) + (size_t)tiles * pbk::kTilePts * 3 * elem;
  // single-pass form: tile status words + one prefix word per generation
  const size_t one = (size_t)(2 * ((n + pbk::kCtaPts - 1) / pbk::kCtaPts) + 4) * 8;
  return pbk::compact_impl() != 2 ? one : (two > one ? two : one);
}

int pbk_project_points(const void* X, int in_dtype, int64_t n, const pbk_project_params* params,
                       void* uv, double* depth, uint8_t* valid, int64_t* count, void* scratch,
                       pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && params != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(in_dtype == PBK_F32 || in_dtype == PBK_F64, PBK_EINVAL, "points must be float32 or float64");
  const pbk_project_params& P = *params;
  PBK_REQUIRE(!P.check_bounds || (P.width > 0 && P.height > 0), PBK_EINVAL,
              "check_bounds cannot proceed. Camera.shape is not set");
  PBK_REQUIRE(P.width < (1 << 24) && P.height < (1 << 24), PBK_EINVAL, "image shape too large");
  PBK_REQUIRE(!P.compact || (count != nullptr && scratch != nullptr), PBK_EINVAL,
              "compact output needs count and scratch");
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    if (count != nullptr) PBK_CUDA(cudaMemsetAsync(count, 0, 8, s));
    return PBK_OK;
  }
  PBK_REQUIRE(X != nullptr && uv != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(X) && aligned16(uv) && aligned16(depth) && aligned16(scratch), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  const bool plain = !P.compact && !P.check_depth && !P.check_bounds && depth == nullptr && valid == nullptr;
  // plain: the kernel writes count = n itself; masks: the warps add into a zeroed word;
  // compact: the kernel writes the total
  if (!P.compact && !plain && count != nullptr) PBK_CUDA(cudaMemsetAsync(count, 0, 8, s));
  const bool dist = P.has_distortion != 0;
  const int mode = plain ? 0 : (P.compact ? 2 : 1);
#define PBK_DISPATCH(T)                                                                       \
  switch (mode * 2 + (dist ? 1 : 0)) {                                                        \
    case 0: return launch_project<T, false, 0>(X, n, P, uv, depth, valid, count, s, di);      \
    case 1: return launch_project<T, true, 0>(X, n, P, uv, depth, valid, count, s, di);       \
    case 2: return launch_project<T, false, 1>(X, n, P, uv, depth, valid, count, s, di);      \
    case 3: return launch_project<T, true, 1>(X, n, P, uv, depth, valid, count, s, di);       \
    case 4: return launch_compact<T, false>(X, n, P, uv, depth, valid, count, scratch, s, di); \
    default: return launch_compact<T, true>(X, n, P, uv, depth, valid, count, scratch, s, di); \
  }
  if (in_dtype == PBK_F32) { PBK_DISPATCH(float) }
  PBK_DISPATCH(double)
#undef PBK_DISPATCH
}

}  // extern "C"
