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
};

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

// ------------------------------------------------- projection + x[valid]
// Camera.project(check_bounds / check_depth) returns x[valid]: an order-preserving
// compaction (camera_utils.py:727-732).  Two phases per chunk of X inside ONE persistent
// cooperative kernel, one grid barrier per chunk:
//
//  phase A  classifies every point of the chunk.  Validity is a handful of sign tests on
//           LINEAR forms of X - |fx xc + (cx - W/2) zc| < (W/2)|zc| and the v / depth
//           analogues - evaluated in fp32 FMAs (30 instructions per point instead of ~90 in
//           fp64) with a rigorous error band b = kappa (|X|+|Y|+|Z|+1): outside the band the
//           fp32 sign IS the reference's decision; the rare point inside it (~1e-5 of C2, also
//           NaN / inf / z ~ 0) takes the exact fp64 route of the uncompacted kernel.  The
//           valid points' coordinates are written, in order, to the front of the warp's own
//           scratch region X' (tile-run-local compaction: no offsets from other warps are
//           needed), the 0/1 flags to valid_out, and the warp's count to shared memory.
//  barrier  every CTA publishes its count; after the barrier each CTA reads the <= 2 x 148
//           counts, and a warp's output offset is chunk base + counts before it.
//  phase B  is the PLAIN dense projection over X' (L2-resident: a chunk's X' is <= 40 MB):
//           fp64 in OpenCV's order, fully coalesced uv / depth stores at the final offsets.
//           No ballots, no parking, no look-back, and the fp64 work is proportional to the
//           number of VALID points.
//
// HBM traffic = the algorithmic bytes (X once, flags, compacted outputs); X' lives in L2.
// History: a single-pass decoupled look-back over 1024-point tiles with
// parked outputs ran at 93 us on C2 however it was shaped - every tile's base depends on all
// earlier tiles, and the fp64 projection of a tile had to finish before its count existed.
constexpr int kCpWarps = 8;
constexpr int kCpThreads = kCpWarps * 32;
#ifndef PBK_CP_MINB
#define PBK_CP_MINB 2
#endif
#ifndef PBK_CP_CHUNK_MB
#define PBK_CP_CHUNK_MB 40
#endif

struct ClassifyParams {
  float z[4];        // zc            = z . (X, Y, Z, 1)
  float cu[4];       // fx xc + (cx - W/2) zc
  float cv[4];       // fy yc + (cy - H/2) zc
  float d[4];        // depth - min_depth
  float hw, hh;      // W/2, H/2
  float kappa;       // band = kappa * (|X| + |Y| + |Z| + 1)
};

// scratch layout of the compacting kernel (bytes): [0] grid-barrier counter (u32), [64 ..) two
// arrays of kCpMaxGrid CTA counts (call parity of the chunk), [kCpCtlBytes ..) the X' region
constexpr int kCpMaxGrid = 2048;
constexpr size_t kCpCtlBytes = 64 + 2 * kCpMaxGrid * 4;       // 16448, a multiple of 64

template <typename T> struct CpTile {
  static constexpr int kBytes = kTilePts * 3 * (int)sizeof(T);
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

// all CTAs of the (co-resident) grid; the counter only grows: target = gridDim.x * (#barriers so far)
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

template <typename T, bool DIST>
__global__ void __launch_bounds__(kCpThreads, PBK_CP_MINB)
project_compact_kernel(const T* __restrict__ X, long long n, T* __restrict__ uv, double* __restrict__ depth,
                       uint8_t* __restrict__ valid_out, long long* __restrict__ count,
                       unsigned char* __restrict__ scratch, long long tiles_per_chunk,
                       const __grid_constant__ pbk_project_params P, const __grid_constant__ ClassifyParams C) {
  constexpr int NST = RingCfg<T>::kStages;
  constexpr int kInBytes = CpTile<T>::kBytes;
  extern __shared__ __align__(128) char dyn_smem[];
  __shared__ int s_wtot[kCpWarps];
  __shared__ long long s_red[2][kCpWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + kCpWarps * NST * kInBytes) + warp * NST;
  WarpRing<kInBytes, NST> ring;
  ring.init(dyn_smem + warp * NST * kInBytes, bars, lane);

  unsigned* bar_ctr = reinterpret_cast<unsigned*>(scratch);
  int* cta_tot = reinterpret_cast<int*>(scratch + 64);
  char* xprime = reinterpret_cast<char*>(scratch + kCpCtlBytes);

  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  const long long nfull = n / kTilePts;
  const int nchunks = (int)((ntiles + tiles_per_chunk - 1) / tiles_per_chunk);
  const long long NW = (long long)gridDim.x * kCpWarps, gw = (long long)blockIdx.x * kCpWarps + warp;
  // this warp's run of tiles inside ANY chunk (clipped to the chunk's length): the same
  // offsets in every chunk, so its X' region is private to it for the whole kernel
  const long long run_a = gw * tiles_per_chunk / NW, run_b = (gw + 1) * tiles_per_chunk / NW;
  char* const xp_mine = xprime + run_a * kInBytes;
  const char* Xb = reinterpret_cast<const char*>(X);
  const unsigned lt = (1u << lane) - 1u;
  const bool valid_words = (reinterpret_cast<uintptr_t>(valid_out) & 3) == 0;
  const uint64_t pol_stream = l2_policy_evict_first();

  long long out_base = 0;          // valid points of the chunks before this one (identical in all CTAs)
  int k = 0;                       // ring sequence number, continues across phases and chunks
  for (int c = 0; c < nchunks; ++c) {
    const long long chunk_t0 = (long long)c * tiles_per_chunk;
    const long long chunk_tiles = (ntiles - chunk_t0) < tiles_per_chunk ? (ntiles - chunk_t0) : tiles_per_chunk;
    const long long a = run_a < chunk_tiles ? run_a : chunk_tiles, b = run_b < chunk_tiles ? run_b : chunk_tiles;
    const int my = (int)(b - a);

    // ================================================================ phase A
    int wtot = 0;
    {
      const long long t0 = chunk_t0 + a;
      if (lane == 0)
        for (int i = 0; i < NST - 1 && i < my; ++i)
          if (t0 + i < nfull) ring.issue_hint(k + i, Xb + (t0 + i) * kInBytes, pol_stream);
      for (int i = 0; i < my; ++i, ++k) {
        const long long tile = t0 + i;
        if (lane == 0 && i + NST - 1 < my && tile + NST - 1 < nfull)
          ring.issue_hint(k + NST - 1, Xb + (tile + NST - 1) * kInBytes, pol_stream);
        const long long first = tile * kTilePts;
        const int npts = (n - first) < kTilePts ? (int)(n - first) : kTilePts;
        char* st = ring.slot(k);
        if (tile >= nfull) {
          warp_load_ragged(Xb + first * 3 * sizeof(T), st, lane, kInBytes, npts * 3 * (int)sizeof(T));
          __syncwarp();
          if (lane == 0) ring.skip(k);        // keep the slot's barrier phase in step
        }
        ring.wait(k);
        T xin[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const T* r = reinterpret_cast<const T*>(st) + 3 * (j * 32 + lane);
          xin[j][0] = r[0]; xin[j][1] = r[1]; xin[j][2] = r[2];
        }
        __syncwarp();                       // slot may be refilled from here on
        bool ok[4], amb[4];
        bool any_amb = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (DIST) {
            ok[j] = false;
            amb[j] = true;
          } else {
            const float x = (float)xin[j][0], y = (float)xin[j][1], z = (float)xin[j][2];
            const float band = C.kappa * (fabsf(x) + fabsf(y) + fabsf(z) + 1.0f);
            bool in = true, out = false;
            if (P.check_depth) {
              const float ld = lin4(C.d, x, y, z);
              in = ld > band;
              out = ld < -band;
            }
            if (P.check_bounds) {
              const float zc = fabsf(lin4(C.z, x, y, z));
              const float lu = fabsf(lin4(C.cu, x, y, z)), lv = fabsf(lin4(C.cv, x, y, z));
              const float hw = C.hw * zc, hh = C.hh * zc;
              const bool zdef = zc > band;
              in = in && zdef && (lu < hw - band) && (lv < hh - band);
              out = out || (zdef && ((lu > hw + band) || (lv > hh + band)));
            }
            ok[j] = in && !out;
            amb[j] = !(in || out);          // NaN / inf / borderline: neither is provable
          }
          if (j * 32 + lane >= npts) { ok[j] = false; amb[j] = false; }
          any_amb = any_amb || amb[j];
        }
        if (__any_sync(0xffffffffu, any_amb)) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (amb[j]) ok[j] = exact_valid<T, DIST>(P, xin[j][0], xin[j][1], xin[j][2]);
        }
        unsigned mv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) mv[j] = __ballot_sync(0xffffffffu, ok[j]);
        if (valid_out != nullptr) {
          if (npts == kTilePts && valid_words) {
            store_valid_rows(valid_out + first, mv, 32, lane);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j * 32 + lane < npts) valid_out[first + j * 32 + lane] = ok[j] ? 1 : 0;
          }
        }
        // the warp's valid points, in order, to the front of its X' region
        T* xp = reinterpret_cast<T*>(xp_mine);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ok[j]) {
            T* dst = xp + 3ll * (wtot + __popc(mv[j] & lt));
            dst[0] = xin[j][0]; dst[1] = xin[j][1]; dst[2] = xin[j][2];
          }
          wtot += __popc(mv[j]);
        }
      }
    }
    if (lane == 0) s_wtot[warp] = wtot;
    // X' was written with generic-proxy stores and is read back by TMA (async proxy)
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kCpWarps; ++w) tot += s_wtot[w];
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(cta_tot + (c & 1) * kCpMaxGrid + blockIdx.x), "r"(tot) : "memory");
    }
    grid_barrier(bar_ctr, gridDim.x * (unsigned)(c + 1));

    // ================================================================ offsets
    long long before = 0, total = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kCpThreads) {
      int v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cta_tot + (c & 1) * kCpMaxGrid + i) : "memory");
      total += v;
      if (i < (int)blockIdx.x) before += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane == 0) { s_red[0][warp] = before; s_red[1][warp] = total; }
    __syncthreads();
    long long my_out = out_base;
#pragma unroll
    for (int w = 0; w < kCpWarps; ++w) {
      my_out += s_red[0][w];
      out_base += s_red[1][w];
      if (w < warp) my_out += s_wtot[w];
    }

    // ================================================================ phase B
    {
      asm volatile("fence.proxy.async;" ::: "memory");
      const int nt = (wtot + kTilePts - 1) / kTilePts;
      auto bytes_of = [&](int t) -> uint32_t {
        const int pts = (wtot - t * kTilePts) < kTilePts ? (wtot - t * kTilePts) : kTilePts;
        return (uint32_t)((pts * 3 * (int)sizeof(T) + 15) & ~15);
      };
      if (lane == 0)
        for (int t = 0; t < NST - 1 && t < nt; ++t) ring.issue_bytes(k + t, xp_mine + (long long)t * kInBytes, bytes_of(t));
      for (int t = 0; t < nt; ++t, ++k) {
        if (lane == 0 && t + NST - 1 < nt)
          ring.issue_bytes(k + NST - 1, xp_mine + (long long)(t + NST - 1) * kInBytes, bytes_of(t + NST - 1));
        const int pts = (wtot - t * kTilePts) < kTilePts ? (wtot - t * kTilePts) : kTilePts;
        ring.wait(k);
        char* st = ring.slot(k);
        double Xp[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j * 32 + lane < pts) stage_read_point<T>(st, j * 32 + lane, Xp[j]);
          else { Xp[j][0] = 0.0; Xp[j][1] = 0.0; Xp[j][2] = 1.0; }
        }
        __syncwarp();
        double ud[4], vd[4];
        project_many<DIST, 4>(P, Xp, ud, vd);
        const long long o0 = my_out + (long long)t * kTilePts;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j * 32 + lane < pts) {
            OutPix<T>::store(uv, o0 + j * 32 + lane, (T)ud[j], (T)vd[j]);
            if (depth != nullptr) st_stream_d1(depth + o0 + j * 32 + lane, depth_of(P, Xp[j]));
          }
        }
      }
    }
    __syncthreads();        // s_wtot / s_red are rewritten by the next chunk
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && count != nullptr) *count = out_base;
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
  C->kappa = (float)(4e-6 * fmax(fmax(mu + W * mz, mv + H * mz), fmax(md, mz)));
}

static long long compact_tiles_per_chunk(long long n, size_t elem) {
  long long mb = PBK_CP_CHUNK_MB;
  if (const char* e = getenv("PBK_COMPACT_CHUNK_MB")) {
    const long long v = atoll(e);
    if (v >= 1 && v <= 4096) mb = v;
  }
  const long long tile_bytes = (long long)kTilePts * 3 * (long long)elem;
  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  long long cap = (mb << 20) / tile_bytes;
  if (cap < 1) cap = 1;
  const long long nchunks = (ntiles + cap - 1) / cap;
  long long tpc = nchunks > 0 ? (ntiles + nchunks - 1) / nchunks : 1;       // equal chunks
  return tpc < 1 ? 1 : tpc;
}

template <typename T, bool DIST>
static int launch_compact(const void* X, int64_t n, const pbk_project_params& P, void* uv,
                          double* depth, uint8_t* valid, int64_t* count, void* scratch, cudaStream_t s,
                          const DeviceInfo* di) {
  constexpr int NST = RingCfg<T>::kStages;
  constexpr int smem = kCpWarps * NST * CpTile<T>::kBytes + kCpWarps * NST * 8;
  auto kern = project_compact_kernel<T, DIST>;
  static KernelCache kc;
  int per_sm;
  int rc = kc.setup(di, kern, kCpThreads, smem, &per_sm);
  if (rc) return rc;
  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  int grid = resident_grid(di, per_sm, (ntiles + kCpWarps - 1) / kCpWarps);
  if (grid > kCpMaxGrid) grid = kCpMaxGrid;
  const T* Xp = static_cast<const T*>(X);
  long long nn = n;
  T* uvp = static_cast<T*>(uv);
  long long* cnt = reinterpret_cast<long long*>(count);
  unsigned char* scr = static_cast<unsigned char*>(scratch);
  long long tpc = compact_tiles_per_chunk(n, sizeof(T));
  pbk_project_params Pc = P;
  ClassifyParams C;
  classify_params(P, &C);
  PBK_CUDA(cudaMemsetAsync(scratch, 0, 64, s));                  // the grid-barrier counter
  void* args[] = {&Xp, &nn, &uvp, &depth, &valid, &cnt, &scr, &tpc, &Pc, &C};
  PBK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(kCpThreads), args,
                                       smem, s));
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
  const long long tpc = pbk::compact_tiles_per_chunk(n, elem);
  // control words + the X' region of one chunk (+ one tile of slack for the 16-byte rounding of
  // the last bulk copy)
  return pbk::kCpCtlBytes + (size_t)(tpc + 1) * pbk::kTilePts * 3 * elem;
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
