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
// Order-preserving compaction (x[valid]) is fused: CTA tiles of 1024 points
// take tickets in order and chain their valid counts with a decoupled
// look-back, so the compacted outputs are written in the same pass.
#include "pbk_common.cuh"

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
  static __device__ __forceinline__ bool in_bounds(float u, float v, int W, int H) {
    return u >= 0.f && u < (float)W && v >= 0.f && v < (float)H;
  }
  static __device__ __forceinline__ void store(float* uv, long long i, float u, float v) {
    st_stream_f2(uv + 2 * i, make_float2(u, v));
  }
};
template <> struct OutPix<double> {
  static __device__ __forceinline__ bool in_bounds(double u, double v, int W, int H) {
    return u >= 0.0 && u < (double)W && v >= 0.0 && v < (double)H;
  }
  static __device__ __forceinline__ void store(double* uv, long long i, double u, double v) {
    double2 o = make_double2(u, v);
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(uv + 2 * i), "d"(o.x), "d"(o.y) : "memory");
  }
};

// OpenCV's per-point loop; every product and sum rounded separately.
template <bool DIST>
__device__ __forceinline__ void project_one(const pbk_project_params& P, const double X[3],
                                            double& u, double& v) {
  double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[0], X[0]), __dmul_rn(P.R[1], X[1])),
                                 __dmul_rn(P.R[2], X[2])), P.t[0]);
  double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[3], X[0]), __dmul_rn(P.R[4], X[1])),
                                 __dmul_rn(P.R[5], X[2])), P.t[1]);
  double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.R[6], X[0]), __dmul_rn(P.R[7], X[1])),
                                 __dmul_rn(P.R[8], X[2])), P.t[2]);
  z = (z != 0.0) ? __drcp_rn(z) : 1.0;
  x = __dmul_rn(x, z);
  y = __dmul_rn(y, z);
  // With all-zero coefficients OpenCV still evaluates the distortion
  // polynomial: it is the identity unless r^6 overflows (0*inf = NaN).  |x|,|y|
  // < 1e50 keeps r^6 finite, so only larger magnitudes (or NaN) take the full
  // formula.  The test is on the exponent word: integer pipe, not fp64.
  const bool huge = ((__double2hiint(x) & 0x7fffffff) >= 0x4A511B0E) ||
                    ((__double2hiint(y) & 0x7fffffff) >= 0x4A511B0E);
  if (DIST || huge) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    const double r4 = __dmul_rn(r2, r2);
    const double r6 = __dmul_rn(r4, r2);
    const double a1 = __dmul_rn(__dmul_rn(2.0, x), y);
    const double a2 = __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x));
    const double a3 = __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y));
    const double cdist = __dadd_rn(__dadd_rn(__dadd_rn(1.0, __dmul_rn(P.k[0], r2)), __dmul_rn(P.k[1], r4)),
                                   __dmul_rn(P.k[4], r6));
    const double icdist2 = __drcp_rn(__dadd_rn(
        __dadd_rn(__dadd_rn(1.0, __dmul_rn(P.k[5], r2)), __dmul_rn(P.k[6], r4)), __dmul_rn(P.k[7], r6)));
    const double xd = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(x, cdist), icdist2), __dmul_rn(P.k[2], a1)),
                                __dmul_rn(P.k[3], a2));
    const double yd = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(y, cdist), icdist2), __dmul_rn(P.k[2], a3)),
                                __dmul_rn(P.k[3], a1));
    x = xd;
    y = yd;
  }
  u = __dadd_rn(__dmul_rn(x, P.fx), P.cx);
  v = __dadd_rn(__dmul_rn(y, P.fy), P.cy);
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

// MODE 0: plain projection (no depth, masks, count).
// MODE 1: depth / masks / valid flags per runtime switches, outputs uncompacted.
// MODE 2: as 1 + order-preserving compaction of uv / depth (decoupled look-back
//         over CTA tiles of 1024 points handed out by an atomic ticket).
// scratch layout (MODE 2): [0] ticket counter (u64), [1 .. ntiles] tile status words.
template <typename T, bool DIST, int MODE>
__global__ void __launch_bounds__(kWarps * 32)
project_points_kernel(const T* __restrict__ X, long long n, T* __restrict__ uv,
                      double* __restrict__ depth, uint8_t* __restrict__ valid_out,
                      long long* __restrict__ count, unsigned long long* __restrict__ scratch,
                      const __grid_constant__ pbk_project_params P) {
  constexpr int NST = RingCfg<T>::kStages;
  constexpr int kInBytes = kTilePts * 3 * sizeof(T);
  extern __shared__ __align__(128) char dyn_smem[];
  __shared__ int warp_count[kWarps];
  __shared__ long long s_ticket[2];
  __shared__ long long s_part[kWarps];
  __shared__ int s_has[kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(dyn_smem + kWarps * NST * kInBytes) + warp * NST;
  WarpRing<kInBytes, NST> ring;
  ring.init(dyn_smem + warp * NST * kInBytes, bars, lane);

  const long long nfull = n / kTilePts;                    // full 128-point tiles
  const long long nwt = (n + kTilePts - 1) / kTilePts;     // warp tiles
  const long long ncta = (n + kCtaPts - 1) / kCtaPts;      // CTA tiles (MODE 2)
  const char* Xb = reinterpret_cast<const char*>(X);
  const bool need_depth = MODE != 0 && ((depth != nullptr) || P.check_depth);

  // warp tile index of step k.  MODE 2: CTA tiles are handed out by an atomic
  // ticket, exactly ONE tile ahead of the one being computed (s_ticket[k & 1]):
  // taking several tickets at once would make a CTA sit on consecutive tiles and
  // serialise the look-back chain of every CTA behind it.
  const long long stride = (long long)gridDim.x * kWarps;
  const long long t0 = (long long)blockIdx.x * kWarps + warp;
  auto wtile = [&](int k) -> long long {
    if (MODE == 2) {
      const long long c = s_ticket[k & 1];
      return c >= ncta ? nwt : c * kWarps + warp;
    }
    return t0 + (long long)k * stride;
  };
  if (MODE == 2) {
    if (threadIdx.x == 0) s_ticket[0] = (long long)atomicAdd(scratch, 1ull);
    __syncthreads();
    if (lane == 0) {
      const long long t = wtile(0);
      if (t < nfull) ring.issue(0, Xb + t * kInBytes);
    }
  } else if (lane == 0) {
    for (int k = 0; k < NST - 1; ++k) {
      const long long t = wtile(k);
      if (t < nfull) ring.issue(k, Xb + t * kInBytes);
    }
  }

  long long local_count = 0;
  for (int k = 0;; ++k) {
    long long cta_tile = 0;
    if (MODE == 2) {
      cta_tile = s_ticket[k & 1];
      if (cta_tile >= ncta) break;
      if (threadIdx.x == 0) s_ticket[(k + 1) & 1] = (long long)atomicAdd(scratch, 1ull);
      __syncthreads();                  // (A0) next ticket visible
      if (lane == 0) {
        const long long t = wtile(k + 1);
        if (t < nfull) ring.issue(k + 1, Xb + t * kInBytes);
      }
    }
    const long long tile = wtile(k);
    if (MODE != 2) {
      if (tile >= nwt) break;
      const long long nxt = tile + (NST - 1) * stride;
      if (lane == 0 && nxt < nfull) ring.issue(k + NST - 1, Xb + nxt * kInBytes);
    }
    const long long first = tile * kTilePts;
    const long long rem = n - first;
    const int npts = rem <= 0 ? 0 : (rem < kTilePts ? (int)rem : kTilePts);
    char* st = ring.slot(k);
    if (npts == kTilePts) ring.wait(k);
    else {
      if (npts > 0) warp_load_ragged(Xb + first * 3 * sizeof(T), st, lane, kInBytes, npts * 3 * (int)sizeof(T));
      else warp_load_ragged(Xb, st, lane, kInBytes, 0);
      __syncwarp();
    }
    double Xp[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) stage_read_point<T>(st, j * 32 + lane, Xp[j]);
    __syncwarp();                       // slot may be refilled from here on

    T uo[4], vo[4];
    double dep[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double u, v;
      project_one<DIST>(P, Xp[j], u, v);
      uo[j] = (T)u;
      vo[j] = (T)v;
      if (MODE != 0) {
        bool good = (j * 32 + lane) < npts;
        if (need_depth) {
          double acc = __dmul_rn(P.Rz[0], Xp[j][0]);
          acc = __fma_rn(P.Rz[1], Xp[j][1], acc);
          acc = __fma_rn(P.Rz[2], Xp[j][2], acc);
          acc = __fma_rn(P.tz, 1.0, acc);
          dep[j] = acc;
          if (P.check_depth) good = good && (acc >= P.min_depth);
        }
        if (P.check_bounds) good = good && OutPix<T>::in_bounds(uo[j], vo[j], P.width, P.height);
        ok[j] = good;
      }
    }

    if (MODE == 0) {
      if (npts == kTilePts) {
#pragma unroll
        for (int j = 0; j < 4; ++j) OutPix<T>::store(uv, first + j * 32 + lane, uo[j], vo[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j * 32 + lane < npts) OutPix<T>::store(uv, first + j * 32 + lane, uo[j], vo[j]);
      }
      continue;
    }

    unsigned m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = __ballot_sync(0xffffffffu, ok[j]);
    const int wc = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    if (valid_out != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = j * 32 + lane;
        if (p < npts) valid_out[first + p] = ok[j] ? 1 : 0;
      }
    }

    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = j * 32 + lane;
        if (p < npts) {
          OutPix<T>::store(uv, first + p, uo[j], vo[j]);
          if (depth != nullptr) st_stream_d1(depth + first + p, dep[j]);
        }
      }
      local_count += wc;
      continue;
    }

    // ---- MODE 2: CTA aggregate, decoupled look-back, ordered scatter.
    // The look-back inspects 256 predecessor tiles per round (one per thread):
    // tiles complete at ~250 per microsecond across the chip, so a 32-wide
    // window could not keep up and every CTA ended up waiting at the barrier.
    if (lane == 0) warp_count[warp] = wc;
    __syncthreads();                    // (A)
    int agg = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) agg += warp_count[w];
    unsigned long long* status = scratch + 1;
    if (threadIdx.x == 0)
      st_status(status + cta_tile, (cta_tile == 0 ? kStatePrefix : kStateAgg) | (unsigned long long)agg);
    long long excl = 0;
    if (cta_tile > 0) {
      long long look = cta_tile - 1;
      while (true) {
        const long long idx = look - (long long)threadIdx.x;      // thread 0 = nearest predecessor
        unsigned long long sw = kStatePrefix;                     // virtual prefix 0 before tile 0
        if (idx >= 0) {
          do { sw = ld_status(status + idx); } while ((sw >> 62) == 0);
        }
        const unsigned is_prefix = __ballot_sync(0xffffffffu, (sw >> 62) == 2);
        long long val = (long long)(sw & kValueMask);
        if (is_prefix && lane > __ffs(is_prefix) - 1) val = 0;   // stop at the nearest prefix
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (lane == 0) {
          s_part[warp] = val;
          s_has[warp] = is_prefix != 0;
        }
        __syncthreads();                // (B1)
        bool found = false;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          if (!found) {
            excl += s_part[w];
            found = s_has[w] != 0;
          }
        }
        if (found) break;
        look -= kWarps * 32;
        __syncthreads();                // (B2) s_part reuse
      }
      if (threadIdx.x == 0) st_status(status + cta_tile, kStatePrefix | (unsigned long long)(excl + agg));
    }
    if (threadIdx.x == 0 && cta_tile == ncta - 1 && count != nullptr) *count = excl + agg;
    long long base = excl;
    for (int w = 0; w < warp; ++w) base += warp_count[w];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (ok[j]) {
        const long long o = base + __popc(m[j] & lt);
        OutPix<T>::store(uv, o, uo[j], vo[j]);
        if (depth != nullptr) st_stream_d1(depth + o, dep[j]);
      }
      base += __popc(m[j]);
    }
    __syncthreads();                    // (C) warp_count / s_part reuse
  }
  if (MODE == 1 && count != nullptr) {
    if (lane == 0 && local_count)
      atomicAdd(reinterpret_cast<unsigned long long*>(count), (unsigned long long)local_count);
  }
}

template <typename K>
static int opt_in_smem(K kernel, int bytes) {
  if (bytes > 48 * 1024) PBK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return PBK_OK;
}

template <typename TI, typename TO>
static int launch_transform(const void* X, int64_t n, const XformParams& P, void* Y,
                            cudaStream_t s, const DeviceInfo* di) {
  constexpr int NST = RingCfg<TI>::kStages;
  const int smem = kWarps * (NST * kTilePts * 3 * (int)sizeof(TI) + kTilePts * 3 * (int)sizeof(TO)) + kWarps * NST * 8;
  auto kern = transform_points_kernel<TI, TO>;
  int rc = opt_in_smem(kern, smem);
  if (rc) return rc;
  const long long ntiles = (n + kTilePts - 1) / kTilePts;
  const int grid = resident_grid(di, kern, kWarps * 32, smem, (ntiles + kWarps - 1) / kWarps);
  kern<<<grid, kWarps * 32, smem, s>>>(static_cast<const TI*>(X), n, static_cast<TO*>(Y), P);
  PBK_CHECK_LAUNCH("transform_points_kernel");
  return PBK_OK;
}

template <typename T, bool DIST, int MODE>
static int launch_project(const void* X, int64_t n, const pbk_project_params& P, void* uv,
                          double* depth, uint8_t* valid, int64_t* count, void* scratch,
                          cudaStream_t s, const DeviceInfo* di) {
  constexpr int NST = RingCfg<T>::kStages;
  const int smem = kWarps * NST * kTilePts * 3 * (int)sizeof(T) + kWarps * NST * 8;
  auto kern = project_points_kernel<T, DIST, MODE>;
  int rc = opt_in_smem(kern, smem);
  if (rc) return rc;
  const long long ncta = (n + kCtaPts - 1) / kCtaPts;
  const int grid = resident_grid(di, kern, kWarps * 32, smem, ncta);
  kern<<<grid, kWarps * 32, smem, s>>>(static_cast<const T*>(X), n, static_cast<T*>(uv), depth, valid,
                                       reinterpret_cast<long long*>(count),
                                       static_cast<unsigned long long*>(scratch), P);
  PBK_CHECK_LAUNCH("project_points_kernel");
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

size_t pbk_project_scratch_bytes(int64_t n) {
  if (n < 0) n = 0;
  const long long ntiles = (n + pbk::kCtaPts - 1) / pbk::kCtaPts;
  return (size_t)(ntiles + 2) * 8;
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
  if (P.compact) PBK_CUDA(cudaMemsetAsync(scratch, 0, pbk_project_scratch_bytes(n), s));
  else if (count != nullptr) {
    if (plain) {
      const int64_t all = n;            // pageable source: staged before the call returns
      PBK_CUDA(cudaMemcpyAsync(count, &all, 8, cudaMemcpyHostToDevice, s));
    } else {
      PBK_CUDA(cudaMemsetAsync(count, 0, 8, s));
    }
  }
  const bool dist = P.has_distortion != 0;
  const int mode = plain ? 0 : (P.compact ? 2 : 1);
#define PBK_DISPATCH(T)                                                                              \
  switch (mode * 2 + (dist ? 1 : 0)) {                                                               \
    case 0: return launch_project<T, false, 0>(X, n, P, uv, depth, valid, count, scratch, s, di);    \
    case 1: return launch_project<T, true, 0>(X, n, P, uv, depth, valid, count, scratch, s, di);     \
    case 2: return launch_project<T, false, 1>(X, n, P, uv, depth, valid, count, scratch, s, di);    \
    case 3: return launch_project<T, true, 1>(X, n, P, uv, depth, valid, count, scratch, s, di);     \
    case 4: return launch_project<T, false, 2>(X, n, P, uv, depth, valid, count, scratch, s, di);    \
    default: return launch_project<T, true, 2>(X, n, P, uv, depth, valid, count, scratch, s, di);    \
  }
  if (in_dtype == PBK_F32) { PBK_DISPATCH(float) }
  PBK_DISPATCH(double)
#undef PBK_DISPATCH
}

}  // extern "C"
