// "Next" rows of the hot path: the callers either side of
// Camera.project / StereoCamera.reconstruct.
//
//  * DepthCamera.reconstruct / reconstruct_sparse   pybot/vision/camera_utils.py:1010-1041
//      depth image -> organised cloud dstack([xs*d, ys*d, d]) in float64, with the
//      reference's own precomputed xs / ys meshes (_build_mesh :1020-1030).
//  * SceneFlow.process                               pybot/vision/optflow_utils.py:114-150
//      project H*W scene points, cast to int32, bounds mask, scatter the source
//      grid coordinate to the hit pixel (last writer wins), subtract the grid.
//      The reference does this as a cv2.projectPoints call plus five numpy
//      passes over H*W points; here it is one projection+scatter kernel and one
//      resolve kernel, and only the (H,W,2) flow leaves the device.
//  * sampson_error / filter_sampson_error            pybot/vision/camera_utils.py:52-90
//      epipolar filter between matching and triangulation; diagonal only.
//  * triangulate_points                              pybot/vision/camera_utils.py:108-116
//      cv2.triangulatePoints (DLT + 4x4 SVD per point) + dehomogenisation.
//  * CameraIntrinsic.undistort_points / ray / reconstruct  pybot/vision/camera_utils.py:499-538
//      cv2.undistortPoints (5 fixed-point iterations, fp64) + ray / rotate / normalise / *Z.
//  * viewer wire colours                             pybot/externals/draw_helpers.py:35-59,
//      float32 RGB in [0,1] exactly as arr_msg ships them  pybot/externals/pb.py:18-38
#include "pbk_common.cuh"
#include "project_math.cuh"

namespace pbk {

// ------------------------------------------------------ DepthCamera.reconstruct
template <typename T> __device__ __forceinline__ double depth_as_double(T v) { return (double)v; }

// four consecutive depth samples starting at a 4-aligned element index, one vector load
template <typename T> struct Depth4;
template <> struct Depth4<float> {
  static __device__ __forceinline__ void load(const float* p, double d[4]) {
    const float4 v = ld_stream_f4(p);
    d[0] = (double)v.x; d[1] = (double)v.y; d[2] = (double)v.z; d[3] = (double)v.w;
  }
};
template <> struct Depth4<uint16_t> {
  static __device__ __forceinline__ void load(const uint16_t* p, double d[4]) {
    const uint2 v = ld_stream_u2(p);
    d[0] = (double)(v.x & 0xffffu); d[1] = (double)(v.x >> 16);
    d[2] = (double)(v.y & 0xffffu); d[3] = (double)(v.y >> 16);
  }
};
template <> struct Depth4<double> {
  static __device__ __forceinline__ void load(const double* p, double d[4]) {
    const uint4 a = ld_stream_u4(p), b = ld_stream_u4(p + 2);
    d[0] = __hiloint2double((int)a.y, (int)a.x); d[1] = __hiloint2double((int)a.w, (int)a.z);
    d[2] = __hiloint2double((int)b.y, (int)b.x); d[3] = __hiloint2double((int)b.w, (int)b.z);
  }
};

constexpr int kDcWarps = 8;
constexpr int kDcTile = 128;                 // output pixels per warp tile (4 per lane)

__device__ __forceinline__ void st_stream_d2(void* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// skip == 1: flat stream over batch*H*W pixels.  One warp owns 128 consecutive
// pixels; lane l computes pixels 4l .. 4l+3 (one 16 B load of float32 depth, 96 B
// of output) and writes them to a warp-private stage that leaves as ONE TMA bulk
// store of 3 KB (cp.async.bulk shared -> global): no LDS / STG for the outputs.
// Two stages alternate so a tile can be written while the previous one drains.
// 4 B in + 24 B out per pixel (float32 depth): HBM-write-bound.
constexpr int kDcPpl = 4;                    // pixels per lane
template <typename T>
__global__ void __launch_bounds__(kDcWarps * 32)
depth_reconstruct_dense_kernel(const T* __restrict__ depth, long long total, int W, int H,
                               const double* __restrict__ xs, const double* __restrict__ ys,
                               double* __restrict__ out) {
  __shared__ __align__(128) double stage[kDcWarps][2][kDcTile * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ntiles = total / kDcTile;
  const long long stride = (long long)gridDim.x * kDcWarps;
  int it = 0;
  for (long long tile = (long long)blockIdx.x * kDcWarps + warp; tile < ntiles; tile += stride, ++it) {
    const long long p0 = tile * kDcTile + kDcPpl * lane;
    double d[kDcPpl];
    Depth4<T>::load(depth + p0, d);
    const long long row = p0 / W;
    int x = (int)(p0 - row * W);
    int y = (int)(row % H);
    double* st = stage[warp][it & 1];
    // the bulk store that used this stage two tiles ago must have drained it
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    double o[kDcPpl * 3];
#pragma unroll
    for (int k = 0; k < kDcPpl; ++k) {
      o[3 * k] = __dmul_rn(__ldg(xs + x), d[k]);
      o[3 * k + 1] = __dmul_rn(__ldg(ys + y), d[k]);
      o[3 * k + 2] = d[k];
      if (++x == W) { x = 0; if (++y == H) y = 0; }
    }
    double2* s2 = reinterpret_cast<double2*>(st) + lane * (kDcPpl * 3 / 2);
#pragma unroll
    for (int j = 0; j < kDcPpl * 3 / 2; ++j) s2[j] = make_double2(o[2 * j], o[2 * j + 1]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_1d(out + tile * (kDcTile * 3), st, kDcTile * 24);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait_all();
}

// One output pixel per thread: skip > 1 and the < 128-pixel tail of the dense kernel.
template <typename T>
__global__ void __launch_bounds__(256)
depth_reconstruct_generic_kernel(const T* __restrict__ depth, int W, int H, int skip, int Ws, int Hs,
                                 const double* __restrict__ xs, const double* __restrict__ ys,
                                 double* __restrict__ out, long long first, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long o = first + i;
  const long long frame = o / ((long long)Ws * Hs);
  const int rem = (int)(o - frame * (long long)Ws * Hs);
  const int yo = rem / Ws, xo = rem - yo * Ws;
  const double d = depth_as_double(depth[(frame * H + (long long)yo * skip) * W + (long long)xo * skip]);
  out[o * 3 + 0] = __dmul_rn(xs[xo], d);
  out[o * 3 + 1] = __dmul_rn(ys[yo], d);
  out[o * 3 + 2] = d;
}

// reconstruct_sparse (:1038-1041): note the reference divides BOTH axes by fx.
__global__ void __launch_bounds__(256)
depth_reconstruct_sparse_kernel(const double* __restrict__ pts, const double* __restrict__ depth,
                                long long n, double fx, double cx, double cy, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = depth[i];
  out[3 * i + 0] = __dmul_rn(__ddiv_rn(__dmul_rn(__dsub_rn(pts[2 * i], cx), 1.0), fx), d);
  out[3 * i + 1] = __dmul_rn(__ddiv_rn(__dmul_rn(__dsub_rn(pts[2 * i + 1], cy), 1.0), fx), d);
  out[3 * i + 2] = d;
}

template <typename T>
static int launch_depth(const void* depth_v, int batch, int H, int W, int skip, const double* xs,
                        const double* ys, double* out, cudaStream_t s, const DeviceInfo* di) {
  const T* depth = static_cast<const T*>(depth_v);
  const int Ws = (W + skip - 1) / skip, Hs = (H + skip - 1) / skip;
  if (skip == 1) {
    const long long total = (long long)batch * H * W;
    const long long ntiles = total / kDcTile;
    if (ntiles > 0) {
      auto kern = depth_reconstruct_dense_kernel<T>;
      static KernelCache kc;
      int per_sm;
      const int rc = kc.setup(di, kern, kDcWarps * 32, 0, &per_sm);
      if (rc) return rc;
      const int grid = resident_grid(di, per_sm, (ntiles + kDcWarps - 1) / kDcWarps);
      kern<<<grid, kDcWarps * 32, 0, s>>>(depth, total, W, H, xs, ys, out);
      PBK_CHECK_LAUNCH("depth_reconstruct_dense_kernel");
    }
    const long long done = ntiles * kDcTile;
    if (done < total) {
      depth_reconstruct_generic_kernel<T><<<1, 256, 0, s>>>(depth, W, H, 1, W, H, xs, ys, out, done, total - done);
      PBK_CHECK_LAUNCH("depth_reconstruct_generic_kernel(tail)");
    }
  } else {
    const long long outs = (long long)batch * Ws * Hs;
    const long long blocks = (outs + 255) / 256;
    PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "too many output pixels");
    depth_reconstruct_generic_kernel<T><<<(int)blocks, 256, 0, s>>>(depth, W, H, skip, Ws, Hs, xs, ys, out, 0, outs);
    PBK_CHECK_LAUNCH("depth_reconstruct_generic_kernel");
  }
  return PBK_OK;
}

// ------------------------------------------------------------ SceneFlow.process
// numpy's .astype(np.int32) on x86: truncation toward zero; NaN and anything
// outside int32 become INT_MIN (cvttss2si / cvttsd2si).
template <typename T>
__device__ __forceinline__ int cast_i32_like_numpy(T v) {
  const double d = (double)v;
  if (!(fabs(d) < 2147483648.0)) return (int)0x80000000;
  return (int)d;                       // cvt.rzi
}

// Four consecutive points per thread: 48 B (float32) arrive as three 16 B loads
// (the per-point version issued twelve 4 B loads at a 12 B stride and fetched every
// sector three times through L1), then project_many interleaves the four fp64 chains.
template <typename T> struct Pts4;
template <> struct Pts4<float> {
  static __device__ __forceinline__ void load(const float* p, double (&X)[4][3]) {
    const float4 a = ld_stream_f4(p), b = ld_stream_f4(p + 4), c = ld_stream_f4(p + 8);
    const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 12; ++k) X[k / 3][k % 3] = (double)v[k];
  }
};
template <> struct Pts4<double> {
  static __device__ __forceinline__ void load(const double* p, double (&X)[4][3]) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const uint4 a = ld_stream_u4(p + 2 * k);
      X[(2 * k) / 3][(2 * k) % 3] = __hiloint2double((int)a.y, (int)a.x);
      X[(2 * k + 1) / 3][(2 * k + 1) % 3] = __hiloint2double((int)a.w, (int)a.z);
    }
  }
};

template <typename T, bool DIST>
__global__ void __launch_bounds__(256)
sceneflow_scatter_kernel(const T* __restrict__ dX, int n, int W, int H,
                         const __grid_constant__ pbk_project_params P, int* __restrict__ winner) {
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  double X[4][3], u[4], v[4];
  const int cnt = n - i0 < 4 ? n - i0 : 4;
  if (cnt == 4) {
    Pts4<T>::load(dX + 3ll * i0, X);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k) X[j][k] = j < cnt ? (double)dX[3ll * (i0 + j) + k] : 1.0;
  }
  project_many<DIST, 4>(P, X, u, v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int xi = cast_i32_like_numpy<T>((T)u[j]), yi = cast_i32_like_numpy<T>((T)v[j]);
    // new_grid[ys[valid], xs[valid]] = grid[gys[valid], gxs[valid]]: on repeated
    // targets numpy keeps the LAST source in index order -> max source index
    if (j < cnt && xi >= 0 && xi < W && yi >= 0 && yi < H) atomicMax(winner + yi * W + xi, i0 + j);
  }
}

// n / W for 0 <= n < 2^31 by multiply-high (Granlund-Montgomery round-up variant):
// q = (umulhi(m, n) + n) >> s with s = ceil(log2 W), m = floor(2^32 (2^s - W) / W) + 1.
struct FastDiv {
  uint32_t m;
  int s;
};
__device__ __forceinline__ int fast_div(int n, FastDiv d) {
  return (int)((__umulhi(d.m, (uint32_t)n) + (uint32_t)n) >> d.s);
}

// four consecutive cells per thread: one 16 B load of winners, two 16 B stores of flow
__global__ void __launch_bounds__(256)
sceneflow_resolve_kernel(const int* __restrict__ winner, int n, int W, FastDiv dW, float* __restrict__ flow) {
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= n) return;
  int w[4];
  const int cnt = n - c0 < 4 ? n - c0 : 4;
  if (cnt == 4) {
    const uint4 v = ld_stream_u4(winner + c0);
    w[0] = (int)v.x; w[1] = (int)v.y; w[2] = (int)v.z; w[3] = (int)v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = j < cnt ? winner[c0 + j] : -1;
  }
  int cy = fast_div(c0, dW), cx = c0 - cy * W;
  float o[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (w[j] < 0) {
      o[2 * j] = o[2 * j + 1] = __int_as_float(0x7fc00000);          // untouched cells stay NaN
    } else {
      // (source grid coordinate) - (this cell's grid coordinate), float32 like the reference
      const int wy = fast_div(w[j], dW), wx = w[j] - wy * W;
      o[2 * j] = (float)wx - (float)cx;
      o[2 * j + 1] = (float)wy - (float)cy;
    }
    if (++cx == W) { cx = 0; ++cy; }
  }
  if (cnt == 4) {
    st_stream_f4(flow + 2ll * c0, make_float4(o[0], o[1], o[2], o[3]));
    st_stream_f4(flow + 2ll * c0 + 4, make_float4(o[4], o[5], o[6], o[7]));
  } else {
    for (int j = 0; j < cnt; ++j) { flow[2ll * (c0 + j)] = o[2 * j]; flow[2ll * (c0 + j) + 1] = o[2 * j + 1]; }
  }
}

struct SampsonParams {
  double F[9];
};

// ---------------------------------------------------------------- sampson_error
// camera_utils.py:52-68.  The reference forms the full N x N product
// x1^T (F x2) and keeps its diagonal; only the diagonal is computed here.
// Products follow the k = 0,1,2 FMA chain of the 3-term dot products numpy's
// BLAS runs (see transform_points_kernel).
__global__ void __launch_bounds__(256)
sampson_error_kernel(const double* __restrict__ p1, const double* __restrict__ p2, long long n,
                     const __grid_constant__ SampsonParams P, double* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x1[3] = {p1[2 * i], p1[2 * i + 1], 1.0}, x2[3] = {p2[2 * i], p2[2 * i + 1], 1.0};
  double a[3], b[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    a[r] = __fma_rn(P.F[3 * r + 2], x1[2], __fma_rn(P.F[3 * r + 1], x1[1], __dmul_rn(P.F[3 * r], x1[0])));
    b[r] = __fma_rn(P.F[3 * r + 2], x2[2], __fma_rn(P.F[3 * r + 1], x2[1], __dmul_rn(P.F[3 * r], x2[0])));
  }
  const double denom = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a[0], a[0]), __dmul_rn(a[1], a[1])),
                                           __dmul_rn(b[0], b[0])), __dmul_rn(b[1], b[1]));
  const double d = __fma_rn(x1[2], b[2], __fma_rn(x1[1], b[1], __dmul_rn(x1[0], b[0])));
  err[i] = __ddiv_rn(__dmul_rn(d, d), denom);
}

// ------------------------------------------------------------- epipolar_line
// camera_utils.py:352-357: cv2.computeCorrespondEpilines(x.reshape(-1,1,2), 1, F).reshape(-1,3).
// OpenCV evaluates l = F [x y 1]^T in double with separately rounded products and sums
// (left to right), scales by 1/sqrt(a^2 + b^2) (1 when that is zero) and rounds once to the
// point dtype; the same operation order here makes the lines bit-identical.
template <typename T>
__global__ void __launch_bounds__(256)
epipolar_lines_kernel(const T* __restrict__ pts, long long n, const __grid_constant__ SampsonParams P,
                      T* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = (double)pts[2 * i], y = (double)pts[2 * i + 1];
  double l[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    l[r] = __dadd_rn(__dadd_rn(__dmul_rn(P.F[3 * r], x), __dmul_rn(P.F[3 * r + 1], y)), P.F[3 * r + 2]);
  double nu = __dadd_rn(__dmul_rn(l[0], l[0]), __dmul_rn(l[1], l[1]));
  nu = (nu != 0.0) ? __ddiv_rn(1.0, __dsqrt_rn(nu)) : 1.0;
#pragma unroll
  for (int r = 0; r < 3; ++r) out[3 * i + r] = (T)__dmul_rn(l[r], nu);
}

// -------------------------------------------------------- triangulate_points
// camera_utils.py:108-116: X = cv2.triangulatePoints(P1, P2, x1, x2).T ; X[:, :3] / X[:, 3].
// Per point the 4x4 DLT matrix A = [x1*P1[2]-P1[0]; y1*P1[2]-P1[1]; x2*P2[2]-P2[0];
// y2*P2[2]-P2[1]] and its right singular vector of the smallest singular value, by
// the one-sided (Hestenes) Jacobi SVD OpenCV uses for small matrices; everything
// stays in registers (fully unrolled column pairs).  Agrees with cv2 to ~1e-14
// relative; parity is a tolerance (an SVD has no bit-exact contract).
struct TriParams {
  double P1[12], P2[12];
};

template <typename T>
__global__ void __launch_bounds__(128)
triangulate_kernel(const T* __restrict__ x1, const T* __restrict__ x2, long long n,
                   const __grid_constant__ TriParams P, T* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const double u1 = (double)x1[2 * idx], v1 = (double)x1[2 * idx + 1];
  const double u2 = (double)x2[2 * idx], v2 = (double)x2[2 * idx + 1];
  double U[4][4], V[4][4];                      // U[row][col]: columns are rotated
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    U[0][c] = u1 * P.P1[8 + c] - P.P1[c];
    U[1][c] = v1 * P.P1[8 + c] - P.P1[4 + c];
    U[2][c] = u2 * P.P2[8 + c] - P.P2[c];
    U[3][c] = v2 * P.P2[8 + c] - P.P2[4 + c];
#pragma unroll
    for (int r = 0; r < 4; ++r) V[r][c] = (r == c) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        double a = 0.0, b = 0.0, p = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          a = fma(U[r][i], U[r][i], a);
          b = fma(U[r][j], U[r][j], b);
          p = fma(U[r][i], U[r][j], p);
        }
        if (fabs(p) > 2.220446049250313e-16 * sqrt(a * b) && fabs(p) > 1e-300) {
          const double p2 = p * 2.0, beta = a - b, gamma = hypot(p2, beta);
          double c, s;
          if (beta < 0.0) {
            const double delta = (gamma - beta) * 0.5;
            s = sqrt(delta / gamma);
            c = p / (gamma * s);
          } else {
            c = sqrt((gamma + beta) / (gamma * 2.0));
            s = p / (gamma * c);
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const double ui = U[r][i], uj = U[r][j], vi = V[r][i], vj = V[r][j];
            U[r][i] = fma(c, ui, s * uj);
            U[r][j] = fma(-s, ui, c * uj);
            V[r][i] = fma(c, vi, s * vj);
            V[r][j] = fma(-s, vi, c * vj);
          }
          changed = true;
        }
      }
    }
    if (!changed) break;
  }
  // column with the smallest norm = smallest singular value
  double best = 0.0, X[4] = {0.0, 0.0, 0.0, 1.0};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double nrm = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) nrm = fma(U[r][c], U[r][c], nrm);
    if (c == 0 || nrm < best) {
      best = nrm;
#pragma unroll
      for (int r = 0; r < 4; ++r) X[r] = V[r][c];
    }
  }
  // cv2 returns the homogeneous point in the input depth; the reference divides in that dtype
  const T w = (T)X[3];
  out[3 * idx + 0] = (T)X[0] / w;
  out[3 * idx + 1] = (T)X[1] / w;
  out[3 * idx + 2] = (T)X[2] / w;
}

// ------------------------------------------- undistort_points / ray / reconstruct
// camera_utils.py:499-538.  cv2.undistortPoints with the default criteria runs
// exactly five iterations of x <- (x0 - delta(x)) / cdist(x) in fp64 and rounds
// once to float32; every product and sum below is rounded separately in OpenCV's
// order (oracle/model.py undistort_points), so the float32 pixels are bit-exact.
__device__ __forceinline__ void undistort_one(const pbk_ray_params& P, double u, double v, float& ou, float& ov) {
  const double* k = P.k;
  const double ifx = __ddiv_rn(1.0, P.fx), ify = __ddiv_rn(1.0, P.fy);
  double x = __dmul_rn(__dsub_rn(u, P.cx), ifx), y = __dmul_rn(__dsub_rn(v, P.cy), ify);
  const double x0 = x, y0 = y;
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
    const double icdist = __ddiv_rn(num, den);
    if (icdist < 0.0) {                      // OpenCV: give up, keep the pinhole estimate
      x = x0;
      y = y0;
      break;
    }
    const double dx = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y),
                                __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
    const double dy = __dadd_rn(__dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y));
    x = __dmul_rn(__dsub_rn(x0, dx), icdist);
    y = __dmul_rn(__dsub_rn(y0, dy), icdist);
  }
  const double xx = __dadd_rn(__dadd_rn(__dmul_rn(P.P[0], x), __dmul_rn(P.P[1], y)), P.P[2]);
  const double yy = __dadd_rn(__dadd_rn(__dmul_rn(P.P[3], x), __dmul_rn(P.P[4], y)), P.P[5]);
  const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(P.P[6], x), __dmul_rn(P.P[7], y)), P.P[8]));
  ou = (float)__dmul_rn(xx, ww);
  ov = (float)__dmul_rn(yy, ww);
}

template <typename T>
__global__ void __launch_bounds__(256)
undistort_points_kernel(const T* __restrict__ pts, int stride, long long n,
                        const __grid_constant__ pbk_ray_params P, float2* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // pts.astype(np.float32) first (:537)
  const float u = (float)pts[i * stride], v = (float)pts[i * stride + 1];
  float2 o;
  undistort_one(P, (double)u, (double)v, o.x, o.y);
  out[i] = o;
}

// One point per thread; the (256,3) float64 block leaves through shared memory as
// contiguous 16 B stores (a lane-per-point store would be 24 B strided).
template <typename T>
__global__ void __launch_bounds__(256)
camera_rays_kernel(const T* __restrict__ pts, int stride, long long n,
                   const __grid_constant__ pbk_ray_params P, double* __restrict__ out) {
  __shared__ __align__(16) double stage[256 * 3];
  const long long base = (long long)blockIdx.x * 256;
  const long long i = base + threadIdx.x;
  if (i < n) {
    double u, v;
    if (P.undistort) {
      float fu, fv;
      undistort_one(P, (double)(float)pts[i * stride], (double)(float)pts[i * stride + 1], fu, fv);
      u = (double)fu;
      v = (double)fv;
    } else {
      u = (double)pts[i * stride];
      v = (double)pts[i * stride + 1];
    }
    double r[3] = {__ddiv_rn(__dsub_rn(u, P.cx), P.fx), __ddiv_rn(__dsub_rn(v, P.cy), P.fy), 1.0};
    if (P.rotate) {                                    // Quaternion.rotate, quaternion.py:69-85
      double q[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        q[a] = __dadd_rn(__dmul_rn(2.0, __dadd_rn(__dadd_rn(__dmul_rn(P.rot[3 * a], r[0]), __dmul_rn(P.rot[3 * a + 1], r[1])),
                                                  __dmul_rn(P.rot[3 * a + 2], r[2]))), r[a]);
      r[0] = q[0]; r[1] = q[1]; r[2] = q[2];
    }
    if (P.normalize) {                                 // np.linalg.norm(axis=1): sqrt((a+b)+c)
      const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1])), __dmul_rn(r[2], r[2])));
      r[0] = __ddiv_rn(r[0], nrm); r[1] = __ddiv_rn(r[1], nrm); r[2] = __ddiv_rn(r[2], nrm);
    }
    if (P.scale_by_z) {
      const double z = (double)pts[i * stride + 2];
      r[0] = __dmul_rn(r[0], z); r[1] = __dmul_rn(r[1], z); r[2] = __dmul_rn(r[2], z);
    }
    stage[3 * threadIdx.x] = r[0];
    stage[3 * threadIdx.x + 1] = r[1];
    stage[3 * threadIdx.x + 2] = r[2];
  }
  __syncthreads();
  const long long left = n - base;
  const int cnt = (int)(left < 256 ? left : 256) * 3;      // doubles in this block; base*24 B is 16 B aligned
  double* dst = out + base * 3;
  for (int e = 2 * threadIdx.x; e + 1 < cnt; e += 512) st_stream_d2(dst + e, stage[e], stage[e + 1]);
  if ((cnt & 1) && threadIdx.x == 0) st_stream_d1(dst + cnt - 1, stage[cnt - 1]);
}

// ------------------------------------------------------- viewer wire colours
// get_color_arr (externals/draw_helpers.py:35-53) for uint8 images: optional
// B<->R flip, then carr.astype(float32) / 255.0 - the float32 (N,3) block
// arr_msg (externals/pb.py:35) sends as msg.colors.  4 pixels (12 B in, 48 B out)
// per thread.
__global__ void __launch_bounds__(256)
wire_colors_kernel(const uint8_t* __restrict__ bgr, long long n_px, int flip_rb, float* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // group of 4 pixels
  const long long p0 = q * 4;
  if (p0 >= n_px) return;
  if (p0 + 4 <= n_px) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(bgr + p0 * 3);
    const uint32_t w0 = ld_stream_u1(src), w1 = ld_stream_u1(src + 1), w2 = ld_stream_u1(src + 2);
    uint8_t c[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) { c[k] = (w0 >> (8 * k)) & 0xff; c[4 + k] = (w1 >> (8 * k)) & 0xff; c[8 + k] = (w2 >> (8 * k)) & 0xff; }
    float f[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t c0 = c[3 * k], c1 = c[3 * k + 1], c2 = c[3 * k + 2];
      f[3 * k] = __fdiv_rn((float)(flip_rb ? c2 : c0), 255.0f);
      f[3 * k + 1] = __fdiv_rn((float)c1, 255.0f);
      f[3 * k + 2] = __fdiv_rn((float)(flip_rb ? c0 : c2), 255.0f);
    }
    float4* dst = reinterpret_cast<float4*>(out + p0 * 3);
    st_stream_f4(dst, make_float4(f[0], f[1], f[2], f[3]));
    st_stream_f4(dst + 1, make_float4(f[4], f[5], f[6], f[7]));
    st_stream_f4(dst + 2, make_float4(f[8], f[9], f[10], f[11]));
  } else {
    for (long long p = p0; p < n_px; ++p) {
      const uint8_t c0 = bgr[3 * p], c1 = bgr[3 * p + 1], c2 = bgr[3 * p + 2];
      out[3 * p] = __fdiv_rn((float)(flip_rb ? c2 : c0), 255.0f);
      out[3 * p + 1] = __fdiv_rn((float)c1, 255.0f);
      out[3 * p + 2] = __fdiv_rn((float)(flip_rb ? c0 : c2), 255.0f);
    }
  }
}

// ------------------------------------------------- check_visibility / get_median_depth
// camera_utils.py:245-266: pts_c = camera.c2w(pts) (np.dot(T, [X;1]): the mul,fma,fma,fma chain
// of transform_points_kernel), v = pts_c / np.linalg.norm(pts_c, axis=1) (sqrt((x^2+y^2)+z^2),
// every operation rounded), hangle = arctan2(v_x, v_z), vangle = arctan2(-v_y, v_z), and the
// four comparisons - fused: the transformed points never leave the SM, one byte per point
// comes back instead of 24 B out + a host pass over them.
struct FovParams {
  double R[9], t[3];
  double half_h, half_v, zmin, zmax;
};
template <typename T>
__global__ void __launch_bounds__(256)
fov_mask_kernel(const T* __restrict__ X, long long n, const __grid_constant__ FovParams P, uint8_t* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = (double)X[3 * i], y = (double)X[3 * i + 1], z = (double)X[3 * i + 2];
  double c[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double acc = __dmul_rn(P.R[3 * r], x);
    acc = __fma_rn(P.R[3 * r + 1], y, acc);
    acc = __fma_rn(P.R[3 * r + 2], z, acc);
    c[r] = __fma_rn(P.t[r], 1.0, acc);
  }
  const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(c[0], c[0]), __dmul_rn(c[1], c[1])), __dmul_rn(c[2], c[2])));
  const double vx = __ddiv_rn(c[0], nrm), vy = __ddiv_rn(c[1], nrm), vz = __ddiv_rn(c[2], nrm);
  const double ha = atan2(vx, vz), va = atan2(-vy, vz);
  mask[i] = (fabs(ha) < P.half_h) && (fabs(va) < P.half_v) && (c[2] >= P.zmin) && (c[2] <= P.zmax);
}

// z of T * X for every stride-th point (get_median_depth, camera_utils.py:269-276:
// (camera * pts[::subsample])[:, 2]): 8 B per sampled point out instead of 24
template <typename T>
__global__ void __launch_bounds__(256)
transform_depth_kernel(const T* __restrict__ X, long long n_out, long long stride, double r0, double r1, double r2,
                       double tz, double* __restrict__ z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const T* p = X + 3 * i * stride;
  double acc = __dmul_rn(r0, (double)p[0]);
  acc = __fma_rn(r1, (double)p[1], acc);
  acc = __fma_rn(r2, (double)p[2], acc);
  z[i] = __fma_rn(tz, 1.0, acc);
}

}  // namespace pbk

extern "C" {

int pbk_fov_mask(const void* X, int in_dtype, int64_t n, const double* Rt, double half_fov_h, double half_fov_v,
                 double zmin, double zmax, uint8_t* mask, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && Rt != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(in_dtype == PBK_F32 || in_dtype == PBK_F64, PBK_EINVAL, "points must be float32 or float64");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(X && mask, PBK_EINVAL, "NULL device pointer");
  FovParams P;
  for (int i = 0; i < 9; ++i) P.R[i] = Rt[i];
  for (int i = 0; i < 3; ++i) P.t[i] = Rt[9 + i];
  P.half_h = half_fov_h; P.half_v = half_fov_v; P.zmin = zmin; P.zmax = zmax;
  const long long blocks = (n + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  cudaStream_t s = as_stream(stream);
  if (in_dtype == PBK_F32) fov_mask_kernel<float><<<(int)blocks, 256, 0, s>>>(static_cast<const float*>(X), n, P, mask);
  else fov_mask_kernel<double><<<(int)blocks, 256, 0, s>>>(static_cast<const double*>(X), n, P, mask);
  PBK_CHECK_LAUNCH("fov_mask_kernel");
  return PBK_OK;
}

int pbk_transform_depth(const void* X, int in_dtype, int64_t n, int64_t stride, const double* Rz_tz, double* z,
                        pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && stride >= 1 && Rz_tz != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(in_dtype == PBK_F32 || in_dtype == PBK_F64, PBK_EINVAL, "points must be float32 or float64");
  const long long n_out = (n + stride - 1) / stride;
  if (n_out == 0) return PBK_OK;
  PBK_REQUIRE(X && z, PBK_EINVAL, "NULL device pointer");
  const long long blocks = (n_out + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  cudaStream_t s = as_stream(stream);
  if (in_dtype == PBK_F32)
    transform_depth_kernel<float><<<(int)blocks, 256, 0, s>>>(static_cast<const float*>(X), n_out, stride, Rz_tz[0],
                                                              Rz_tz[1], Rz_tz[2], Rz_tz[3], z);
  else
    transform_depth_kernel<double><<<(int)blocks, 256, 0, s>>>(static_cast<const double*>(X), n_out, stride, Rz_tz[0],
                                                               Rz_tz[1], Rz_tz[2], Rz_tz[3], z);
  PBK_CHECK_LAUNCH("transform_depth_kernel");
  return PBK_OK;
}

int pbk_sampson_error(const double* F, const double* pts1, const double* pts2, int64_t n, double* err,
                      pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && F != nullptr, PBK_EINVAL, "bad arguments");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(pts1 && pts2 && err, PBK_EINVAL, "NULL device pointer");
  SampsonParams P;
  for (int i = 0; i < 9; ++i) P.F[i] = F[i];
  const long long blocks = (n + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  sampson_error_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(pts1, pts2, n, P, err);
  PBK_CHECK_LAUNCH("sampson_error_kernel");
  return PBK_OK;
}

int pbk_epipolar_lines(const double* F, const void* pts, int dtype, int64_t n, void* lines, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && F != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(dtype == PBK_F32 || dtype == PBK_F64, PBK_EINVAL, "points must be float32 or float64");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(pts && lines, PBK_EINVAL, "NULL device pointer");
  SampsonParams P;
  for (int i = 0; i < 9; ++i) P.F[i] = F[i];
  const long long blocks = (n + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  cudaStream_t s = as_stream(stream);
  if (dtype == PBK_F32)
    epipolar_lines_kernel<float><<<(int)blocks, 256, 0, s>>>(static_cast<const float*>(pts), n, P, static_cast<float*>(lines));
  else
    epipolar_lines_kernel<double><<<(int)blocks, 256, 0, s>>>(static_cast<const double*>(pts), n, P, static_cast<double*>(lines));
  PBK_CHECK_LAUNCH("epipolar_lines_kernel");
  return PBK_OK;
}

int pbk_triangulate_points(const double* P1, const double* P2, const void* x1, const void* x2, int dtype,
                           int64_t n, void* out, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && P1 != nullptr && P2 != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(dtype == PBK_F32 || dtype == PBK_F64, PBK_EINVAL, "points must be float32 or float64");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(x1 && x2 && out, PBK_EINVAL, "NULL device pointer");
  TriParams P;
  for (int i = 0; i < 12; ++i) { P.P1[i] = P1[i]; P.P2[i] = P2[i]; }
  const long long blocks = (n + 127) / 128;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  cudaStream_t s = as_stream(stream);
  if (dtype == PBK_F32)
    triangulate_kernel<float><<<(int)blocks, 128, 0, s>>>(static_cast<const float*>(x1), static_cast<const float*>(x2), n, P, static_cast<float*>(out));
  else
    triangulate_kernel<double><<<(int)blocks, 128, 0, s>>>(static_cast<const double*>(x1), static_cast<const double*>(x2), n, P, static_cast<double*>(out));
  PBK_CHECK_LAUNCH("triangulate_kernel");
  return PBK_OK;
}

static int check_ray_args(const void* pts, int dtype, int stride, int64_t n, const pbk_ray_params* params,
                          const void* out, int min_stride) {
  PBK_REQUIRE(n >= 0 && params != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(dtype == PBK_F32 || dtype == PBK_F64, PBK_EINVAL, "points must be float32 or float64");
  PBK_REQUIRE(stride >= min_stride, PBK_EINVAL, "row stride too small for the requested columns");
  PBK_REQUIRE(n == 0 || (pts && out), PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE((n + 255) / 256 < (1ll << 31), PBK_EINVAL, "n too large");
  return PBK_OK;
}

int pbk_undistort_points(const void* pts, int dtype, int stride, int64_t n, const pbk_ray_params* params,
                         float* out, pbk_stream_t stream) {
  using namespace pbk;
  const int rc = check_ray_args(pts, dtype, stride, n, params, out, 2);
  if (rc != PBK_OK || n == 0) return rc;
  PBK_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7) == 0, PBK_EALIGN, "out must be 8-byte aligned");
  const int blocks = (int)((n + 255) / 256);
  cudaStream_t s = as_stream(stream);
  if (dtype == PBK_F32)
    undistort_points_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(pts), stride, n, *params, reinterpret_cast<float2*>(out));
  else
    undistort_points_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(pts), stride, n, *params, reinterpret_cast<float2*>(out));
  PBK_CHECK_LAUNCH("undistort_points_kernel");
  return PBK_OK;
}

int pbk_camera_rays(const void* pts, int dtype, int stride, int64_t n, const pbk_ray_params* params, double* out,
                    pbk_stream_t stream) {
  using namespace pbk;
  const int rc = check_ray_args(pts, dtype, stride, n, params, out, (params && params->scale_by_z) ? 3 : 2);
  if (rc != PBK_OK || n == 0) return rc;
  PBK_REQUIRE(aligned16(out), PBK_EALIGN, "out must be 16-byte aligned");
  const int blocks = (int)((n + 255) / 256);
  cudaStream_t s = as_stream(stream);
  if (dtype == PBK_F32)
    camera_rays_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(pts), stride, n, *params, out);
  else
    camera_rays_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(pts), stride, n, *params, out);
  PBK_CHECK_LAUNCH("camera_rays_kernel");
  return PBK_OK;
}

int pbk_wire_colors(const uint8_t* bgr, int64_t n_px, int flip_rb, float* out, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n_px >= 0, PBK_EINVAL, "negative size");
  if (n_px == 0) return PBK_OK;
  PBK_REQUIRE(bgr && out, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(bgr) && aligned16(out), PBK_EALIGN, "device pointers must be 16-byte aligned");
  const long long groups = (n_px + 3) / 4, blocks = (groups + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "too many pixels");
  wire_colors_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(bgr, n_px, flip_rb, out);
  PBK_CHECK_LAUNCH("wire_colors_kernel");
  return PBK_OK;
}


int pbk_depth_reconstruct(const void* depth, int depth_dtype, int batch, int H, int W, int skip,
                          const double* xs, const double* ys, double* out, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(batch >= 0 && H >= 0 && W >= 0 && skip >= 1, PBK_EINVAL, "bad shape/skip");
  PBK_REQUIRE(depth_dtype == PBK_U16 || depth_dtype == PBK_F32 || depth_dtype == PBK_F64, PBK_EINVAL,
              "depth dtype must be uint16, float32 or float64");
  if ((long long)batch * H * W == 0) return PBK_OK;
  PBK_REQUIRE(depth && xs && ys && out, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(depth) && aligned16(out), PBK_EALIGN, "device pointers must be 16-byte aligned");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  if (depth_dtype == PBK_U16) return launch_depth<uint16_t>(depth, batch, H, W, skip, xs, ys, out, s, di);
  if (depth_dtype == PBK_F32) return launch_depth<float>(depth, batch, H, W, skip, xs, ys, out, s, di);
  return launch_depth<double>(depth, batch, H, W, skip, xs, ys, out, s, di);
}

int pbk_depth_reconstruct_sparse(const double* pts, const double* depth, int64_t n, double fx, double cx,
                                 double cy, double* out, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0, PBK_EINVAL, "negative size");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(pts && depth && out, PBK_EINVAL, "NULL device pointer");
  const long long blocks = (n + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  depth_reconstruct_sparse_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(pts, depth, n, fx, cx, cy, out);
  PBK_CHECK_LAUNCH("depth_reconstruct_sparse_kernel");
  return PBK_OK;
}

int pbk_sceneflow(const void* dX, int dtype, int H, int W, const pbk_project_params* params,
                  int32_t* winner, float* flow, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(H >= 0 && W >= 0 && params != nullptr, PBK_EINVAL, "bad arguments");
  PBK_REQUIRE(dtype == PBK_F32 || dtype == PBK_F64, PBK_EINVAL, "scene points must be float32 or float64");
  const long long n = (long long)H * W;
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(n < (1ll << 31), PBK_EINVAL, "image too large");
  PBK_REQUIRE(dX && winner && flow, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(flow), PBK_EALIGN, "device pointers must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  PBK_REQUIRE(aligned16(dX) && aligned16(winner), PBK_EALIGN, "device pointers must be 16-byte aligned");
  PBK_CUDA(cudaMemsetAsync(winner, 0xff, (size_t)n * 4, s));          // -1 = no source
  const int blocks = (int)(((n + 3) / 4 + 255) / 256);
  const bool dist = params->has_distortion != 0;
  if (dtype == PBK_F32) {
    if (dist) sceneflow_scatter_kernel<float, true><<<blocks, 256, 0, s>>>(static_cast<const float*>(dX), (int)n, W, H, *params, winner);
    else sceneflow_scatter_kernel<float, false><<<blocks, 256, 0, s>>>(static_cast<const float*>(dX), (int)n, W, H, *params, winner);
  } else {
    if (dist) sceneflow_scatter_kernel<double, true><<<blocks, 256, 0, s>>>(static_cast<const double*>(dX), (int)n, W, H, *params, winner);
    else sceneflow_scatter_kernel<double, false><<<blocks, 256, 0, s>>>(static_cast<const double*>(dX), (int)n, W, H, *params, winner);
  }
  PBK_CHECK_LAUNCH("sceneflow_scatter_kernel");
  FastDiv dW;
  dW.s = 0;
  while ((1ll << dW.s) < W) ++dW.s;
  dW.m = (uint32_t)((((1ull << dW.s) - (unsigned long long)W) << 32) / (unsigned long long)W + 1ull);
  sceneflow_resolve_kernel<<<blocks, 256, 0, s>>>(winner, (int)n, W, dW, flow);
  PBK_CHECK_LAUNCH("sceneflow_resolve_kernel");
  return PBK_OK;
}

}  // extern "C"
