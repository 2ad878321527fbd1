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
#include "pbk_common.cuh"
#include "project_math.cuh"

namespace pbk {

// ------------------------------------------------------ DepthCamera.reconstruct
template <typename T> __device__ __forceinline__ double depth_as_double(T v) { return (double)v; }

constexpr int kDcWarps = 8;
constexpr int kDcTile = 64;                  // output pixels per warp tile (2 per lane)

__device__ __forceinline__ void st_stream_d2(void* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// skip == 1: flat stream over batch*H*W pixels.  One warp owns 64 consecutive
// pixels; lane l computes pixels 2l, 2l+1 (48 B of output), the records are
// transposed through a warp-private stage so every global store is a coalesced
// 16 B row.  4 B in + 24 B out per pixel (float32 depth): HBM-write-bound.
template <typename T>
__global__ void __launch_bounds__(kDcWarps * 32)
depth_reconstruct_dense_kernel(const T* __restrict__ depth, long long total, int W, int H,
                               const double* __restrict__ xs, const double* __restrict__ ys,
                               double* __restrict__ out) {
  __shared__ __align__(16) double stage[kDcWarps][kDcTile * 3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* st = stage[warp];
  const long long ntiles = total / kDcTile;
  const long long stride = (long long)gridDim.x * kDcWarps;
  for (long long tile = (long long)blockIdx.x * kDcWarps + warp; tile < ntiles; tile += stride) {
    const long long p0 = tile * kDcTile + 2 * lane;
    const long long row = p0 / W;
    int x = (int)(p0 - row * W);
    int y = (int)(row % H);
    const double d0 = depth_as_double(depth[p0]), d1 = depth_as_double(depth[p0 + 1]);
    const double xa = __ldg(xs + x), ya = __ldg(ys + y);
    int x1 = x + 1, y1 = y;
    if (x1 == W) { x1 = 0; y1 = (y + 1 == H) ? 0 : y + 1; }
    const double xb = __ldg(xs + x1), yb = __ldg(ys + y1);
    double2* s2 = reinterpret_cast<double2*>(st);
    s2[lane * 3 + 0] = make_double2(__dmul_rn(xa, d0), __dmul_rn(ya, d0));
    s2[lane * 3 + 1] = make_double2(d0, __dmul_rn(xb, d1));
    s2[lane * 3 + 2] = make_double2(__dmul_rn(yb, d1), d1);
    __syncwarp();
    double* dst = out + tile * (kDcTile * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double2 v = s2[j * 32 + lane];
      st_stream_d2(dst + (j * 32 + lane) * 2, v.x, v.y);
    }
    __syncwarp();
  }
}

// One output pixel per thread: skip > 1 and the < 64-pixel tail of the dense kernel.
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
      const int grid = resident_grid(di, kern, kDcWarps * 32, 0, (ntiles + kDcWarps - 1) / kDcWarps);
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

template <typename T, bool DIST>
__global__ void __launch_bounds__(256)
sceneflow_scatter_kernel(const T* __restrict__ dX, int n, int W, int H,
                         const __grid_constant__ pbk_project_params P, int* __restrict__ winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double X[3] = {(double)dX[3ll * i], (double)dX[3ll * i + 1], (double)dX[3ll * i + 2]};
  double u, v;
  project_one<DIST>(P, X, u, v);
  const int xi = cast_i32_like_numpy<T>((T)u), yi = cast_i32_like_numpy<T>((T)v);
  // new_grid[ys[valid], xs[valid]] = grid[gys[valid], gxs[valid]]: on repeated
  // targets numpy keeps the LAST source in index order -> max source index
  if (xi >= 0 && xi < W && yi >= 0 && yi < H) atomicMax(winner + yi * W + xi, i);
}

__global__ void __launch_bounds__(256)
sceneflow_resolve_kernel(const int* __restrict__ winner, int n, int W, float2* __restrict__ flow) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int w = winner[c];
  float2 o;
  if (w < 0) {
    o.x = o.y = __int_as_float(0x7fc00000);          // untouched cells stay NaN
  } else {
    // (source grid coordinate) - (this cell's grid coordinate), float32 like the reference
    o.x = (float)(w % W) - (float)(c % W);
    o.y = (float)(w / W) - (float)(c / W);
  }
  flow[c] = o;
}

}  // namespace pbk

extern "C" {

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
  PBK_CUDA(cudaMemsetAsync(winner, 0xff, (size_t)n * 4, s));          // -1 = no source
  const int blocks = (int)((n + 255) / 256);
  const bool dist = params->has_distortion != 0;
  if (dtype == PBK_F32) {
    if (dist) sceneflow_scatter_kernel<float, true><<<blocks, 256, 0, s>>>(static_cast<const float*>(dX), (int)n, W, H, *params, winner);
    else sceneflow_scatter_kernel<float, false><<<blocks, 256, 0, s>>>(static_cast<const float*>(dX), (int)n, W, H, *params, winner);
  } else {
    if (dist) sceneflow_scatter_kernel<double, true><<<blocks, 256, 0, s>>>(static_cast<const double*>(dX), (int)n, W, H, *params, winner);
    else sceneflow_scatter_kernel<double, false><<<blocks, 256, 0, s>>>(static_cast<const double*>(dX), (int)n, W, H, *params, winner);
  }
  PBK_CHECK_LAUNCH("sceneflow_scatter_kernel");
  sceneflow_resolve_kernel<<<blocks, 256, 0, s>>>(winner, (int)n, W, reinterpret_cast<float2*>(flow));
  PBK_CHECK_LAUNCH("sceneflow_resolve_kernel");
  return PBK_OK;
}

}  // extern "C"
