// Disparity -> XYZ point clouds.
//
// Replaces cv2.reprojectImageTo3D(disp, self.Q) at
// pybot/vision/camera_utils.py:968 (StereoCamera.reconstruct) and :987
// (reconstruct_with_texture, plus its np.copy passes :988-989), and the numpy
// version reconstruct_sparse :971-980.
//
// Arithmetic (bit-exact with OpenCV 4.13, see oracle/model.py):
//   h_i = (((0 + Q[i,0]*x) + Q[i,1]*y) + Q[i,2]*double(d)) + Q[i,3]   fp64, no FMA
//   out_i = float( double(float(h_i)) * (1.0 / h_3) )
// The kernels are HBM streaming kernels: 4 B/px in, 12 B/px out (+3+3 B/px
// with colour).  One warp owns a tile of 128 consecutive pixels: one coalesced
// 16 B load per lane, the 12 output floats per lane are transposed through a
// warp-private shared-memory stage so every global store is a full 512 B
// coalesced STG.128.
#include <string.h>

#include "pbk_common.cuh"

namespace pbk {

struct ReprojParams {
  double q[16];     // Q as double, row-major
  double n2f;       // FAST: double(float(Q[2][3]))
  long long total;  // batch*H*W input pixels
  int W, H;
  int dx, dy;       // FAST/dense: advance of (x, y) per grid stride
};

constexpr int kTilePx = 128;       // pixels per warp tile
constexpr int kWarpsPerCta = 8;

template <typename T> struct Disp4;
template <> struct Disp4<float> {
  static __device__ __forceinline__ void load(const float* b, long long p, float d[4]) {
    float4 v = ld_stream_f4(b + p);
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
};
template <> struct Disp4<int32_t> {
  static __device__ __forceinline__ void load(const int32_t* b, long long p, float d[4]) {
    uint4 v = ld_stream_u4(b + p);
    d[0] = (float)(int)v.x; d[1] = (float)(int)v.y;
    d[2] = (float)(int)v.z; d[3] = (float)(int)v.w;
  }
};
template <> struct Disp4<int16_t> {
  static __device__ __forceinline__ void load(const int16_t* b, long long p, float d[4]) {
    uint2 v = ld_stream_u2(b + p);
    d[0] = (float)(short)(v.x & 0xffff); d[1] = (float)(short)(v.x >> 16);
    d[2] = (float)(short)(v.y & 0xffff); d[3] = (float)(short)(v.y >> 16);
  }
};
template <> struct Disp4<uint8_t> {
  static __device__ __forceinline__ void load(const uint8_t* b, long long p, float d[4]) {
    uint32_t v = ld_stream_u1(b + p);
    d[0] = (float)(v & 0xff); d[1] = (float)((v >> 8) & 0xff);
    d[2] = (float)((v >> 16) & 0xff); d[3] = (float)(v >> 24);
  }
};

// One homogeneous coordinate, OpenCV's Matx product order, no contraction.
__device__ __forceinline__ double hom_row(const double* q, double x, double y, double d) {
  double s = __dadd_rn(0.0, __dmul_rn(q[0], x));
  s = __dadd_rn(s, __dmul_rn(q[1], y));
  s = __dadd_rn(s, __dmul_rn(q[2], d));
  return __dadd_rn(s, q[3]);            // q[3]*1.0 == q[3]
}

__device__ __forceinline__ double round_f32(double v) {
  return (double)__double2float_rn(v);
}

// General 4x4 Q.
__device__ __forceinline__ void reproject_px_general(const ReprojParams& P, int x, int y,
                                                     float d, float out[3]) {
  const double xd = (double)x, yd = (double)y, dd = (double)d;
  const double w = hom_row(P.q + 12, xd, yd, dd);
  const double iw = __drcp_rn(w);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    out[i] = __double2float_rn(__dmul_rn(round_f32(hom_row(P.q + 4 * i, xd, yd, dd)), iw));
}

// Q with the sparsity StereoCamera.Q always has (camera_utils.py:893-896):
//   [[a,0,0,b],[0,c,0,e],[0,0,0,f],[0,0,g,0]] with all structural zeros +0.
// For finite d the zero products add exact +0, so h0 = a*x+b, h1 = c*y+e,
// h2 = f, W = g*d + 0; a non-finite d turns every coordinate into NaN through
// the 0*d products, which the two selects below reproduce.
__device__ __forceinline__ double fast_iw(const ReprojParams& P, float d) {
  const double w = __dadd_rn(__dmul_rn(P.q[14], (double)d), 0.0);
  double iw = __drcp_rn(w);
  if (fabsf(d) == __int_as_float(0x7f800000)) iw = __longlong_as_double(0x7ff8000000000000ll);
  return iw;
}

template <typename T, bool FAST>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
reproject_dense_kernel(const T* __restrict__ disp, float* __restrict__ xyz,
                       const uint8_t* __restrict__ bgr, uint8_t* __restrict__ bgr_out,
                       const __grid_constant__ ReprojParams P) {
  __shared__ __align__(16) float stage[kWarpsPerCta][kTilePx * 3];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float* st = stage[warp];
  const long long ntiles = P.total / kTilePx;
  const long long stride = (long long)gridDim.x * kWarpsPerCta;
  long long tile = (long long)blockIdx.x * kWarpsPerCta + warp;
  if (tile >= ntiles) return;

  // (x, y) of this lane's first pixel; advanced incrementally afterwards.
  int x, y;
  {
    const long long p0 = tile * kTilePx + lane * 4;
    const long long row = p0 / P.W;
    x = (int)(p0 - row * P.W);
    y = (int)(row % P.H);
  }
  float d[4];
  Disp4<T>::load(disp, tile * kTilePx + lane * 4, d);

  for (; tile < ntiles; tile += stride) {
    // software prefetch of the next tile's disparities
    float dn[4] = {0.f, 0.f, 0.f, 0.f};
    const long long nt = tile + stride;
    if (nt < ntiles) Disp4<T>::load(disp, nt * kTilePx + lane * 4, dn);
    uint4 col;
    const bool do_col = (bgr != nullptr) && lane < (kTilePx * 3 / 16);
    if (do_col) col = ld_stream_u4(bgr + tile * (kTilePx * 3) + lane * 16);

    float o[12];
    if (FAST) {
      // row terms for y and the next row (a 4-pixel group crosses at most one
      // row boundary because W >= 4)
      const int y1 = (y + 1 == P.H) ? 0 : y + 1;
      const double n1a = round_f32(__dadd_rn(__dmul_rn(P.q[5], (double)y), P.q[7]));
      const double n1b = round_f32(__dadd_rn(__dmul_rn(P.q[5], (double)y1), P.q[7]));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int xk = x + k;
        double n1 = n1a;
        if (xk >= P.W) { xk -= P.W; n1 = n1b; }
        const double n0 = round_f32(__dadd_rn(__dmul_rn(P.q[0], (double)xk), P.q[3]));
        const double iw = fast_iw(P, d[k]);
        o[3 * k + 0] = __double2float_rn(__dmul_rn(n0, iw));
        o[3 * k + 1] = __double2float_rn(__dmul_rn(n1, iw));
        o[3 * k + 2] = __double2float_rn(__dmul_rn(P.n2f, iw));
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int xk = x + k, yk = y;
        if (xk >= P.W) { xk -= P.W; yk = (y + 1 == P.H) ? 0 : y + 1; }
        reproject_px_general(P, xk, yk, d[k], &o[3 * k]);
      }
    }

    // transpose through the warp-private stage: lane-contiguous 48 B records
    // in (conflict-free: 12-word stride), coalesced 16 B rows out.
    float4* s4 = reinterpret_cast<float4*>(st);
    s4[lane * 3 + 0] = make_float4(o[0], o[1], o[2], o[3]);
    s4[lane * 3 + 1] = make_float4(o[4], o[5], o[6], o[7]);
    s4[lane * 3 + 2] = make_float4(o[8], o[9], o[10], o[11]);
    __syncwarp();
    float* dst = xyz + tile * (kTilePx * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j)
      st_stream_f4(dst + (j * 32 + lane) * 4, s4[j * 32 + lane]);
    if (do_col) st_stream_u4(bgr_out + tile * (kTilePx * 3) + lane * 16, col);
    __syncwarp();

#pragma unroll
    for (int k = 0; k < 4; ++k) d[k] = dn[k];
    x += P.dx;
    y += P.dy;
    if (x >= P.W) { x -= P.W; ++y; }
    if (y >= P.H) y -= P.H;
  }
}

// One output pixel per thread: sub-sampled clouds (sample > 1), images with
// W < 4 and the < 128-pixel tail of the dense kernel.  `first`/`count` select
// a range of OUTPUT pixels of the (batch, Hs, Ws) grid.
template <typename T>
__global__ void __launch_bounds__(256)
reproject_generic_kernel(const T* __restrict__ disp, float* __restrict__ xyz,
                         const uint8_t* __restrict__ bgr, uint8_t* __restrict__ bgr_out,
                         const __grid_constant__ ReprojParams P, int sample, int Ws, int Hs,
                         long long first, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long o = first + i;
  const long long frame = o / ((long long)Ws * Hs);
  const int rem = (int)(o - frame * (long long)Ws * Hs);
  const int ys = rem / Ws, xs = rem - ys * Ws;
  const int x = xs * sample, y = ys * sample;
  const long long src = (frame * P.H + y) * (long long)P.W + x;
  float out[3];
  reproject_px_general(P, x, y, (float)disp[src], out);
  xyz[o * 3 + 0] = out[0];
  xyz[o * 3 + 1] = out[1];
  xyz[o * 3 + 2] = out[2];
  if (bgr != nullptr) {
    bgr_out[o * 3 + 0] = bgr[src * 3 + 0];
    bgr_out[o * 3 + 1] = bgr[src * 3 + 1];
    bgr_out[o * 3 + 2] = bgr[src * 3 + 2];
  }
}

// reconstruct_sparse: float64 rows [x, y, d] -> (Q*[x,y,d,1])[:3] / W.
__global__ void __launch_bounds__(256)
reproject_sparse_kernel(const double* __restrict__ xyd, long long n,
                        const __grid_constant__ ReprojParams P, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xyd[3 * i], y = xyd[3 * i + 1], d = xyd[3 * i + 2];
  double h[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    h[r] = P.q[4 * r] * x + P.q[4 * r + 1] * y + P.q[4 * r + 2] * d + P.q[4 * r + 3];
  out[3 * i + 0] = h[0] / h[3];
  out[3 * i + 1] = h[1] / h[3];
  out[3 * i + 2] = h[2] / h[3];
}

static bool is_pos_zero(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return b == 0;
}
static bool finite_not_negzero(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return b != 0x8000000000000000ull && v == v && v - v == 0.0;
}

static bool q_is_stereo_sparse(const double* q) {
  const int zeros[] = {1, 2, 4, 6, 8, 9, 10, 12, 13, 15};
  for (int i : zeros)
    if (!is_pos_zero(q[i])) return false;
  return finite_not_negzero(q[0]) && finite_not_negzero(q[3]) && finite_not_negzero(q[5]) &&
         finite_not_negzero(q[7]) && finite_not_negzero(q[11]) && finite_not_negzero(q[14]);
}

template <typename T>
static int launch_reproject(const void* disp_v, int batch, int H, int W, const ReprojParams& P0,
                            int sample, float* xyz, const uint8_t* bgr, uint8_t* bgr_out,
                            cudaStream_t stream, const DeviceInfo* di) {
  const T* disp = static_cast<const T*>(disp_v);
  ReprojParams P = P0;
  const long long total = (long long)batch * H * W;
  if (sample == 1 && W >= 4) {
    const long long ntiles = total / kTilePx;
    if (ntiles > 0) {
      const long long ctas_needed = (ntiles + kWarpsPerCta - 1) / kWarpsPerCta;
      const bool fast = q_is_stereo_sparse(P.q);
      auto kern = fast ? reproject_dense_kernel<T, true> : reproject_dense_kernel<T, false>;
      static KernelCache kc[2];
      int per_sm;
      const int rc = kc[fast ? 1 : 0].setup(di, kern, kWarpsPerCta * 32, 0, &per_sm);
      if (rc) return rc;
      const int grid = resident_grid(di, per_sm, ctas_needed);
      const long long stride_px = (long long)grid * kWarpsPerCta * kTilePx;
      P.dx = (int)(stride_px % W);
      P.dy = (int)((stride_px / W) % H);
      kern<<<grid, kWarpsPerCta * 32, 0, stream>>>(disp, xyz, bgr, bgr_out, P);
      PBK_CHECK_LAUNCH("reproject_dense_kernel");
    }
    const long long done = ntiles * kTilePx;
    if (done < total) {
      const long long rest = total - done;
      reproject_generic_kernel<T><<<(int)((rest + 255) / 256), 256, 0, stream>>>(
          disp, xyz, bgr, bgr_out, P, 1, W, H, done, rest);
      PBK_CHECK_LAUNCH("reproject_generic_kernel(tail)");
    }
  } else {
    const int Ws = (W + sample - 1) / sample, Hs = (H + sample - 1) / sample;
    const long long outs = (long long)batch * Ws * Hs;
    const long long blocks = (outs + 255) / 256;
    PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "too many output pixels");
    reproject_generic_kernel<T><<<(int)blocks, 256, 0, stream>>>(disp, xyz, bgr, bgr_out, P, sample,
                                                                 Ws, Hs, 0, outs);
    PBK_CHECK_LAUNCH("reproject_generic_kernel");
  }
  return PBK_OK;
}

}  // namespace pbk

extern "C" {

int pbk_reproject_disp(const void* disp, int disp_dtype, int batch, int H, int W, const float* Q,
                       int sample, float* xyz, const uint8_t* bgr, uint8_t* bgr_out,
                       pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(disp_dtype != PBK_F64, PBK_EUNSUPPORTED,
              "float64 disparity is rejected (cv2.reprojectImageTo3D rejects it too)");
  PBK_REQUIRE(disp_dtype >= PBK_U8 && disp_dtype <= PBK_F32, PBK_EINVAL, "bad disp_dtype %d", disp_dtype);
  PBK_REQUIRE(batch >= 0 && H >= 0 && W >= 0 && sample >= 1, PBK_EINVAL, "bad shape/sample");
  PBK_REQUIRE(Q != nullptr, PBK_EINVAL, "Q is NULL");
  PBK_REQUIRE((bgr == nullptr) == (bgr_out == nullptr), PBK_EINVAL, "bgr and bgr_out must both be given or both NULL");
  if ((long long)batch * H * W == 0) return PBK_OK;
  PBK_REQUIRE(disp != nullptr && xyz != nullptr, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(disp) && aligned16(xyz) && aligned16(bgr) && aligned16(bgr_out), PBK_EALIGN,
              "device pointers must be 16-byte aligned");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  ReprojParams P;
  for (int i = 0; i < 16; ++i) P.q[i] = (double)Q[i];
  P.n2f = (double)(float)P.q[11];
  P.total = (long long)batch * H * W;
  P.W = W; P.H = H; P.dx = 0; P.dy = 0;
  cudaStream_t s = as_stream(stream);
  switch (disp_dtype) {
    case PBK_U8: return launch_reproject<uint8_t>(disp, batch, H, W, P, sample, xyz, bgr, bgr_out, s, di);
    case PBK_S16: return launch_reproject<int16_t>(disp, batch, H, W, P, sample, xyz, bgr, bgr_out, s, di);
    case PBK_S32: return launch_reproject<int32_t>(disp, batch, H, W, P, sample, xyz, bgr, bgr_out, s, di);
    default: return launch_reproject<float>(disp, batch, H, W, P, sample, xyz, bgr, bgr_out, s, di);
  }
}

int pbk_reproject_sparse(const double* xyd, int64_t n, const float* Q, double* out,
                         pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n >= 0 && Q != nullptr, PBK_EINVAL, "bad arguments");
  if (n == 0) return PBK_OK;
  PBK_REQUIRE(xyd != nullptr && out != nullptr, PBK_EINVAL, "NULL device pointer");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  ReprojParams P = {};
  for (int i = 0; i < 16; ++i) P.q[i] = (double)Q[i];
  const long long blocks = (n + 255) / 256;
  PBK_REQUIRE(blocks < (1ll << 31), PBK_EINVAL, "n too large");
  reproject_sparse_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(xyd, n, P, out);
  PBK_CHECK_LAUNCH("reproject_sparse_kernel");
  return PBK_OK;
}

}  // extern "C"
