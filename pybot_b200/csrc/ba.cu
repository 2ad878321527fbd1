// Bundle-adjustment reprojection residual + 2x6 / 2x3 Jacobians
// (SURVEY.md 8a row a10, Appendix A.5).
//
// Replaces, per observation (camera c, point p, measured u v):
//   x, J = cv2.projectPoints(X_p, rvec_c, tvec_c, K, D=0)       (the Jacobian
//   pybot/vision/camera_utils.py:697 computes and discards, columns 0:6)
//   r = x - z_obs ;  J_cam = [dx/drvec | dx/dtvec] ;  J_pt = dx/dX = J_t * R
// The reference has no BA; parity is pinned against OpenCV.
//
// Two kernels.  ba_camera_precompute turns each rvec into R (9) and dR/dr (27)
// by OpenCV's Rodrigues closed form (one thread per camera, fp64).  The
// streaming kernel is HBM-write-bound: 16 B read + 80 B written per
// observation; cameras (320 B each) and points (12 B each) are gathered through
// L1/L2.  One warp owns 32 observations: the 16 B observation records load
// coalesced, fp64 math in registers, J_cam / J_pt rows are transposed through a
// warp-private shared-memory stage so global stores are coalesced 16 B, and
// sum ||r||^2 is reduced warp -> CTA -> one fp64 partial per CTA, summed in
// fixed order by a second tiny kernel (deterministic, no atomics).
#include "pbk_common.cuh"

namespace pbk {

constexpr int kBaWarps = 8;
constexpr int kCam = PBK_BA_CAM_DOUBLES;   // R(9) t(3) dRdr(27) pad(1)

__global__ void __launch_bounds__(128)
ba_camera_precompute_kernel(const float* __restrict__ rvec, const float* __restrict__ tvec, int n,
                            double* __restrict__ blocks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double rx = rvec[3 * c], ry = rvec[3 * c + 1], rz = rvec[3 * c + 2];
  double* o = blocks + (long long)c * kCam;
  o[9] = tvec[3 * c];
  o[10] = tvec[3 * c + 1];
  o[11] = tvec[3 * c + 2];
  o[39] = 0.0;
  const double theta = sqrt(rx * rx + ry * ry + rz * rz);
  double R[9], J[27];
  if (theta < 2.220446049250313e-16) {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 27; ++i) J[i] = 0.0;
    J[5] = -1; J[7] = 1; J[9 + 2] = 1; J[9 + 6] = -1; J[18 + 1] = -1; J[18 + 3] = 1;
  } else {
    double s, c_;
    sincos(theta, &s, &c_);
    const double c1 = 1.0 - c_, itheta = 1.0 / theta;
    const double u[3] = {rx * itheta, ry * itheta, rz * itheta};
    const double rrt[9] = {u[0] * u[0], u[0] * u[1], u[0] * u[2], u[0] * u[1], u[1] * u[1],
                           u[1] * u[2], u[0] * u[2], u[1] * u[2], u[2] * u[2]};
    const double r_x[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; ++k) R[k] = c_ * I[k] + c1 * rrt[k] + s * r_x[k];
    const double drrt[27] = {u[0] + u[0], u[1], u[2], u[1], 0, 0, u[2], 0, 0,
                             0, u[0], 0, u[0], u[1] + u[1], u[2], 0, u[2], 0,
                             0, 0, u[0], 0, 0, u[1], u[0], u[1], u[2] + u[2]};
    const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0,
                              0, 0, 1, 0, 0, 0, -1, 0, 0,
                              0, -1, 0, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      const double ri = u[i];
      const double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta,
                   a3 = (c_ - s * itheta) * ri, a4 = s * itheta;
      for (int k = 0; k < 9; ++k)
        J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
    }
  }
  for (int i = 0; i < 9; ++i) o[i] = R[i];
  for (int i = 0; i < 27; ++i) o[12 + i] = J[i];
}

struct BaParams {
  const int4* obs;            // {cam, pt, u bits, v bits}
  const float* points;
  const double* cams;
  float* r;
  float* Jc;
  float* Jp;
  double* partials;
  long long n_obs;
  long long n_points;
  int n_cams;
  double fx, fy, cx, cy;
};

__global__ void __launch_bounds__(kBaWarps * 32)
ba_residual_jacobian_kernel(const __grid_constant__ BaParams P) {
  __shared__ __align__(16) float stage[kBaWarps][32 * 18];   // 12 (Jc) + 6 (Jp) floats per obs
  __shared__ double warp_cost[kBaWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* st = stage[warp];
  const long long ntiles = (P.n_obs + 31) / 32;
  const long long stride = (long long)gridDim.x * kBaWarps;
  double cost = 0.0;

  for (long long tile = (long long)blockIdx.x * kBaWarps + warp; tile < ntiles; tile += stride) {
    const long long i = tile * 32 + lane;
    const bool live = i < P.n_obs;
    float jc[12], jp[6];
    float2 res = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 12; ++k) jc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) jp[k] = 0.f;
    if (live) {
      const uint4 ob = ld_stream_u4(P.obs + i);
      int cam = (int)ob.x, pt = (int)ob.y;
      // out-of-range indices are clamped (the wrapper validates them)
      cam = cam < 0 ? 0 : (cam >= P.n_cams ? P.n_cams - 1 : cam);
      pt = pt < 0 ? 0 : ((long long)pt >= P.n_points ? (int)(P.n_points - 1) : pt);
      const double uo = (double)__uint_as_float(ob.z), vo = (double)__uint_as_float(ob.w);
      const float* pp = P.points + 3ll * pt;
      const double X = (double)__ldg(pp), Y = (double)__ldg(pp + 1), Z = (double)__ldg(pp + 2);
      const double* cb = P.cams + (long long)cam * kCam;
      double R[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = __ldg(cb + k);
      const double xc = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, __ldg(cb + 9))));
      const double yc = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, __ldg(cb + 10))));
      const double zc = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, __ldg(cb + 11))));
      const double iz = (zc != 0.0) ? 1.0 / zc : 1.0;
      const double x = xc * iz, y = yc * iz;
      const double ru = fma(P.fx, x, P.cx) - uo;
      const double rv = fma(P.fy, y, P.cy) - vo;
      res = make_float2((float)ru, (float)rv);
      cost = fma(ru, ru, fma(rv, rv, cost));
      // J_t = [[a, 0, b], [0, c, d]]
      const double a = P.fx * iz, b = -P.fx * x * iz;
      const double c = P.fy * iz, d = -P.fy * y * iz;
      jc[3] = (float)a;  jc[4] = 0.f;       jc[5] = (float)b;
      jc[9] = 0.f;       jc[10] = (float)c; jc[11] = (float)d;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        // d(R X)/d r_m : rows 0..2 of (dR/dr_m) X
        const double* Jm = cb + 12 + 9 * m;
        const double g0 = fma(__ldg(Jm + 0), X, fma(__ldg(Jm + 1), Y, __ldg(Jm + 2) * Z));
        const double g1 = fma(__ldg(Jm + 3), X, fma(__ldg(Jm + 4), Y, __ldg(Jm + 5) * Z));
        const double g2 = fma(__ldg(Jm + 6), X, fma(__ldg(Jm + 7), Y, __ldg(Jm + 8) * Z));
        jc[m] = (float)fma(a, g0, b * g2);
        jc[6 + m] = (float)fma(c, g1, d * g2);
        jp[m] = (float)fma(a, R[m], b * R[6 + m]);
        jp[3 + m] = (float)fma(c, R[3 + m], d * R[6 + m]);
      }
    }
    if (P.r != nullptr && live) st_stream_f2(P.r + 2 * i, res);

    // stage: Jc rows at [lane*12, +12), Jp rows at [384 + lane*6, +6)
    float4* s4 = reinterpret_cast<float4*>(st);
    s4[lane * 3 + 0] = make_float4(jc[0], jc[1], jc[2], jc[3]);
    s4[lane * 3 + 1] = make_float4(jc[4], jc[5], jc[6], jc[7]);
    s4[lane * 3 + 2] = make_float4(jc[8], jc[9], jc[10], jc[11]);
    float2* s2 = reinterpret_cast<float2*>(st + 384);
    s2[lane * 3 + 0] = make_float2(jp[0], jp[1]);
    s2[lane * 3 + 1] = make_float2(jp[2], jp[3]);
    s2[lane * 3 + 2] = make_float2(jp[4], jp[5]);
    __syncwarp();
    const long long o0 = tile * 32;
    const int nlive = (int)((P.n_obs - o0) < 32 ? (P.n_obs - o0) : 32);
    if (P.Jc != nullptr) {
      float* dst = P.Jc + o0 * 12;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int v4 = j * 32 + lane;               // float4 index within the tile
        if (v4 * 4 < nlive * 12) st_stream_f4(dst + v4 * 4, s4[v4]);
      }
    }
    if (P.Jp != nullptr) {
      float* dst = P.Jp + o0 * 6;                   // 24 B per obs: tile base is 16 B aligned
      const float4* sp = reinterpret_cast<const float4*>(st + 384);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int v4 = j * 32 + lane;
        if (v4 < 48) {
          if (v4 * 4 + 4 <= nlive * 6) st_stream_f4(dst + v4 * 4, sp[v4]);
          else if (v4 * 4 < nlive * 6) {            // ragged last tile: 8 B granularity
            st_stream_f2(dst + v4 * 4, make_float2(sp[v4].x, sp[v4].y));
          }
        }
      }
    }
    __syncwarp();
  }

  // cost: warp -> CTA -> partial
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
  if (lane == 0) warp_cost[warp] = cost;
  __syncthreads();
  if (threadIdx.x == 0 && P.partials != nullptr) {
    double s = 0.0;
    for (int w = 0; w < kBaWarps; ++w) s += warp_cost[w];
    P.partials[blockIdx.x] = s;
  }
}

__global__ void ba_cost_sum_kernel(const double* __restrict__ partials, int n, double* __restrict__ cost) {
  // one warp, fixed order: lane-strided partial sums then a shuffle tree
  const int lane = threadIdx.x;
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) *cost = s;
}

static int ba_grid(const DeviceInfo* di, long long n_obs) {
  const long long ntiles = (n_obs + 31) / 32;
  return stream_grid(di, (ntiles + kBaWarps - 1) / kBaWarps, 4);
}

}  // namespace pbk

extern "C" {

int pbk_ba_camera_precompute(const float* rvec, const float* tvec, int n_cams, double* cam_blocks,
                             pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n_cams >= 0, PBK_EINVAL, "negative camera count");
  if (n_cams == 0) return PBK_OK;
  PBK_REQUIRE(rvec && tvec && cam_blocks, PBK_EINVAL, "NULL device pointer");
  ba_camera_precompute_kernel<<<(n_cams + 127) / 128, 128, 0, as_stream(stream)>>>(rvec, tvec, n_cams,
                                                                                  cam_blocks);
  PBK_CHECK_LAUNCH("ba_camera_precompute_kernel");
  return PBK_OK;
}

int pbk_ba_cost_slots(int64_t n_obs) {
  using namespace pbk;
  const DeviceInfo* di;
  if (current_device_info(&di)) return -1;
  return ba_grid(di, n_obs > 0 ? n_obs : 1);
}

int pbk_ba_residual_jacobian(const void* obs, int64_t n_obs, const float* points, int64_t n_points,
                             const double* cam_blocks, int n_cams, double fx, double fy, double cx,
                             double cy, float* r, float* J_cam, float* J_pt, double* cost_partials,
                             double* cost, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n_obs >= 0 && n_points >= 0 && n_cams >= 0, PBK_EINVAL, "negative size");
  PBK_REQUIRE((cost == nullptr) || (cost_partials != nullptr), PBK_EINVAL, "cost needs cost_partials");
  cudaStream_t s = as_stream(stream);
  if (n_obs == 0) {
    if (cost != nullptr) PBK_CUDA(cudaMemsetAsync(cost, 0, 8, s));
    return PBK_OK;
  }
  PBK_REQUIRE(n_cams > 0 && n_points > 0, PBK_EINVAL, "observations without cameras or points");
  PBK_REQUIRE(obs && points && cam_blocks, PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(obs) && aligned16(r) && aligned16(J_cam) && aligned16(J_pt) && aligned16(cam_blocks),
              PBK_EALIGN, "device pointers must be 16-byte aligned");
  const DeviceInfo* di;
  int rc = current_device_info(&di);
  if (rc) return rc;
  BaParams P;
  P.obs = static_cast<const int4*>(obs);
  P.points = points; P.cams = cam_blocks;
  P.r = r; P.Jc = J_cam; P.Jp = J_pt; P.partials = cost_partials;
  P.n_obs = n_obs; P.n_points = n_points; P.n_cams = n_cams;
  P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
  const int grid = ba_grid(di, n_obs);
  ba_residual_jacobian_kernel<<<grid, kBaWarps * 32, 0, s>>>(P);
  PBK_CHECK_LAUNCH("ba_residual_jacobian_kernel");
  if (cost != nullptr) {
    ba_cost_sum_kernel<<<1, 32, 0, s>>>(cost_partials, grid, cost);
    PBK_CHECK_LAUNCH("ba_cost_sum_kernel");
  }
  return PBK_OK;
}

}  // extern "C"
