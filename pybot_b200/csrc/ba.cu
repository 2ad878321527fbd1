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
// observation; points (12 B each) are gathered through L2 one tile ahead and the
// warp's camera block (320 B) is cached in shared memory.  One warp owns 32
// observations: the 16 B observation records load coalesced (prefetched two
// tiles ahead), fp64 math in registers, J_cam / J_pt rows are transposed through a
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

__device__ __forceinline__ uint4 ba_load_obs(const BaParams& P, long long i) {
  return i < P.n_obs ? ld_stream_u4(P.obs + i) : make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ int ba_clamp_pt(const BaParams& P, uint32_t raw) {
  const int pt = (int)raw;      // out-of-range indices are clamped (the wrapper validates them)
  return pt < 0 ? 0 : ((long long)pt >= P.n_points ? (int)(P.n_points - 1) : pt);
}

__global__ void __launch_bounds__(kBaWarps * 32, 3)
ba_residual_jacobian_kernel(const __grid_constant__ BaParams P) {
  __shared__ __align__(16) float stage[kBaWarps][32 * 18];   // 12 (Jc) + 6 (Jp) floats per obs
  __shared__ __align__(16) double cam_cache[kBaWarps][kCam]; // the warp's current camera block
  __shared__ double warp_cost[kBaWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* st = stage[warp];
  const double* C = cam_cache[warp];
  const long long ntiles = (P.n_obs + 31) / 32;
  const long long stride = (long long)gridDim.x * kBaWarps;
  double cost = 0.0;
  int cached_cam = -1;

  // software pipeline: the observation record is fetched two tiles ahead and
  // the point one tile ahead, so the DRAM latency of the 16 B record and the
  // L2 latency of the dependent gather overlap with the fp64 math.
  long long tile = (long long)blockIdx.x * kBaWarps + warp;
  uint4 ob = ba_load_obs(P, tile * 32 + lane);
  uint4 ob_n = ba_load_obs(P, (tile + stride) * 32 + lane);
  float px, py, pz;
  {
    const float* pp = P.points + 3ll * ba_clamp_pt(P, ob.y);
    px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
  }

  for (; tile < ntiles; tile += stride) {
    const long long i = tile * 32 + lane;
    const bool live = i < P.n_obs;
    const uint4 ob_nn = ba_load_obs(P, (tile + 2 * stride) * 32 + lane);
    float pxn, pyn, pzn;
    {
      const float* pp = P.points + 3ll * ba_clamp_pt(P, ob_n.y);
      pxn = __ldg(pp); pyn = __ldg(pp + 1); pzn = __ldg(pp + 2);
    }
    int cam = (int)ob.x;
    cam = cam < 0 ? 0 : (cam >= P.n_cams ? P.n_cams - 1 : cam);
    const double uo = (double)__uint_as_float(ob.z), vo = (double)__uint_as_float(ob.w);
    const double X = (double)px, Y = (double)py, Z = (double)pz;
    float2 res = make_float2(0.f, 0.f);
    float4* s4 = reinterpret_cast<float4*>(st);
    float2* s2 = reinterpret_cast<float2*>(st + 384);

    // Observations are sorted by camera in practice, so a warp's 32 records
    // normally share one camera: its 320 B block is staged in shared memory
    // once and read by broadcast LDS at the point of use (no registers pinned,
    // no per-record global gathers).  A tile that straddles cameras takes one
    // pass per distinct camera.
    unsigned todo = 0xffffffffu;
    while (todo) {
      const int c = __shfl_sync(0xffffffffu, cam, __ffs(todo) - 1);
      const bool mine = (cam == c) && ((todo >> lane) & 1u);
      if (c != cached_cam) {
        __syncwarp();
        if (lane < kCam / 2)
          reinterpret_cast<double2*>(cam_cache[warp])[lane] =
              __ldg(reinterpret_cast<const double2*>(P.cams + (long long)c * kCam) + lane);
        cached_cam = c;
        __syncwarp();
      }
      if (mine) {
        const double xc = fma(C[0], X, fma(C[1], Y, fma(C[2], Z, C[9])));
        const double yc = fma(C[3], X, fma(C[4], Y, fma(C[5], Z, C[10])));
        const double zc = fma(C[6], X, fma(C[7], Y, fma(C[8], Z, C[11])));
        const double iz = (zc != 0.0) ? 1.0 / zc : 1.0;
        const double x = xc * iz, y = yc * iz;
        const double ru = fma(P.fx, x, P.cx) - uo;
        const double rv = fma(P.fy, y, P.cy) - vo;
        if (live) {
          res = make_float2((float)ru, (float)rv);
          cost = fma(ru, ru, fma(rv, rv, cost));
        }
        // J_t = [[a, 0, b], [0, c, d]]
        const double a = P.fx * iz, b = -P.fx * x * iz;
        const double cc = P.fy * iz, d = -P.fy * y * iz;
        float jr[6];                                     // d/drvec: [row u | row v]
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          // d(R X)/d r_m = (dR/dr_m) X
          const double* Jm = C + 12 + 9 * m;
          const double g0 = fma(Jm[0], X, fma(Jm[1], Y, Jm[2] * Z));
          const double g1 = fma(Jm[3], X, fma(Jm[4], Y, Jm[5] * Z));
          const double g2 = fma(Jm[6], X, fma(Jm[7], Y, Jm[8] * Z));
          jr[m] = (float)fma(a, g0, b * g2);
          jr[3 + m] = (float)fma(cc, g1, d * g2);
        }
        // stage: Jc rows at [lane*12, +12), Jp rows at [384 + lane*6, +6)
        s4[lane * 3 + 0] = make_float4(jr[0], jr[1], jr[2], (float)a);
        s4[lane * 3 + 1] = make_float4(0.f, (float)b, jr[3], jr[4]);
        s4[lane * 3 + 2] = make_float4(jr[5], 0.f, (float)cc, (float)d);
        s2[lane * 3 + 0] = make_float2((float)fma(a, C[0], b * C[6]), (float)fma(a, C[1], b * C[7]));
        s2[lane * 3 + 1] = make_float2((float)fma(a, C[2], b * C[8]), (float)fma(cc, C[3], d * C[6]));
        s2[lane * 3 + 2] = make_float2((float)fma(cc, C[4], d * C[7]), (float)fma(cc, C[5], d * C[8]));
      }
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (P.r != nullptr && live) st_stream_f2(P.r + 2 * i, res);
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
    ob = ob_n; ob_n = ob_nn;
    px = pxn; py = pyn; pz = pzn;
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
  return resident_grid(di, ba_residual_jacobian_kernel, kBaWarps * 32, 0, (ntiles + kBaWarps - 1) / kBaWarps);
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
