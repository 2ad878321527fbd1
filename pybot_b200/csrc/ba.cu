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
// warp's camera block (320 B) is cached in shared memory.  One warp owns 64
// observations (2 per lane): the 16 B observation records arrive through a
// per-warp TMA ring (three tiles in flight), fp64 math in registers with every camera
// coefficient read once for both observations, r / J_cam / J_pt rows are written
// to a warp-private shared-memory stage and leave as TMA bulk stores
// (cp.async.bulk shared -> global), which keeps the output bytes off the LSU
// data path - the first version of this kernel was L1-wavefront-bound (ncu
// l1tex 83 %), not HBM-bound.  sum ||r||^2 is reduced warp -> CTA -> one fp64
// partial per CTA; the last CTA to finish adds the partials in fixed order
// (deterministic, no floating-point atomics, no second launch).
#include "pbk_common.cuh"

namespace pbk {

constexpr int kBaWarps = 8;
constexpr int kBaMaxWorld = 16;
constexpr int kCam = PBK_BA_CAM_DOUBLES;   // R(9) t(3) dRdr(27) pad(1)

// One camera block = R (9), t (3), dR/dr (27), pad: OpenCV's Rodrigues closed form in fp64,
// computed by ONE WARP (entry e of the 36 on lane e % 32): the latency is that of a single
// sincos + a dozen dependent fp64 operations instead of a 300-operation serial chain.  Used
// by the stand-alone precompute kernel and by the prologue of the streaming kernel (same
// function, same SASS: the two routes give bit-identical blocks).
__device__ __forceinline__ double sel3(int i, double a, double b, double c) { return i == 0 ? a : (i == 1 ? b : c); }

__device__ __noinline__ void ba_camera_block_warp(const float* __restrict__ rvec, const float* __restrict__ tvec,
                                                  int c, double* __restrict__ o, double* o2, int lane) {
  // o: the camera's block in global memory; o2: optional second copy (a warp's shared-memory cache)
  const double rx = rvec[3 * c], ry = rvec[3 * c + 1], rz = rvec[3 * c + 2];
  if (lane < 3) {
    const double t = (double)tvec[3 * c + lane];
    o[9 + lane] = t;
    if (o2 != nullptr) o2[9 + lane] = t;
  }
  if (lane == 3) {
    o[39] = 0.0;
    if (o2 != nullptr) o2[39] = 0.0;
  }
  const double theta = sqrt(rx * rx + ry * ry + rz * rz);
  const bool tiny = theta < 2.220446049250313e-16;
  double s = 0.0, c_ = 1.0;
  if (!tiny) sincos(theta, &s, &c_);
  const double c1 = 1.0 - c_, itheta = tiny ? 0.0 : 1.0 / theta;
  const double u0 = rx * itheta, u1 = ry * itheta, u2 = rz * itheta;
  for (int e = lane; e < 36; e += 32) {
    const int k = e < 9 ? e : (e - 9) % 9, i = e < 9 ? 0 : (e - 9) / 9;
    const int row = k / 3, col = k % 3;
    const double Ik = row == col ? 1.0 : 0.0;
    const double ur = sel3(row, u0, u1, u2), uc = sel3(col, u0, u1, u2);
    const double rrt = ur * uc;
    // [u]x: entry (row, col) = -+ u_m, m the third index; sign - for (0,1) (1,2) (2,0)
    const int m = 3 - row - col;
    const double sgn = row == col ? 0.0 : (((col - row + 3) % 3 == 1) ? -1.0 : 1.0);
    const double r_x = row == col ? 0.0 : sgn * sel3(m, u0, u1, u2);
    double v;
    if (e < 9) {
      v = tiny ? Ik : c_ * Ik + c1 * rrt + s * r_x;
    } else {
      const double d_r_x = (row != col && m == i) ? sgn : 0.0;          // d [u]x / d u_i
      if (tiny) {
        v = d_r_x;
      } else {
        const double ri = sel3(i, u0, u1, u2);
        const double drrt = (row == i ? uc : 0.0) + (col == i ? ur : 0.0);
        const double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta,
                     a3 = (c_ - s * itheta) * ri, a4 = s * itheta;
        v = a0 * Ik + a1 * rrt + a2 * drrt + a3 * r_x + a4 * d_r_x;
      }
    }
    o[e < 9 ? e : e + 3] = v;
    if (o2 != nullptr) o2[e < 9 ? e : e + 3] = v;
  }
}

// stand-alone precompute (callers that keep the two-kernel route, unsorted observations):
// one warp per camera
__global__ void __launch_bounds__(128)
ba_camera_precompute_kernel(const float* __restrict__ rvec, const float* __restrict__ tvec, int n,
                            double* __restrict__ blocks) {
  // programmatic dependent launch: let the streaming kernel (pbk_ba_evaluate launches it with
  // the programmatic-serialization attribute) start its prologue now; it executes
  // griddepcontrol.wait before it reads the first camera block.  No-op otherwise.
  asm volatile("griddepcontrol.launch_dependents;");
  const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= n) return;
  ba_camera_block_warp(rvec, tvec, c, blocks + (long long)c * kCam, nullptr, threadIdx.x & 31);
}

struct BaParams {
  const int4* obs;            // {cam, pt, u bits, v bits}
  const float* points;
  const double* cams;
  // fused camera precompute (observations sorted by camera): every CTA derives the blocks of
  // the cameras ITS observation range references into cams_w (== cams) in its prologue
  const float* rvec;
  const float* tvec;
  double* cams_w;
  float* r;
  float* Jc;
  float* Jp;
  double* partials;           // [gridDim.x] per-CTA sums, [gridDim.x] a u64 arrival counter (zero between
                              // calls), [gridDim.x + 1] the device-resident call sequence (seq == 0)
  double* cost;
  // multi-GPU cost exchange over peer-mapped memory (world > 1): peers[p] = rank p's exchange
  // buffer as mapped here, 2 (call parity) x world x {value bits, sequence number} u64 words
  unsigned long long* peers[kBaMaxWorld];
  int rank, world;
  unsigned long long seq;     // 0: use (and advance) the device-resident sequence in `partials`
  long long timeout_clk;      // bound of the peer wait (SM clocks)
  uint32_t* status;           // set to kStatusPeerTimeout when the wait ran into the bound
  long long n_obs;
  long long n_points;
  int n_cams;
  double fx, fy, cx, cy;
};

// Compile-time shape of the streaming kernel: OPL observations per lane.
template <int OPL>
struct BaCfg {
  static constexpr int kTile = 32 * OPL;              // observations per warp tile
  static constexpr int kStJc = 0;                     // stage layout (bytes): J_cam rows
  static constexpr int kStJp = kTile * 48;            //                       J_pt rows
  static constexpr int kStR = kStJp + kTile * 24;     //                       residuals
  static constexpr int kStage = kStR + kTile * 8;     // 80 B per observation
#ifndef PBK_BA_STAGE_BUFS
#define PBK_BA_STAGE_BUFS 1
#endif
  static constexpr int kStageBufs = PBK_BA_STAGE_BUFS; // output stages per warp (bulk stores of tile k-1 drain while tile k is computed)
  static constexpr int kRingStages = 3;
  static constexpr int kObsBytes = kTile * 16;        // one tile of observation records
  // dynamic shared memory layout
  static constexpr int kSmemRing = kBaWarps * kStageBufs * kStage;
  static constexpr int kSmemCam = kSmemRing + kBaWarps * kRingStages * kObsBytes;
  static constexpr int kSmemBars = kSmemCam + kBaWarps * kCam * 8;
  static constexpr int kSmemCost = kSmemBars + kBaWarps * kRingStages * 8;
  static constexpr int kSmemBytes = kSmemCost + kBaWarps * 8;
};

__device__ __forceinline__ uint4 ba_load_obs(const BaParams& P, long long i) {
  return i < P.n_obs ? ld_stream_u4(P.obs + i) : make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void ba_load_point(const BaParams& P, uint32_t raw, float (&p)[3]) {
  int pt = (int)raw;            // out-of-range indices are clamped (the wrapper validates them)
#ifdef PBK_BA_PROBE_NOGATHER
  pt = (int)((threadIdx.x + blockIdx.x * 256) & 0xffff);      // probe: coalesced, L1-resident "gather"
#endif
  pt = pt < 0 ? 0 : ((long long)pt >= P.n_points ? (int)(P.n_points - 1) : pt);
  const float* pp = P.points + 3ll * pt;
  // volatile: these loads must not be hoisted above the conversion of the previous
  // tile's points (see process()), or both tiles' loads share a scoreboard
  // the 12 MB point array is re-read ~4x while 320 MB of outputs stream through L2: keep it
  // resident (evict_last); measured 84.2 -> 82.0 us on C4
  const uint64_t pol = l2_policy_evict_last();
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(p[0]) : "l"(pp), "l"(pol));
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(p[1]) : "l"(pp + 1), "l"(pol));
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(p[2]) : "l"(pp + 2), "l"(pol));

}

// Residual + Jacobian blocks of N observations that share the camera block C
// (shared memory, 40 doubles).  The block is pulled into registers in three
// groups of 16 B broadcast loads (R|t, dR/dr0|dR/dr1, dR/dr2) and every
// coefficient is used for all N observations, so the shared-memory traffic per
// observation falls with N.  Results go straight to the warp's stage:
//   J_cam row [lane-slot*12, +12), J_pt row [slot*6, +6), r [slot*2, +2).
template <int N, int TILE>
__device__ __forceinline__ void ba_eval(const double* C, const BaParams& P, const double (&X)[N][3],
                                        const uint4 (&ob)[N], const int (&slot)[N], const bool (&act)[N],
                                        const bool (&live)[N], char* st, double& cost) {
  const double2* C2 = reinterpret_cast<const double2*>(C);
  float4* sJc = reinterpret_cast<float4*>(st);
  float2* sJp = reinterpret_cast<float2*>(st + TILE * 48);
  float2* sR = reinterpret_cast<float2*>(st + TILE * 72);
#ifdef PBK_BA_PROBE_NOMATH
  // probe: the memory skeleton alone (records in, gathers, stage writes, bulk stores out)
#pragma unroll
  for (int j = 0; j < N; ++j) {
    if (act[j]) {
      const float x0 = (float)X[j][0], x1 = (float)X[j][1], x2 = (float)X[j][2] + (float)C[0];
      const float uo = __uint_as_float(ob[j].z), vo = __uint_as_float(ob[j].w);
      if (live[j]) cost += (double)uo;
      sR[slot[j]] = make_float2(uo, vo);
      sJp[slot[j] * 3 + 0] = make_float2(x0, x1);
      sJp[slot[j] * 3 + 1] = make_float2(x2, x0);
      sJp[slot[j] * 3 + 2] = make_float2(x1, x2);
      sJc[slot[j] * 3 + 0] = make_float4(x0, x1, x2, uo);
      sJc[slot[j] * 3 + 1] = make_float4(x1, x2, uo, x0);
      sJc[slot[j] * 3 + 2] = make_float4(x2, uo, x0, x1);
    }
  }
  return;
#endif
  double a[N], b[N], cc[N], d[N];
  {
    const double2 c0 = C2[0], c1 = C2[1], c2 = C2[2], c3 = C2[3], c4 = C2[4], c5 = C2[5];
    const double R0 = c0.x, R1 = c0.y, R2 = c1.x, R3 = c1.y, R4 = c2.x, R5 = c2.y, R6 = c3.x, R7 = c3.y,
                 R8 = c4.x, t0 = c4.y, t1 = c5.x, t2 = c5.y;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double uo = (double)__uint_as_float(ob[j].z), vo = (double)__uint_as_float(ob[j].w);
      const double xc = fma(R0, X[j][0], fma(R1, X[j][1], fma(R2, X[j][2], t0)));
      const double yc = fma(R3, X[j][0], fma(R4, X[j][1], fma(R5, X[j][2], t1)));
      const double zc = fma(R6, X[j][0], fma(R7, X[j][1], fma(R8, X[j][2], t2)));
      const double iz = (zc != 0.0) ? 1.0 / zc : 1.0;
      const double x = xc * iz, y = yc * iz;
      const double ru = fma(P.fx, x, P.cx) - uo;
      const double rv = fma(P.fy, y, P.cy) - vo;
      // J_t = [[a, 0, b], [0, cc, d]]
      a[j] = P.fx * iz; b[j] = -P.fx * x * iz;
      cc[j] = P.fy * iz; d[j] = -P.fy * y * iz;
      if (act[j]) {
        if (live[j]) cost = fma(ru, ru, fma(rv, rv, cost));
        sR[slot[j]] = make_float2((float)ru, (float)rv);
        // J_pt = J_t * R
        sJp[slot[j] * 3 + 0] = make_float2((float)fma(a[j], R0, b[j] * R6), (float)fma(a[j], R1, b[j] * R7));
        sJp[slot[j] * 3 + 1] = make_float2((float)fma(a[j], R2, b[j] * R8), (float)fma(cc[j], R3, d[j] * R6));
        sJp[slot[j] * 3 + 2] = make_float2((float)fma(cc[j], R4, d[j] * R7), (float)fma(cc[j], R5, d[j] * R8));
      }
    }
  }
  float jr[N][6];                                  // d/drvec: [row u | row v]
  {
    // dR/dr_0 = C[12..20], dR/dr_1 = C[21..29]  (18 doubles, 16 B aligned)
    double J[18];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const double2 v = C2[6 + k];
      J[2 * k] = v.x; J[2 * k + 1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const double* Jm = J + 9 * m;                // d(R X)/d r_m = (dR/dr_m) X
        const double g0 = fma(Jm[0], X[j][0], fma(Jm[1], X[j][1], Jm[2] * X[j][2]));
        const double g1 = fma(Jm[3], X[j][0], fma(Jm[4], X[j][1], Jm[5] * X[j][2]));
        const double g2 = fma(Jm[6], X[j][0], fma(Jm[7], X[j][1], Jm[8] * X[j][2]));
        jr[j][m] = (float)fma(a[j], g0, b[j] * g2);
        jr[j][3 + m] = (float)fma(cc[j], g1, d[j] * g2);
      }
    }
  }
  {
    // dR/dr_2 = C[30..38] (+ pad)
    double J[10];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double2 v = C2[15 + k];
      J[2 * k] = v.x; J[2 * k + 1] = v.y;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double g0 = fma(J[0], X[j][0], fma(J[1], X[j][1], J[2] * X[j][2]));
      const double g1 = fma(J[3], X[j][0], fma(J[4], X[j][1], J[5] * X[j][2]));
      const double g2 = fma(J[6], X[j][0], fma(J[7], X[j][1], J[8] * X[j][2]));
      jr[j][2] = (float)fma(a[j], g0, b[j] * g2);
      jr[j][5] = (float)fma(cc[j], g1, d[j] * g2);
      if (act[j]) {
        sJc[slot[j] * 3 + 0] = make_float4(jr[j][0], jr[j][1], jr[j][2], (float)a[j]);
        sJc[slot[j] * 3 + 1] = make_float4(0.f, (float)b[j], jr[j][3], jr[j][4]);
        sJc[slot[j] * 3 + 2] = make_float4(jr[j][5], 0.f, (float)cc[j], (float)d[j]);
      }
    }
  }
}

// plain copy of the first `bytes` (multiple of 8) of a stage region to global
__device__ __forceinline__ void ba_copy_out(float* gdst, const char* ssrc, int bytes, int lane) {
  for (int off = lane * 8; off < bytes; off += 256)
    st_stream_f2(reinterpret_cast<char*>(gdst) + off, *reinterpret_cast<const float2*>(ssrc + off));
}

template <int OPL, int MINB>
__global__ void __launch_bounds__(kBaWarps * 32, MINB)
ba_residual_jacobian_kernel(const __grid_constant__ BaParams P) {
  using Cfg = BaCfg<OPL>;
  constexpr int kBaOpl = OPL, kBaTile = Cfg::kTile, kBaStage = Cfg::kStage, kBaRingStages = Cfg::kRingStages,
                kBaObsBytes = Cfg::kObsBytes, kBaStJc = Cfg::kStJc, kBaStJp = Cfg::kStJp, kBaStR = Cfg::kStR,
                kBaSmemRing = Cfg::kSmemRing, kBaSmemCam = Cfg::kSmemCam, kBaSmemBars = Cfg::kSmemBars,
                kBaSmemCost = Cfg::kSmemCost;
  extern __shared__ __align__(128) char ba_smem[];
  constexpr int kBaStageBufs = Cfg::kStageBufs;
  char (*stage)[kBaStageBufs * kBaStage] = reinterpret_cast<char (*)[kBaStageBufs * kBaStage]>(ba_smem);
  char (*obs_ring)[kBaRingStages * kBaObsBytes] =
      reinterpret_cast<char (*)[kBaRingStages * kBaObsBytes]>(ba_smem + kBaSmemRing);
  double (*cam_cache)[kCam] = reinterpret_cast<double (*)[kCam]>(ba_smem + kBaSmemCam);
  uint64_t (*ring_bars)[kBaRingStages] = reinterpret_cast<uint64_t (*)[kBaRingStages]>(ba_smem + kBaSmemBars);
  double* warp_cost = reinterpret_cast<double*>(ba_smem + kBaSmemCost);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* const st_base = stage[warp];
  const double* C = cam_cache[warp];
  double cost = 0.0;
  int cached_cam = -1;

  // Every warp owns a CONTIGUOUS range of 64-observation tiles (observations
  // are sorted by camera, so consecutive tiles share the camera block and the
  // 320 B cache is refilled once per ~60 tiles, not once per tile).
  const long long ntiles = (P.n_obs + kBaTile - 1) / kBaTile;
  const long long nwarps = (long long)gridDim.x * kBaWarps;
  const long long gw = (long long)blockIdx.x * kBaWarps + warp;
  const long long t_begin = gw * ntiles / nwarps, t_end = (gw + 1) * ntiles / nwarps;
  const int my_tiles = (int)(t_end - t_begin);

  // Observation records stream through a warp-private 4-stage TMA ring (1 KB per
  // tile, three tiles in flight, no registers held); the dependent point gathers
  // (L2-resident) are issued one tile ahead into ping-pong registers.
  WarpRing<kBaObsBytes, kBaRingStages> ring;
  ring.init(obs_ring[warp], ring_bars[warp], lane);
  const char* obs_b = reinterpret_cast<const char*>(P.obs);
  auto issue = [&](int k) {        // lane 0
    const long long o0 = (t_begin + k) * kBaTile;
    const long long rem = P.n_obs - o0;
    if (rem >= kBaTile) ring.issue(k, obs_b + o0 * 16);
    else ring.issue_bytes(k, obs_b + o0 * 16, (uint32_t)rem * 16u);
  };
  if (lane == 0)
    for (int k = 0; k < kBaRingStages - 1 && k < my_tiles; ++k) issue(k);

  // record of observation slot j*32+lane of tile k (zero beyond n_obs)
  auto read_obs = [&](int k, int j) -> uint4 {
    const long long i = (t_begin + k) * kBaTile + j * 32 + lane;
    uint4 v = *reinterpret_cast<const uint4*>(ring.slot(k) + (j * 32 + lane) * 16);
    if (i >= P.n_obs) v = make_uint4(0, 0, 0, 0);
    return v;
  };
  auto gather = [&](int k, float (&pt)[kBaOpl][3]) {
#pragma unroll
    for (int j = 0; j < kBaOpl; ++j) ba_load_point(P, read_obs(k, j).y, pt[j]);
  };
  auto load_cam = [&](int c) {
    if (c != cached_cam) {
      __syncwarp();
      if (lane < kCam / 2)
        reinterpret_cast<double2*>(cam_cache[warp])[lane] =
            __ldcg(reinterpret_cast<const double2*>(P.cams + (long long)c * kCam) + lane);
      cached_cam = c;
      __syncwarp();
    }
  };

  // one tile: `pt` holds its points on entry and the next tile's on exit.  The
  // points are converted to fp64 FIRST and only then are the next tile's gathers
  // issued into the same registers: issued the other way round, the conversions
  // wait on a scoreboard shared with the brand-new loads (11 % of the stall samples).
  float pt[kBaOpl][3];
  auto process = [&](int k) {
    const long long o0 = (t_begin + k) * kBaTile;
    if (lane == 0 && k + kBaRingStages - 1 < my_tiles) issue(k + kBaRingStages - 1);
    double Xd[kBaOpl][3];
#pragma unroll
    for (int j = 0; j < kBaOpl; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Xd[j][i] = (double)pt[j][i];
        asm volatile("" : "+d"(Xd[j][i]));
      }
    if (k + 1 < my_tiles) {
      ring.wait(k + 1);
      gather(k + 1, pt);
    }
    uint4 ob[kBaOpl];
    int cam[kBaOpl], slot[kBaOpl];
    bool live[kBaOpl];
    bool same = true;
#pragma unroll
    for (int j = 0; j < kBaOpl; ++j) ob[j] = read_obs(k, j);
    const int c0 = __shfl_sync(0xffffffffu, (int)ob[0].x, 0);
#pragma unroll
    for (int j = 0; j < kBaOpl; ++j) {
      slot[j] = j * 32 + lane;
      live[j] = o0 + slot[j] < P.n_obs;
      cam[j] = live[j] ? (int)ob[j].x : c0;
      same = same && (cam[j] == c0);
      cam[j] = cam[j] < 0 ? 0 : (cam[j] >= P.n_cams ? P.n_cams - 1 : cam[j]);
    }
    const bool uniform = __all_sync(0xffffffffu, same) && c0 >= 0 && c0 < P.n_cams;

    // the bulk stores that last read this stage buffer (tile k - kStageBufs) must have drained it
    // before it is rewritten; the stores of the tiles in between may still be in flight
    char* const st = st_base + (k % kBaStageBufs) * kBaStage;
    if (lane == 0) tma_store_wait_read_but<kBaStageBufs - 1>();
    __syncwarp();

    if (uniform) {
      // the 64 records share one camera: every coefficient of its block is
      // read from shared memory once for both observations of the lane
      load_cam(c0);
      bool act[kBaOpl];
#pragma unroll
      for (int j = 0; j < kBaOpl; ++j) act[j] = true;
      ba_eval<kBaOpl, kBaTile>(C, P, Xd, ob, slot, act, live, st, cost);
    } else {
      // a tile that straddles cameras takes one pass per distinct camera of each 32-record row
#pragma unroll
      for (int j = 0; j < kBaOpl; ++j) {
        unsigned todo = 0xffffffffu;
        while (todo) {
          const int c = __shfl_sync(0xffffffffu, cam[j], __ffs(todo) - 1);
          load_cam(c);
          const double p1[1][3] = {{Xd[j][0], Xd[j][1], Xd[j][2]}};
          const uint4 o1[1] = {ob[j]};
          const int s1[1] = {slot[j]};
          const bool a1[1] = {(cam[j] == c) && ((todo >> lane) & 1u)};
          const bool l1[1] = {live[j]};
          ba_eval<1, kBaTile>(C, P, p1, o1, s1, a1, l1, st, cost);
          todo &= ~__ballot_sync(0xffffffffu, a1[0]);
        }
      }
    }

    // ---- stage -> global.  Full tiles leave as three TMA bulk stores issued by
    // lane 0 (the LSU never touches the outputs again); the ragged last tile is
    // copied with plain stores.
    const long long rem = P.n_obs - o0;
    if (rem >= kBaTile) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (P.Jc != nullptr) tma_store_1d(P.Jc + o0 * 12, st + kBaStJc, kBaTile * 48);
        if (P.Jp != nullptr) tma_store_1d(P.Jp + o0 * 6, st + kBaStJp, kBaTile * 24);
        if (P.r != nullptr) tma_store_1d(P.r + o0 * 2, st + kBaStR, kBaTile * 8);
        tma_store_commit();
      }
    } else {
      __syncwarp();
      const int nlive = (int)rem;
      if (P.Jc != nullptr) ba_copy_out(P.Jc + o0 * 12, st + kBaStJc, nlive * 48, lane);
      if (P.Jp != nullptr) ba_copy_out(P.Jp + o0 * 6, st + kBaStJp, nlive * 24, lane);
      if (P.r != nullptr) ba_copy_out(P.r + o0 * 2, st + kBaStR, nlive * 8, lane);
      __syncwarp();
    }
  };

  if (my_tiles > 0) {
    ring.wait(0);
    gather(0, pt);
  }
  if (P.cams_w != nullptr && my_tiles > 0) {
    // Fused precompute: the observations are sorted by camera, so the tiles of THIS WARP
    // reference the cameras [its first record's, its last record's] - one or two at C4 - and
    // the warp derives those blocks now (lane-parallel Rodrigues) while the record tiles and
    // point gathers issued above are in flight: the first one straight into its shared-memory
    // cache, all of them into cam_blocks (read back by this same warp when its camera changes;
    // other warps may write the same block with the same bits).  No second kernel, no launch
    // boundary, no CTA barrier in the step.
    const long long oa = t_begin * kBaTile;
    long long ob = t_end * kBaTile;
    ob = (ob < P.n_obs ? ob : P.n_obs) - 1;
    int ca = __ldg(reinterpret_cast<const int*>(P.obs + oa)), cb = __ldg(reinterpret_cast<const int*>(P.obs + ob));
    ca = ca < 0 ? 0 : (ca >= P.n_cams ? P.n_cams - 1 : ca);
    cb = cb < 0 ? 0 : (cb >= P.n_cams ? P.n_cams - 1 : cb);
    for (int c = ca; c <= cb; ++c)
      ba_camera_block_warp(P.rvec, P.tvec, c, P.cams_w + (long long)c * kCam, c == ca ? cam_cache[warp] : nullptr, lane);
    cached_cam = ca;
    __syncwarp();                 // the warp's blocks (shared and global) are visible to all of its lanes
  }
  // Launched behind the stand-alone camera precompute with programmatic stream serialization
  // (two-kernel route), everything above - barrier init, the first record tiles, the first
  // point gathers, none of which the precompute writes - has overlapped its tail; the camera
  // blocks are complete and visible after this.  A no-op for an ordinary launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int k = 0; k < my_tiles; ++k) process(k);
  // the stage must outlive the bulk stores that read it; their completion in global memory is
  // the kernel boundary's business, so the cost chain below does not wait for it
  if (lane == 0) tma_store_wait_read();

  // cost: warp -> CTA -> partial
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
  if (lane == 0) warp_cost[warp] = cost;
  __syncthreads();
  if (P.partials == nullptr) return;
  __shared__ int s_last;
  unsigned long long* arrived = reinterpret_cast<unsigned long long*>(P.partials + gridDim.x);
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kBaWarps; ++w) s += warp_cost[w];
    P.partials[blockIdx.x] = s;
    __threadfence();
    s_last = atomicAdd(arrived, 1ull) == gridDim.x - 1;
  }
  __syncthreads();
  // The last CTA to finish adds the per-CTA partials in a FIXED order (lane-strided
  // sums, then a shuffle tree): deterministic, no floating-point atomics, and no
  // second launch for an 8-byte result.
  if (s_last && warp == 0 && P.cost != nullptr) {
    __threadfence();
    double s = 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) s += __ldcg(P.partials + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (P.world > 1) {
      // Fused sum-all-reduce over NVLink peer memory.  Lane p stores this rank's cost into rank
      // p's exchange buffer as TWO SELF-VALIDATING words, [tag : 32 | half of the double : 32]
      // with tag = the call's sequence number, in one 16-byte store; lane q then polls rank q's
      // pair in the LOCAL buffer until both tags are this call's.  No fence and no separate flag:
      // one NVLink crossing per call instead of a store, a system fence (a round trip) and a
      // release store.  All ranks add the world values in rank order, so the result is
      // bit-identical everywhere.  Slots alternate with the call parity: a rank can only reuse
      // a slot after every peer has finished the call in between, i.e. has read the previous pair.
      // The call sequence is either the caller's (identical on all ranks) or the device-resident
      // counter behind the arrival word, which lets a captured CUDA graph replay the call.
      unsigned long long* dev_seq = arrived + 1;
      const unsigned long long seq = P.seq != 0 ? P.seq : (*dev_seq + 1ull);
      __syncwarp();
      if (lane == 0 && P.seq == 0) *dev_seq = seq;
      const int par = (int)(seq & 1ull);
      const unsigned long long tag = (seq & 0xffffffffull) << 32;
      if (lane < P.world) {
        unsigned long long* slot = P.peers[lane] + (size_t)(par * P.world + P.rank) * 2;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
        asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(tag | (bits & 0xffffffffull)),
                     "l"(tag | (bits >> 32)) : "memory");
      }
      double v = 0.0;
      if (lane < P.world) {
        const unsigned long long* mine = P.peers[P.rank] + (size_t)(par * P.world + lane) * 2;
        unsigned long long w0 = 0, w1 = 0;
        const long long t0 = clock64();
        bool ok;
        do {
          asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(mine) : "memory");
          ok = ((w0 ^ tag) >> 32) == 0 && ((w1 ^ tag) >> 32) == 0;
        } while (!ok && clock64() - t0 < P.timeout_clk);                    // bounded instead of a hang
        unsigned long long bits = 0x7ff8000000000000ull;                    // NaN if a peer never arrives
        if (ok) bits = (w0 & 0xffffffffull) | (w1 << 32);
        else raise_status(P.status, kStatusPeerTimeout);                    // the cost is NaN and the host raises
        v = __longlong_as_double((long long)bits);
      }
      s = 0.0;
      for (int q = 0; q < P.world; ++q) s += __shfl_sync(0xffffffffu, v, q);
    }
    if (lane == 0) *P.cost = s;
  }
  if (s_last && threadIdx.x == 0) *arrived = 0ull;      // ready for the next call
}

// Kernel shape: 2 observations per lane, 2 CTAs per SM (128 registers).  Measured
// alternatives on C4 (4M observations): <2,3> 80 registers with spills 148 us,
// <1,3> 93 us, <1,4> 107 us against 90 us for this one (3-launch step).
struct BaLaunch {
  void (*kernel)(BaParams);
  int smem;
  int tile;
};
static BaLaunch ba_variant() {
  return {ba_residual_jacobian_kernel<2, 2>, BaCfg<2>::kSmemBytes, BaCfg<2>::kTile};
}

// resident grid of the streaming kernel; < 0 on failure (shared-memory opt-in and occupancy
// are resolved once per device)
static int ba_grid(const DeviceInfo* di, long long n_obs) {
  const BaLaunch L = ba_variant();
  const long long ntiles = (n_obs + L.tile - 1) / L.tile;
  static KernelCache kc;
  int per_sm;
  if (kc.setup(di, L.kernel, kBaWarps * 32, L.smem, &per_sm) != PBK_OK) return -1;
  return resident_grid(di, per_sm, (ntiles + kBaWarps - 1) / kBaWarps);
}

}  // namespace pbk

extern "C" {

int pbk_ba_camera_precompute(const float* rvec, const float* tvec, int n_cams, double* cam_blocks,
                             pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n_cams >= 0, PBK_EINVAL, "negative camera count");
  if (n_cams == 0) return PBK_OK;
  PBK_REQUIRE(rvec && tvec && cam_blocks, PBK_EINVAL, "NULL device pointer");
  ba_camera_precompute_kernel<<<(n_cams + 3) / 4, 128, 0, as_stream(stream)>>>(rvec, tvec, n_cams, cam_blocks);
  PBK_CHECK_LAUNCH("ba_camera_precompute_kernel");
  return PBK_OK;
}

int pbk_ba_cost_slots(int64_t n_obs) {
  using namespace pbk;
  const DeviceInfo* di;
  if (current_device_info(&di)) return -1;
  const int g = ba_grid(di, n_obs > 0 ? n_obs : 1);
  return g < 0 ? g : g + 2;            // + the arrival counter + the device-resident call sequence
}

// fuse_rvec / fuse_tvec non-NULL: the streaming kernel derives the camera blocks itself (sorted
// observations); after_precompute: it was launched behind the stand-alone precompute kernel
static int ba_launch(const void* obs, int64_t n_obs, const float* points, int64_t n_points,
                     double* cam_blocks, int n_cams, double fx, double fy, double cx, double cy, float* r,
                     float* J_cam, float* J_pt, double* cost_partials, double* cost, const uint64_t* peer_slots,
                     int rank, int world, uint64_t seq, uint32_t* status, pbk_stream_t stream,
                     const float* fuse_rvec = nullptr, const float* fuse_tvec = nullptr,
                     bool after_precompute = false) {
  using namespace pbk;
  PBK_REQUIRE(n_obs >= 0 && n_points >= 0 && n_cams >= 0, PBK_EINVAL, "negative size");
  PBK_REQUIRE((cost == nullptr) || (cost_partials != nullptr), PBK_EINVAL, "cost needs cost_partials");
  PBK_REQUIRE(world >= 1 && world <= kBaMaxWorld && rank >= 0 && rank < world, PBK_EINVAL, "bad rank/world");
  PBK_REQUIRE(world == 1 || (peer_slots != nullptr && cost != nullptr), PBK_EINVAL,
              "the fused exchange needs peer buffers and a cost output");
  cudaStream_t s = as_stream(stream);
  if (n_obs == 0 && world == 1) {
    if (cost != nullptr) PBK_CUDA(cudaMemsetAsync(cost, 0, 8, s));
    return PBK_OK;
  }
  PBK_REQUIRE(n_obs > 0, PBK_EINVAL, "a rank of a sharded problem needs at least one observation");
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
  P.rvec = fuse_rvec; P.tvec = fuse_tvec; P.cams_w = (fuse_rvec != nullptr) ? cam_blocks : nullptr;
  P.r = r; P.Jc = J_cam; P.Jp = J_pt; P.partials = cost_partials; P.cost = cost;
  P.n_obs = n_obs; P.n_points = n_points; P.n_cams = n_cams;
  P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
  P.rank = rank; P.world = world; P.seq = seq;
  P.timeout_clk = exchange_timeout_clk();
  P.status = status;
  for (int p = 0; p < kBaMaxWorld; ++p)
    P.peers[p] = (world > 1 && p < world) ? reinterpret_cast<unsigned long long*>(peer_slots[p]) : nullptr;
  const int grid = ba_grid(di, n_obs);
  const BaLaunch L = ba_variant();
  PBK_REQUIRE(grid > 0, PBK_ECUDA, "cannot opt in to %d bytes of shared memory", L.smem);
  if (after_precompute) {
    // may begin while the precompute kernel in front of it is still running (see the kernel)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kBaWarps * 32);
    cfg.dynamicSmemBytes = L.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PBK_CUDA(cudaLaunchKernelEx(&cfg, L.kernel, P));
  } else {
    L.kernel<<<grid, kBaWarps * 32, L.smem, s>>>(P);
  }
  PBK_CHECK_LAUNCH("ba_residual_jacobian_kernel");
  return PBK_OK;
}

int pbk_ba_residual_jacobian(const void* obs, int64_t n_obs, const float* points, int64_t n_points,
                             const double* cam_blocks, int n_cams, double fx, double fy, double cx,
                             double cy, float* r, float* J_cam, float* J_pt, double* cost_partials,
                             double* cost, pbk_stream_t stream) {
  return ba_launch(obs, n_obs, points, n_points, const_cast<double*>(cam_blocks), n_cams, fx, fy, cx, cy, r, J_cam,
                   J_pt, cost_partials, cost, nullptr, 0, 1, 0, nullptr, stream);
}

int pbk_ba_evaluate(const void* obs, int64_t n_obs, int obs_sorted, const float* points, int64_t n_points,
                    const float* rvec, const float* tvec, int n_cams, double* cam_blocks, double fx, double fy,
                    double cx, double cy, float* r, float* J_cam, float* J_pt, double* cost_partials, double* cost,
                    const uint64_t* peer_slots, int rank, int world, uint64_t seq, uint32_t* status,
                    pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(n_cams == 0 || (rvec && tvec), PBK_EINVAL, "NULL device pointer");
  const uint64_t* peers = world > 1 ? peer_slots : nullptr;
  if (obs_sorted && n_cams > 0 && n_obs > 0)
    // ONE kernel: every CTA derives the camera blocks its (camera-sorted) observations reference
    return ba_launch(obs, n_obs, points, n_points, cam_blocks, n_cams, fx, fy, cx, cy, r, J_cam, J_pt,
                     cost_partials, cost, peers, rank, world, seq, status, stream, rvec, tvec);
  int rc = pbk_ba_camera_precompute(rvec, tvec, n_cams, cam_blocks, stream);
  if (rc) return rc;
  return ba_launch(obs, n_obs, points, n_points, cam_blocks, n_cams, fx, fy, cx, cy, r, J_cam, J_pt,
                   cost_partials, cost, peers, rank, world, seq, status, stream, nullptr, nullptr,
                   /*after_precompute=*/n_cams > 0);
}

int pbk_ba_residual_jacobian_dist(const void* obs, int64_t n_obs, const float* points, int64_t n_points,
                                  const double* cam_blocks, int n_cams, double fx, double fy, double cx,
                                  double cy, float* r, float* J_cam, float* J_pt, double* cost_partials,
                                  double* cost, const uint64_t* peer_slots, int rank, int world, uint64_t seq,
                                  uint32_t* status, pbk_stream_t stream) {
  return ba_launch(obs, n_obs, points, n_points, const_cast<double*>(cam_blocks), n_cams, fx, fy, cx, cy, r, J_cam,
                   J_pt, cost_partials, cost, peer_slots, rank, world, seq, status, stream);
}

}  // extern "C"
