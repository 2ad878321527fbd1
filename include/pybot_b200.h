/*
 * pybot_b200.h - C-ABI of libpybot_b200.so: the B200 (sm_100a) implementation of
 * the data-parallel geometry-and-matching hot path of spillai/pybot.
 *
 * The reference has no FFI for this path: its "operator API" is the Python
 * method surface, which bottoms out in numpy BLAS and three OpenCV entry
 * points.  Each entry point below names the reference site (file:line under
 * /root/reference) whose arithmetic it replaces; pybot_b200/_ffi.py is the
 * ctypes binding and INTEGRATION.md shows the stub a pybot maintainer adds.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer documented "device" must be
 *    a CUDA device pointer valid in the calling thread's current context;
 *    "host" pointers are read before the call returns.
 *  - all launches are asynchronous on `stream` (a cudaStream_t / CUstream
 *    handle cast to void*; NULL = legacy default stream).
 *  - the library never allocates or frees device memory and keeps no pointers
 *    after a call returns; all outputs and scratch are caller-allocated.
 *  - every function returns PBK_OK (0) or a negative pbk_status; the message
 *    of the last failure on the calling thread is pbk_last_error_string().
 *  - device pointers must be 16-byte aligned (PBK_EALIGN otherwise).
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    returns PBK_ECUDA.
 */
#ifndef PYBOT_B200_H_
#define PYBOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBK_VERSION 200

#if defined(__GNUC__)
#define PBK_API __attribute__((visibility("default")))
#else
#define PBK_API
#endif

typedef enum pbk_status {
  PBK_OK = 0,
  PBK_EINVAL = -1,       /* bad argument (size, null pointer, enum)         */
  PBK_ECUDA = -2,        /* CUDA runtime error; see pbk_last_error_string   */
  PBK_EALIGN = -3,       /* device pointer not 16-byte aligned              */
  PBK_EUNSUPPORTED = -4  /* valid in the reference but not implemented here */
} pbk_status;

typedef enum pbk_dtype {
  PBK_U8 = 0, PBK_S16 = 1, PBK_S32 = 2, PBK_F32 = 3, PBK_F64 = 4, PBK_U16 = 5
} pbk_dtype;

typedef void* pbk_stream_t;

/* ---------------------------------------------------------------- runtime */
PBK_API int pbk_version(void);
PBK_API const char* pbk_last_error_string(void);
/* Bound of the cross-GPU waits of the fused exchanges (pbk_hamming_*_dist,
 * pbk_ba_*: world > 1), default ~2 s.  A peer that does not arrive within the bound
 * makes the call fail: its `status` word is set to PBK_STATUS_PEER_TIMEOUT and the
 * outputs are poisoned (no-match keys, zero rows, NaN cost) - never stale data.   */
#define PBK_STATUS_PEER_TIMEOUT 1u
PBK_API int pbk_set_exchange_timeout_ms(double ms);
/* props[0]=SM count, [1]=cc major, [2]=cc minor, [3]=max smem/block optin,
 * [4]=L2 bytes, [5]=SM clock kHz, [6]=memory clock kHz, [7]=bus width bits */
PBK_API int pbk_device_props(int device, int64_t props[8]);

/* HBM bandwidth probe (measurement only).  mode 0: read `bytes` from src;
 * 1: write `bytes` to dst; 2: copy src -> dst.  16 B vector accesses, grid =
 * 8 CTAs per SM.  sink: device uint32[1].                                      */
PBK_API int pbk_hbm_probe(int mode, const void* src, void* dst, int64_t bytes,
                          uint32_t* sink, pbk_stream_t stream);

/* ------------------------------------------------- disparity -> XYZ (a6,a7)
 * Replaces cv2.reprojectImageTo3D(disp, self.Q) in StereoCamera.reconstruct,
 * pybot/vision/camera_utils.py:964-969, and reconstruct_with_texture :982-990
 * (including its two np.copy(...[::s, ::s]).reshape(-1,3) passes :988-989).
 *
 * disp   device, (batch, H, W) contiguous, dtype in {U8,S16,S32,F32}
 *        (float64 is rejected like OpenCV rejects it: PBK_EUNSUPPORTED)
 * Q      host, 16 floats row-major (StereoCamera.Q is float32, :893)
 * xyz    device, (batch, Hs, Ws, 3) float32, Hs=ceil(H/sample), Ws=ceil(W/sample)
 * bgr    device or NULL, (batch, H, W, 3) uint8; bgr_out (batch, Hs, Ws, 3)
 * Per pixel, bit-exact with OpenCV 4.13: h = Q*[x,y,d,1] in fp64 without FMA,
 * out = float( double(float(h_i)) * (1.0/h_3) ).
 */
PBK_API int pbk_reproject_disp(const void* disp, int disp_dtype, int batch, int H, int W,
                       const float* Q, int sample, float* xyz,
                       const uint8_t* bgr, uint8_t* bgr_out,
                       pbk_stream_t stream);

/* StereoCamera.reconstruct_sparse, camera_utils.py:971-980: float64
 * (Q * [x,y,d,1])[:3] / W per row of xyd (N,3) float64 -> out (N,3) float64. */
PBK_API int pbk_reproject_sparse(const double* xyd, int64_t n, const float* Q,
                         double* out, pbk_stream_t stream);

/* ---------------------------------------- SE(3)/Sim(3) point transform (a1,a3,a5)
 * Replaces np.hstack + np.dot(self.matrix, X) in RigidTransform.__mul__,
 * pybot/geometry/rigid_transform.py:139-141 (and CameraExtrinsic.c2w/w2c,
 * camera_utils.py:585-595).  Y = s*R*X + t (s = 1 for SE(3); Sim(3) semantics
 * per SURVEY.md A.6).
 *
 * X      device, (n,3) AoS, in_dtype in {F32,F64}
 * Rt     host, 12 doubles: R row-major (9) then t (3); scale host double
 * Y      device, (n,3) AoS, out_dtype in {F32,F64} (reference returns F64)
 */
PBK_API int pbk_transform_points(const void* X, int in_dtype, int64_t n,
                         const double* Rt, double scale,
                         void* Y, int out_dtype, pbk_stream_t stream);

/* ------------------------------------------------------ Camera.project (a4,a5)
 * Replaces cv2.projectPoints + depth_from_projection + the depth / bounds
 * masks + the x[valid] compaction of Camera.project,
 * pybot/vision/camera_utils.py:687-734, in one pass.
 */
typedef struct pbk_project_params {
  double R[9];        /* rotation used for the pixel (cv2.Rodrigues round trip) */
  double t[3];
  double Rz[3];       /* third row of the rotation used for the depth          */
  double tz;          /*   (extrinsics round trip, camera_utils.py:683-685)     */
  double fx, fy, cx, cy;
  double k[8];        /* k1 k2 p1 p2 k3 k4 k5 k6 ; all zero -> pinhole only     */
  double min_depth;
  int32_t width, height;
  int32_t check_depth;   /* valid &= depth >= min_depth        (:708-710)      */
  int32_t check_bounds;  /* valid &= 0<=u<W && 0<=v<H          (:714-725)      */
  int32_t has_distortion;
  int32_t compact;       /* 1: uv/depth outputs are x[valid] order-preserving  */
} pbk_project_params;

/* X        device (n,3) AoS, in_dtype in {F32,F64}
 * uv       device (n,2) same dtype as X (cv2 semantics), compacted if compact
 * depth    device (n,) float64 or NULL
 * valid    device (n,) uint8 or NULL (never compacted)
 * count    device int64[1] or NULL: number of valid points (needed if compact)
 * scratch  device, >= pbk_project_scratch_bytes(n, in_dtype) bytes, only if compact:
 *          64 B of control words, the per-CTA counts and the L2-resident region the
 *          classified points of one chunk are compacted into before they are projected
 *          (two phases in one cooperative kernel; see csrc/transform_project.cu).
 *          The environment variable PBK_COMPACT_CHUNK_MB (default 40) sets the chunk size.
 */
PBK_API size_t pbk_project_scratch_bytes(int64_t n, int in_dtype);
PBK_API int pbk_project_points(const void* X, int in_dtype, int64_t n,
                       const pbk_project_params* params,
                       void* uv, double* depth, uint8_t* valid,
                       int64_t* count, void* scratch, pbk_stream_t stream);

/* ------------------------------------------------------- Hamming kNN (a9)
 * Replaces cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, k=2) over a collection
 * (no call site in the reference: SURVEY.md 0.3; semantics pinned against
 * OpenCV 4.13).  256-bit descriptors, 32 bytes per row.
 *
 * Result keys are (distance << 32) | global_train_row, so an unsigned 64-bit
 * min reproduces BFMatcher's lowest-train-index tie-break across tiles,
 * splits and GPUs.  PBK_KEY_NONE marks "fewer than k neighbours".
 */
#define PBK_KEY_NONE 0xFFFFFFFFFFFFFFFFull

/* number of train splits the kernel will use for (nq, nt) on this device, and
 * the scratch it needs: partial keys (nq, splits, 2) uint64                    */
PBK_API int pbk_hamming_plan(int64_t nq, int64_t nt, int* splits, size_t* scratch_bytes);

/* q (nq,32) u8 device; t (nt,32) u8 device; train_base = global row index of
 * t[0] (database shard offset); keys (nq,2) uint64 device out;
 * scratch from pbk_hamming_plan.                                              */
PBK_API int pbk_hamming_knn2(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt,
                     uint64_t train_base, uint64_t* keys, void* scratch,
                     pbk_stream_t stream);

/* merge `parts` partial results laid out (parts, nq, 2) -> keys (nq,2): the
 * step after an all-gather of per-shard keys.                                 */
PBK_API int pbk_hamming_merge_top2(const uint64_t* partial, int parts, int64_t nq,
                           uint64_t* keys, pbk_stream_t stream);

/* rows[i] = t[(keys[i,0] & 0xffffffff) - train_base] if that row lies in
 * [train_base, train_base+nt), else 32 zero bytes; rows (nq,32) u8.           */
PBK_API int pbk_hamming_gather_best(const uint8_t* t, int64_t nt, uint64_t train_base,
                            const uint64_t* keys, int64_t nq, uint8_t* rows,
                            pbk_stream_t stream);

/* Lowe ratio + mutual check.  rev_keys (nq,2): result of pbk_hamming_knn2 with
 * the gathered best rows as queries and q as train (only [:,0] is used).
 * ratio_ok[i]  = has 2 neighbours && (double)d1 < ratio*(double)d2
 * mutual_ok[i] = has 1 neighbour  && (rev_keys[i,0] & 0xffffffff) == i
 * valid = ratio_ok & mutual_ok ; idx (nq,2) int64 (-1 = none), dist (nq,2) int32 */
PBK_API int pbk_hamming_ratio_mutual(const uint64_t* keys, const uint64_t* rev_keys,
                             int64_t nq, double ratio, int64_t* idx,
                             int32_t* dist, uint8_t* ratio_ok,
                             uint8_t* mutual_ok, uint8_t* valid,
                             pbk_stream_t stream);

/* The rest of cv2.BFMatcher.knnMatch(query, train, k, mask)'s signature: any
 * 1 <= k <= PBK_HAMMING_KNNK_MAX_K and an optional per-pair mask, in BFMatcher's order
 * (ascending distance, ties to the lowest global train row).  Not the hot path (k = 2 without
 * a mask goes through pbk_hamming_knn2 / pbk_hamming_tc_knn2); one CTA per query, two passes.
 * mask   NULL, or (nq, nt) uint8 row-major on the device: pair (i, j) is a candidate iff
 *        mask[i*nt + j] != 0 (OpenCV's convention)
 * idx    (nq, k) int64 out: train_base + row, -1 past the neighbours found
 * dist   (nq, k) int32 out, -1 past the neighbours found
 * count  (nq)    int32 out: min(k, allowed rows of that query)                          */
#define PBK_HAMMING_KNNK_MAX_K 1024
PBK_API int pbk_hamming_knnk(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int k,
                     const uint8_t* mask, uint64_t train_base, int64_t* idx, int32_t* dist,
                     int32_t* count, pbk_stream_t stream);

/* Sharded matching with the exchanges FUSED into the kernels: the
 * split-merge kernel stores this rank's top-2 keys into every peer's exchange
 * buffer over NVLink peer memory, the gather kernel does the same with the
 * matched train rows it owns; the consumers wait for the peers' sequence
 * numbers.  No NCCL call on the data path; results are bit-identical to
 * pbk_hamming_knn2 over the whole database.
 * peer_bufs    host array of `world` device pointers: peer_bufs[p] = rank p's
 *              exchange buffer AS MAPPED INTO THIS PROCESS, each >=
 *              pbk_hamming_exchange_words(nq, world) uint64, zeroed before the
 *              first call
 * seq          call counter, identical on all ranks, 1, 2, 3 ... ; one value is
 *              shared by the knn2_dist and gather_best_dist calls of a frame
 * done_counter device uint32[1], zero before the first call
 * status       uint32[1] the kernels can write (device memory, or pinned host memory
 *              mapped into the device so the host can poll it without a copy), zero
 *              before the first call; set to PBK_STATUS_PEER_TIMEOUT when a peer's data
 *              did not arrive within the bound - the call's outputs are then all
 *              "no match" (keys PBK_KEY_NONE, rows zero).  May be NULL.
 * knn2_dist:        keys (nq,2) = GLOBAL top-2 on every rank
 * gather_best_dist: rows (nq,32) = the matched train rows on every rank          */
PBK_API size_t pbk_hamming_exchange_words(int64_t nq, int world);
PBK_API int pbk_hamming_knn2_dist(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt,
                     uint64_t train_base, uint64_t* keys, void* scratch,
                     const uint64_t* peer_bufs, int rank, int world, uint64_t seq,
                     uint32_t* done_counter, uint32_t* status, pbk_stream_t stream);
PBK_API int pbk_hamming_gather_best_dist(const uint8_t* t, int64_t nt, uint64_t train_base,
                     const uint64_t* keys, int64_t nq, uint8_t* rows,
                     const uint64_t* peer_bufs, int rank, int world, uint64_t seq,
                     uint32_t* done_counter, uint32_t* status, pbk_stream_t stream);

/* Tensor-core route of the same scan: tcgen05.mma over descriptors expanded ONCE into a signed
 * element per bit, accumulators in tensor memory, the same top-2 keys and merges - results
 * bit-identical to pbk_hamming_knn2 / pbk_hamming_knn2_dist.  Two encodings (pbk_hamming_tc_format):
 *   4 (default)  e2m1 nibbles, tcgen05.mma kind::mxf4 with unit block scales (csrc/hamming_tc4.cu):
 *                128 B per row, fp32 accumulator = 2 d - 256, twice the int8 MMA rate
 *   8            signed bytes, tcgen05.mma kind::i8 (csrc/hamming_tc.cu): 256 B per row,
 *                int32 accumulator = 128 d - 16384
 * The format is a property of the expanded buffers: select it (PBK_TC_FORMAT=4|8 in the
 * environment, or pbk_hamming_tc_set_format) BEFORE expanding and keep it for every scan of that
 * copy.  The expanded copy of a database is made once (pbk_hamming_tc_expand, in the tile layout
 * the kernel's UMMA descriptors describe, padded to a multiple of 512 rows) and kept resident;
 * queries are expanded inside the call.  An appended database is extended in place:
 * first_row = (old n / 512) * 512 rewrites only the rows from there on (d always points at row 0).
 * The scan must be given the whole buffer with the nt it was expanded for.
 * signs     device, >= pbk_hamming_tc_signs_bytes(n) bytes, 16-byte aligned
 * scratch   device, >= the size pbk_hamming_tc_plan reports (partial keys + expanded queries)   */
PBK_API int pbk_hamming_tc_format(void);
PBK_API int pbk_hamming_tc_set_format(int format);
PBK_API size_t pbk_hamming_tc_signs_bytes(int64_t n);
PBK_API int pbk_hamming_tc_expand(const uint8_t* d, int64_t n, int64_t first_row, int8_t* signs,
                                  pbk_stream_t stream);
PBK_API int pbk_hamming_tc_plan(int64_t nq, int64_t nt, int* parts, size_t* scratch_bytes);
PBK_API int pbk_hamming_tc_knn2(const uint8_t* q, int64_t nq, const int8_t* t_signs, int64_t nt,
                     uint64_t train_base, uint64_t* keys, void* scratch, pbk_stream_t stream);
PBK_API int pbk_hamming_tc_knn2_dist(const uint8_t* q, int64_t nq, const int8_t* t_signs, int64_t nt,
                     uint64_t train_base, uint64_t* keys, void* scratch,
                     const uint64_t* peer_bufs, int rank, int world, uint64_t seq,
                     uint32_t* done_counter, uint32_t* status, pbk_stream_t stream);

/* tensor-pipe peak probe (measurement only) for the selected format: one CTA per SM issues
 * iters x (4 tcgen05.mma kind::mxf4, K 64 | 8 tcgen05.mma kind::i8, K 32) of the matcher's shape
 * (M 128, N 128) on resident shared memory, no data movement, no epilogue.  *macs_per_launch
 * receives the multiply-adds issued (iters x 128 x 128 x 256 per SM in either format).
 * sink: device uint32[SM count] or NULL.                                                      */
PBK_API int pbk_hamming_tc_peak_probe(int iters, uint32_t* sink, double* macs_per_launch,
                                      pbk_stream_t stream);

/* popc-pipe peak probe (measurement only): ctas x 256 threads run `iters`
 * rounds of mode 0: independent POPC chains; 1: POPC + one LOP3 each;
 * 2 / 3: the matcher's inner instruction mix (8 / 4 POPC per pair) on
 * registers.  *popc_per_launch receives the number of 32-bit POPC executed
 * (modes 2/3: 16*iters pairs per thread).  sink: device uint32[8 + ctas*256],
 * first 8 words are read as inputs.                                           */
PBK_API int pbk_popc_peak_probe(int mode, int ctas, int iters, uint32_t* sink,
                        double* popc_per_launch, pbk_stream_t stream);

/* ------------------------------------- BA residual + Jacobians (a10)
 * Replaces cv2.projectPoints(X, rvec, tvec, K, D=0) -> (x, J) minus the
 * observation: the Jacobian Camera.project computes and discards at
 * pybot/vision/camera_utils.py:697 (columns 0:6), plus J_pt = J_t * R.
 *
 * ba_camera_precompute: rvec (C,3) f32, tvec (C,3) f32 device ->
 *   cam_blocks (C, PBK_BA_CAM_DOUBLES) f64 device: R(9) t(3) dR/dr(27).
 * ba_residual_jacobian: obs (O,4) {int32 cam, int32 pt, f32 u, f32 v} device,
 *   points (P,3) f32 device -> r (O,2) f32, J_cam (O,2,6) f32, J_pt (O,2,3)
 *   f32 (any may be NULL), cost_partials (pbk_ba_cost_slots(O),) f64 device (per-CTA sums,
 *   an arrival counter and the device-resident call sequence), ZERO on the
 *   first call (the kernel leaves it reusable), cost (1,) f64 device = sum ||r||^2
 *   (deterministic: per-CTA partials added in fixed order by the last CTA).
 */
#define PBK_BA_CAM_DOUBLES 40
PBK_API int pbk_ba_camera_precompute(const float* rvec, const float* tvec, int n_cams,
                             double* cam_blocks, pbk_stream_t stream);
PBK_API int pbk_ba_cost_slots(int64_t n_obs);
PBK_API int pbk_ba_residual_jacobian(const void* obs, int64_t n_obs,
                             const float* points, int64_t n_points,
                             const double* cam_blocks, int n_cams,
                             double fx, double fy, double cx, double cy,
                             float* r, float* J_cam, float* J_pt,
                             double* cost_partials, double* cost,
                             pbk_stream_t stream);

/* Sharded BA: the same kernel with the sum-all-reduce of the cost
 * FUSED in - the last CTA stores this rank's cost into every peer's exchange
 * buffer over NVLink peer memory, waits for all peers' stores and adds the
 * `world` values in rank order (bit-identical on every rank).  No NCCL call on
 * the data path.
 * peer_slots  host array of `world` device pointers: peer_slots[p] is rank p's
 *             exchange buffer AS MAPPED INTO THIS PROCESS (e.g. torch symmetric
 *             memory buffer_ptrs), >= PBK_BA_EXCHANGE_WORDS(world) uint64, zeroed
 *             before the first call
 * seq         call counter, identical on all ranks, 1, 2, 3 ... ; 0 = use the
 *             device-resident counter kept behind cost_partials (every rank then
 *             simply makes the same number of calls; this is what lets a captured
 *             CUDA graph replay the call)
 * status      uint32[1] the kernel can write (device or mapped pinned host memory), zero
 *             before the first call, set to PBK_STATUS_PEER_TIMEOUT on a timeout; may be NULL
 * cost        receives the GLOBAL cost on every rank (NaN if a peer did not
 *             arrive within the bound).  Every rank must own >= 1 observation.  */
#define PBK_BA_EXCHANGE_WORDS(world) (4 * (world))
PBK_API int pbk_ba_residual_jacobian_dist(const void* obs, int64_t n_obs,
                             const float* points, int64_t n_points,
                             const double* cam_blocks, int n_cams,
                             double fx, double fy, double cx, double cy,
                             float* r, float* J_cam, float* J_pt,
                             double* cost_partials, double* cost,
                             const uint64_t* peer_slots, int rank, int world,
                             uint64_t seq, uint32_t* status, pbk_stream_t stream);

/* One call for one evaluation, straight from rvec / tvec.
 * obs_sorted != 0 (the observations are sorted by camera index - the caller checks): ONE
 *   kernel.  Every CTA of the streaming kernel derives, in its prologue, the camera blocks
 *   its own observation range references (its eight warps, one camera each, lane-parallel
 *   Rodrigues) into cam_blocks while its first record tiles are in flight - no precompute
 *   launch, no launch boundary inside the step; only the blocks of referenced cameras are
 *   (re)written.
 * obs_sorted == 0: pbk_ba_camera_precompute followed by the streaming kernel launched with
 *   programmatic stream serialization so that its prologue overlaps the precompute.
 * world == 1: peer_slots NULL, seq / status ignored.  Same arithmetic on both routes (the
 * camera block is one device function), bit-identical outputs.                            */
PBK_API int pbk_ba_evaluate(const void* obs, int64_t n_obs, int obs_sorted,
                             const float* points, int64_t n_points,
                             const float* rvec, const float* tvec, int n_cams,
                             double* cam_blocks,
                             double fx, double fy, double cx, double cy,
                             float* r, float* J_cam, float* J_pt,
                             double* cost_partials, double* cost,
                             const uint64_t* peer_slots, int rank, int world,
                             uint64_t seq, uint32_t* status, pbk_stream_t stream);

/* ------------------------------------------- next rows
 * DepthCamera.reconstruct, pybot/vision/camera_utils.py:1032-1036:
 *   out[b, y, x] = (xs[x]*d, ys[y]*d, d), d = depth[b, y*skip, x*skip], float64.
 * depth  device (batch, H, W), dtype in {U16, F32, F64}
 * xs, ys device float64 meshes of DepthCamera._build_mesh (:1020-1030), already
 *        sub-sampled: Ws = ceil(W/skip), Hs = ceil(H/skip) entries
 * out    device (batch, Hs, Ws, 3) float64                                     */
PBK_API int pbk_depth_reconstruct(const void* depth, int depth_dtype, int batch, int H, int W,
                                  int skip, const double* xs, const double* ys, double* out,
                                  pbk_stream_t stream);

/* DepthCamera.reconstruct_sparse, camera_utils.py:1038-1041 (both axes divided
 * by fx, as the reference does): pts (n,2) f64, depth (n,) f64 -> out (n,3) f64 */
PBK_API int pbk_depth_reconstruct_sparse(const double* pts, const double* depth, int64_t n,
                                         double fx, double cx, double cy, double* out,
                                         pbk_stream_t stream);

/* SceneFlow.process, pybot/vision/optflow_utils.py:129-150: project the (H,W,3)
 * scene points dX with params (R, t, K, D; the reference uses identity
 * extrinsics), cast to int32 like numpy, keep in-bounds hits, scatter the source
 * grid coordinate to the hit pixel (last source wins) and subtract the grid.
 * dX      device (H, W, 3), dtype in {F32, F64}
 * winner  device int32 (H*W) scratch
 * flow    device (H, W, 2) float32; NaN where no point landed                   */
PBK_API int pbk_sceneflow(const void* dX, int dtype, int H, int W,
                          const pbk_project_params* params, int32_t* winner, float* flow,
                          pbk_stream_t stream);

/* check_visibility, pybot/vision/camera_utils.py:245-266, fused: mask[i] = |atan2(v_x, v_z)| <
 * half_fov_h && |atan2(-v_y, v_z)| < half_fov_v && zmin <= z <= zmax with pts_c = R X + t
 * (camera.c2w, the same mul,fma,fma,fma chain as pbk_transform_points), v = pts_c / |pts_c|,
 * z = pts_c[2].  Every operation rounded as numpy rounds it; the two atan2 are CUDA's double
 * atan2 (<= 2 ulp; numpy calls the C library's): a decision can differ from the reference only
 * for an angle within ~1e-15 rad of the field-of-view limit.
 * X device (n,3) F32/F64; Rt host 12 doubles (R row-major, t); half_fov_* = the reference's
 * float32 camera.fov[i] * 0.5 as a double; mask device (n,) uint8 (0 / 1).               */
PBK_API int pbk_fov_mask(const void* X, int in_dtype, int64_t n, const double* Rt,
                         double half_fov_h, double half_fov_v, double zmin, double zmax,
                         uint8_t* mask, pbk_stream_t stream);

/* get_median_depth, camera_utils.py:269-276: z[i] = (T * X[i * stride])[2] for every stride-th
 * point (np.dot chain, float64) - the (camera * pts[::subsample])[:, 2] column alone.
 * Rz_tz host 4 doubles: third row of R, then t_z.  z device (ceil(n / stride),) float64.   */
PBK_API int pbk_transform_depth(const void* X, int in_dtype, int64_t n, int64_t stride,
                                const double* Rz_tz, double* z, pbk_stream_t stream);

/* sampson_error, pybot/vision/camera_utils.py:52-68: F host 9 doubles row-major,
 * pts1 / pts2 device (n,2) float64 -> err (n,) float64 =
 * (x1^T F x2)^2 / ((F x1)_0^2 + (F x1)_1^2 + (F x2)_0^2 + (F x2)_1^2), exactly the
 * reference's expression (it uses F x2, not F^T x2).                           */
PBK_API int pbk_sampson_error(const double* F, const double* pts1, const double* pts2, int64_t n,
                              double* err, pbk_stream_t stream);

/* epipolar_line, pybot/vision/camera_utils.py:352-357
 * (cv2.computeCorrespondEpilines(x, 1, F)): F host 9 doubles row-major, pts device (n,2)
 * float32 / float64 -> lines (n,3) of the same dtype, l = F [x y 1]^T scaled so that
 * a^2 + b^2 = 1 (unscaled when a = b = 0); double arithmetic in OpenCV's order, one
 * rounding to the point dtype.                                                      */
PBK_API int pbk_epipolar_lines(const double* F, const void* pts, int dtype, int64_t n, void* lines,
                               pbk_stream_t stream);

/* triangulate_points, pybot/vision/camera_utils.py:108-116:
 * cv2.triangulatePoints(P1, P2, x1, x2).T dehomogenised.  P1, P2 host 12 doubles
 * (3x4 row-major); x1, x2 device (n,2), dtype in {F32, F64}; out device (n,3) of
 * the same dtype.  One-sided Jacobi SVD per point, ~1e-14 relative of OpenCV.    */
PBK_API int pbk_triangulate_points(const double* P1, const double* P2, const void* x1,
                                   const void* x2, int dtype, int64_t n, void* out,
                                   pbk_stream_t stream);

/* CameraIntrinsic.undistort_points / ray / reconstruct,
 * pybot/vision/camera_utils.py:499-538.                                         */
typedef struct pbk_ray_params {
  double fx, fy, cx, cy;   /* camera matrix K (skew ignored, like OpenCV)                 */
  double k[8];             /* k1 k2 p1 p2 k3 k4 k5 k6                                      */
  double P[9];             /* new camera matrix of cv2.undistortPoints(P=...), row-major   */
  double rot[9];           /* Quaternion.rotate coefficient rows (quaternion.py:69-85):
                            * out_i = 2*(rot[3i]*v0 + rot[3i+1]*v1 + rot[3i+2]*v2) + v_i    */
  int32_t undistort;       /* ray: run undistort_points first                 (:505)       */
  int32_t rotate;          /* ray: extrinsics.rotate_vec                      (:511-512)   */
  int32_t normalize;       /* ray: divide by the Euclidean norm               (:514-515)   */
  int32_t scale_by_z;      /* reconstruct: multiply by pts[:, 2]              (:519-524)   */
} pbk_ray_params;

/* cv2.undistortPoints(pts.astype(float32), K, D, P=P) (:536-537): five fixed-point
 * iterations of the inverse distortion in fp64, then P; float32 out, bit-exact.
 * pts device (n, stride) rows of dtype in {F32, F64}, stride >= 2 elements.      */
PBK_API int pbk_undistort_points(const void* pts, int dtype, int stride, int64_t n,
                                 const pbk_ray_params* params, float* out, pbk_stream_t stream);

/* ray()/reconstruct(): [(u-cx)/fx, (v-cy)/fy, 1] (optionally of the undistorted
 * float32 pixel), rotated, normalised, scaled by the row's third element.
 * out device (n,3) float64.                                                      */
PBK_API int pbk_camera_rays(const void* pts, int dtype, int stride, int64_t n,
                            const pbk_ray_params* params, double* out, pbk_stream_t stream);

/* Point-cloud colours in the viewer wire format: get_color_arr,
 * pybot/externals/draw_helpers.py:35-53, for uint8 (n_px,3) images -> float32
 * (n_px,3) = c / 255.0 with optional B<->R flip; this is the block arr_msg
 * (pybot/externals/pb.py:35) serialises as msg.colors.                          */
PBK_API int pbk_wire_colors(const uint8_t* bgr, int64_t n_px, int flip_rb, float* out,
                            pbk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PYBOT_B200_H_ */
