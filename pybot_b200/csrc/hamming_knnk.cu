// General-k / masked brute-force Hamming kNN (sm_100a): the rest of
// cv2.BFMatcher.knnMatch(query, train, k, mask, compactResult)'s signature.
// The hot path (k = 2, no mask) is hamming_tc4.cu / hamming_tc.cu / hamming.cu; this kernel is
// the complete-signature path next to it and shares their result order: ascending
// (distance, global train row), i.e. BFMatcher's lowest-train-index tie-break.
//
// One CTA per query, two passes over the train rows (640 MB of descriptors stay in the 126 MB L2
// only partly - this path is for per-frame sizes and masked re-matching, not the database scan):
//   pass 1  histogram of the 257 possible distances over the allowed rows (per-warp copies in
//           shared memory) -> the distance d* of the k-th neighbour, the number of rows strictly
//           below it and how many rows AT d* are still needed;
//   pass 2  rows below d* are taken in any order; rows at d* are taken in ROW order (a ballot /
//           prefix rank per 256-row chunk) until the k are filled - so ties resolve to the lowest
//           row exactly as a (distance, row) sort would;
//   then the <= k keys are ranked by counting in shared memory and written in order.
// Distances are recomputed in pass 2 (8 POPC per pair) rather than parked in HBM: a parked
// distance is 2 B against 32 B of descriptor that L2 mostly still holds.
#include "pbk_common.cuh"

namespace pbk {

constexpr int KNNK_THREADS = 256;
constexpr int KNNK_WARPS = KNNK_THREADS / 32;

__device__ __forceinline__ unsigned int knnk_distance(const uint4& qa, const uint4& qb, const uint8_t* row) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(row));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(row) + 1);
  return __popc(qa.x ^ a.x) + __popc(qa.y ^ a.y) + __popc(qa.z ^ a.z) + __popc(qa.w ^ a.w) +
         __popc(qb.x ^ b.x) + __popc(qb.y ^ b.y) + __popc(qb.z ^ b.z) + __popc(qb.w ^ b.w);
}

__global__ void __launch_bounds__(KNNK_THREADS)
hamming_knnk_kernel(const uint8_t* __restrict__ q, long long nq, const uint8_t* __restrict__ t, long long nt,
                    int k, const uint8_t* __restrict__ mask, unsigned long long train_base,
                    long long* __restrict__ idx, int* __restrict__ dist, int* __restrict__ count) {
  __shared__ unsigned int hist[KNNK_WARPS][257];
  __shared__ unsigned long long cand[PBK_HAMMING_KNNK_MAX_K];
  __shared__ unsigned int s_wcnt[2][KNNK_WARPS];
  __shared__ unsigned int s_sel[3];          // d*, rows below d*, neighbours to return
  __shared__ unsigned int s_below;           // rows below d* stored so far
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (long long qi = blockIdx.x; qi < nq; qi += gridDim.x) {
    const uint4 qa = __ldg(reinterpret_cast<const uint4*>(q + qi * 32));
    const uint4 qb = __ldg(reinterpret_cast<const uint4*>(q + qi * 32) + 1);
    const uint8_t* mrow = mask ? mask + qi * nt : nullptr;

    for (int i = tid; i < KNNK_WARPS * 257; i += KNNK_THREADS) (&hist[0][0])[i] = 0u;
    if (tid == 0) s_below = 0u;
    __syncthreads();

    // ---- pass 1: distance histogram of the allowed rows
    for (long long j = tid; j < nt; j += KNNK_THREADS) {
      if (mrow && mrow[j] == 0) continue;
      atomicAdd(&hist[warp][knnk_distance(qa, qb, t + j * 32)], 1u);
    }
    __syncthreads();
    for (int b = tid; b < 257; b += KNNK_THREADS) {
      unsigned int s = 0;
#pragma unroll
      for (int w = 0; w < KNNK_WARPS; ++w) s += hist[w][b];
      hist[0][b] = s;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long total = 0;
      for (int b = 0; b < 257; ++b) total += hist[0][b];
      const unsigned int kk = (unsigned int)(total < (unsigned long long)k ? total : (unsigned long long)k);
      unsigned int dstar = 0, below = 0;
      if (kk > 0) {
        unsigned long long cum = 0;
        for (int b = 0; b < 257; ++b) {
          if (cum + hist[0][b] >= kk) { dstar = (unsigned int)b; below = (unsigned int)cum; break; }
          cum += hist[0][b];
        }
      }
      s_sel[0] = dstar; s_sel[1] = below; s_sel[2] = kk;
    }
    __syncthreads();
    const unsigned int dstar = s_sel[0], below = s_sel[1], kk = s_sel[2];
    const unsigned int need_at = kk - below;      // rows AT d* still to take, in row order

    // ---- pass 2: collect
    unsigned int at_taken = 0;                    // identical in every thread
    int buf = 0;
    for (long long base = 0; base < nt && kk > 0; base += KNNK_THREADS) {
      const long long j = base + tid;
      unsigned int d = 0xFFFFu;
      if (j < nt && !(mrow && mrow[j] == 0)) d = knnk_distance(qa, qb, t + j * 32);
      const unsigned long long key = ((unsigned long long)d << 32) | (train_base + (unsigned long long)j);
      if (d < dstar) cand[atomicAdd(&s_below, 1u)] = key;
      if (at_taken < need_at) {                   // uniform branch
        const bool at = (d == dstar);
        const unsigned int votes = __ballot_sync(0xffffffffu, at);
        if (lane == 0) s_wcnt[buf][warp] = __popc(votes);
        __syncthreads();
        unsigned int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < KNNK_WARPS; ++w) {
          const unsigned int c = s_wcnt[buf][w];
          before += (w < warp) ? c : 0u;
          all += c;
        }
        const unsigned int rank = at_taken + before + __popc(votes & ((1u << lane) - 1u));
        if (at && rank < need_at) cand[below + rank] = key;
        at_taken += all;
        buf ^= 1;                                 // the other copy is free: its readers passed the barrier above
      }
    }
    __syncthreads();

    // ---- order the kk keys by counting and write the row
    long long* irow = idx + qi * k;
    int* drow = dist + qi * k;
    for (unsigned int i = tid; i < kk; i += KNNK_THREADS) {
      const unsigned long long key = cand[i];
      unsigned int rank = 0;
      for (unsigned int x = 0; x < kk; ++x) rank += (cand[x] < key) ? 1u : 0u;
      irow[rank] = (long long)(key & 0xFFFFFFFFull);
      drow[rank] = (int)(key >> 32);
    }
    for (unsigned int i = kk + tid; i < (unsigned int)k; i += KNNK_THREADS) { irow[i] = -1; drow[i] = -1; }
    if (tid == 0) count[qi] = (int)kk;
    __syncthreads();                              // cand / hist are reused by the next query
  }
}

}  // namespace pbk

extern "C" {

int pbk_hamming_knnk(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt, int k, const uint8_t* mask,
                     uint64_t train_base, int64_t* idx, int32_t* dist, int32_t* count, pbk_stream_t stream) {
  using namespace pbk;
  PBK_REQUIRE(nq >= 0 && nt >= 0, PBK_EINVAL, "negative size");
  PBK_REQUIRE(k >= 1, PBK_EINVAL, "k must be >= 1");
  PBK_REQUIRE(k <= PBK_HAMMING_KNNK_MAX_K, PBK_EUNSUPPORTED, "k > %d is not supported", PBK_HAMMING_KNNK_MAX_K);
  PBK_REQUIRE(nt < (1ll << 32) && train_base + (uint64_t)nt <= (1ull << 32), PBK_EINVAL,
              "train rows must be addressable in 32 bits");
  if (nq == 0) return PBK_OK;
  PBK_REQUIRE(q != nullptr && idx != nullptr && dist != nullptr && count != nullptr && (nt == 0 || t != nullptr),
              PBK_EINVAL, "NULL device pointer");
  PBK_REQUIRE(aligned16(q) && aligned16(t), PBK_EALIGN, "descriptor pointers must be 16-byte aligned");
  const DeviceInfo* di = nullptr;
  int rc = current_device_info(&di);
  if (rc != PBK_OK) return rc;
  const long long cap = (long long)di->sm_count * 8;    // 8 CTAs of 256 threads per SM
  const unsigned grid = (unsigned)(nq < cap ? nq : cap);
  hamming_knnk_kernel<<<grid, KNNK_THREADS, 0, as_stream(stream)>>>(
      q, nq, t, nt, k, mask, train_base, reinterpret_cast<long long*>(idx), dist, count);
  PBK_CHECK_LAUNCH("hamming_knnk_kernel");
  return PBK_OK;
}

}  // extern "C"
