/*
 * C restatement of cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, k=2) for
 * 256-bit descriptors.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * OpenCV (pinned opencv3=3.1.0 in /root/reference/environment.yml:53; probed
 * here: 4.13.0) implements it as batchDistance(..., NORM_HAMMING, K=2): for each
 * query it scans the train rows in ascending index order and inserts a
 * candidate only when its distance is STRICTLY smaller than a kept one, so
 * ties keep the lowest train index.  OpenCV's source
 * is not under /root/reference; parity is pinned by tests/test_oracle_pinning.py
 * against the installed cv2 on tie-heavy inputs.
 *
 * idx: (nq,2) int64 global train rows (-1 = none); dist: (nq,2) int32.
 */
#include <stdint.h>
#include <string.h>

void pbo_hamming_knn2(const uint8_t* q, int64_t nq, const uint8_t* t, int64_t nt,
                      int64_t* idx, int32_t* dist) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nq; ++i) {
    uint64_t a[4];
    memcpy(a, q + i * 32, 32);
    int32_t d1 = 1 << 30, d2 = 1 << 30;
    int64_t i1 = -1, i2 = -1;
    for (int64_t j = 0; j < nt; ++j) {
      uint64_t b[4];
      memcpy(b, t + j * 32, 32);
      const int32_t d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
                        __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
      if (d < d2) {
        if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
        else { d2 = d; i2 = j; }
      }
    }
    idx[2 * i] = i1; idx[2 * i + 1] = i2;
    dist[2 * i] = i1 < 0 ? -1 : d1;
    dist[2 * i + 1] = i2 < 0 ? -1 : d2;
  }
}
