// epgx_host.cu -- host-side data marshalling of the C ABI (no arithmetic of the EPG path): widening rows of real samples
// to the complex rows of the reference's API with several threads and streaming stores (include/epgx.h).
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

#include "epgx.h"

namespace {

template <typename real> void expand_rows(const real *src, int64_t src_pitch, real *dst, int64_t dst_pitch, int64_t r0, int64_t r1, int64_t cols) {
  for (int64_t r = r0; r < r1; ++r) {
    const real *s = src + r * src_pitch;
    real *d = dst + 2 * r * dst_pitch; // complex elements: two reals each
    int64_t c = 0;
#if defined(__AVX2__)
    if (sizeof(real) == 8) {
      const __m256d zero = _mm256_setzero_pd();
      while (c < cols && ((uintptr_t)(d + 2 * c) & 31)) { d[2 * c] = s[c]; d[2 * c + 1] = 0; ++c; }
      for (; c + 4 <= cols; c += 4) {
        const __m256d v = _mm256_loadu_pd((const double *)s + c);      // a b c d
        const __m256d lo = _mm256_unpacklo_pd(v, zero);                // a 0 c 0
        const __m256d hi = _mm256_unpackhi_pd(v, zero);                // b 0 d 0
        _mm256_stream_pd((double *)d + 2 * c, _mm256_permute2f128_pd(lo, hi, 0x20));     // a 0 b 0
        _mm256_stream_pd((double *)d + 2 * c + 4, _mm256_permute2f128_pd(lo, hi, 0x31)); // c 0 d 0
      }
    }
#endif
    for (; c < cols; ++c) { d[2 * c] = s[c]; d[2 * c + 1] = 0; }
  }
#if defined(__AVX2__)
  _mm_sfence();
#endif
}

template <typename real> void expand(const void *src, int64_t src_pitch, void *dst, int64_t dst_pitch, int64_t rows, int64_t cols, int nthreads) {
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, rows));
  if (nthreads == 1) return expand_rows<real>((const real *)src, src_pitch, (real *)dst, dst_pitch, 0, rows, cols);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; ++t) {
    const int64_t r0 = rows * t / nthreads, r1 = rows * (t + 1) / nthreads;
    pool.emplace_back(expand_rows<real>, (const real *)src, src_pitch, (real *)dst, dst_pitch, r0, r1, cols);
  }
  for (auto &th : pool) th.join();
}

} // namespace

extern "C" int epgx_expand_real(int dtype, const void *src, int64_t src_pitch, void *dst, int64_t dst_pitch, int64_t rows,
                                int64_t cols, int nthreads) {
  if (!src || !dst || rows < 0 || cols < 0 || src_pitch < cols || dst_pitch < cols) return EPGX_ERR_INVALID;
  if (dtype == EPGX_F64) expand<double>(src, src_pitch, dst, dst_pitch, rows, cols, nthreads);
  else if (dtype == EPGX_F32) expand<float>(src, src_pitch, dst, dst_pitch, rows, cols, nthreads);
  else return EPGX_ERR_INVALID;
  return EPGX_OK;
}
