// host_bw.cpp -- host-side bandwidth of the real -> complex expansion (8 B read, 16 B written per sample) with T threads
// and non-temporal stores: decides whether shipping only the real parts over PCIe and expanding on the host can beat
// the direct 16 B / sample device->host copy.   g++ -O3 -mavx2 -pthread -o host_bw host_bw.cpp
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <thread>
#include <vector>

static void expand(const double *src, double *dst, size_t n) { // dst[2 i] = src[i], dst[2 i + 1] = 0
  const __m256d zero = _mm256_setzero_pd();
  size_t i = 0;
  for (; i + 4 <= n; i += 4) {
    const __m256d v = _mm256_loadu_pd(src + i);                  // a b c d
    const __m256d lo = _mm256_unpacklo_pd(v, zero);              // a 0 c 0
    const __m256d hi = _mm256_unpackhi_pd(v, zero);              // b 0 d 0
    _mm256_stream_pd(dst + 2 * i, _mm256_permute2f128_pd(lo, hi, 0x20));     // a 0 b 0
    _mm256_stream_pd(dst + 2 * i + 4, _mm256_permute2f128_pd(lo, hi, 0x31)); // c 0 d 0
  }
  for (; i < n; ++i) { dst[2 * i] = src[i]; dst[2 * i + 1] = 0; }
}

int main(int argc, char **argv) {
  const size_t n = (size_t)1 << 28; // 2 GiB of reals -> 4 GiB of complex
  double *src = (double *)aligned_alloc(64, n * 8), *dst = (double *)aligned_alloc(64, n * 16);
  memset(src, 1, n * 8);
  memset(dst, 0, n * 16); // fault the pages in
  for (int T : {1, 2, 4, 8, 12, 16, 32}) {
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t) {
        const size_t per = (n / T) & ~(size_t)3, b = per * t, e = t == T - 1 ? n : per * (t + 1);
        th.emplace_back(expand, src + b, dst + 2 * b, e - b);
      }
      for (auto &x : th) x.join();
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (n * 16 / s / 1e9 > best) best = n * 16 / s / 1e9;
    }
    printf("threads %2d: %.1f GB/s of complex written (%.1f GB/s of memory traffic)\n", T, best, best * 1.5);
  }
  // plain memcpy for comparison
  auto t0 = std::chrono::steady_clock::now();
  memcpy(dst, src, n * 8);
  printf("memcpy 1 thread: %.1f GB/s\n", n * 8 / std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / 1e9);
  return 0;
}
