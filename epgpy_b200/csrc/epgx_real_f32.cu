// epgx_real_f32.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_real.cuh"

namespace epgx {
// shared-memory carve-out: enough for the CTAs the register budget admits (the driver's default heuristic may prefer L1
// and leave one CTA out); the rest stays L1 for the coefficient table
template <typename K> static void carve(K kernel, int smem, int threads, int maxt, int blocks) {
  const int ctas = blocks * (threads > 0 && maxt / threads > 1 ? maxt / threads : 1);
  const int pct = (int)(((long long)(smem + 1024) * ctas * 100 + 228 * 1024 - 1) / (228 * 1024));
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); // (opt-in above 48 kB)
}
#define LAUNCH(NS, MAXT)                                                                                       \
  if (kp.bounded) {                                                                                            \
    carve(real_kernel<float, NS, MAXT, true>, smem, threads, MAXT, real_min_blocks(sizeof(float) * NS, MAXT));        \
    real_kernel<float, NS, MAXT, true><<<grid, threads, smem, st>>>(kp);                                          \
  } else {                                                                                                     \
    carve(real_kernel<float, NS, MAXT, false>, smem, threads, MAXT, real_min_blocks(sizeof(float) * NS, MAXT));       \
    real_kernel<float, NS, MAXT, false><<<grid, threads, smem, st>>>(kp);                                         \
  }
template <> cudaError_t launch_real<float>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  switch (slots) {
  case 2: LAUNCH(2, 256) break;
  case 4: LAUNCH(4, 256) break;
  case 8:
    if (threads <= 128) { LAUNCH(8, 128) } else { LAUNCH(8, 256) }
    break;
  case 16:
    if (threads <= 128) { LAUNCH(16, 128) } else { LAUNCH(16, 256) }
    break;
  case 32: LAUNCH(32, 256) break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
#undef LAUNCH
} // namespace epgx
