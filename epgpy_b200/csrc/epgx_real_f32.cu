// epgx_real_f32.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_real.cuh"

namespace epgx {
#define LAUNCH(NS, MAXT)                                                                         \
  if (kp.bounded) real_kernel<float, NS, MAXT, true><<<grid, threads, smem, st>>>(kp);           \
  else real_kernel<float, NS, MAXT, false><<<grid, threads, smem, st>>>(kp);
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
