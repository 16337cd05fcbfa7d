// epgx_real_f64.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_real.cuh"

namespace epgx {
template <> cudaError_t launch_real<double>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  switch (slots) {
  case 2: real_kernel<double, 2><<<grid, threads, smem, st>>>(kp); break;
  case 4: real_kernel<double, 4><<<grid, threads, smem, st>>>(kp); break;
  case 8:
    if (threads <= 128) real_kernel<double, 8, 128><<<grid, threads, smem, st>>>(kp);
    else real_kernel<double, 8><<<grid, threads, smem, st>>>(kp);
    break;
  case 16:
    if (threads <= 128) real_kernel<double, 16, 128><<<grid, threads, smem, st>>>(kp);
    else real_kernel<double, 16><<<grid, threads, smem, st>>>(kp);
    break;
  case 32: real_kernel<double, 32><<<grid, threads, smem, st>>>(kp); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
} // namespace epgx
