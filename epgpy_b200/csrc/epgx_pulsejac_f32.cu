// epgx_pulsejac_f32.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_pulsejac.cuh"

namespace epgx {
template <> cudaError_t launch_pulsejac<float>(int orders, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  if (threads != PJ_THREADS) return cudaErrorInvalidValue;
  switch (orders) {
  case 4: pulsejac_kernel<float, 4><<<grid, threads, smem, st>>>(kp); break;
  case 8: pulsejac_kernel<float, 8><<<grid, threads, smem, st>>>(kp); break;
  case 12: pulsejac_kernel<float, 12><<<grid, threads, smem, st>>>(kp); break;
  case 16: pulsejac_kernel<float, 16><<<grid, threads, smem, st>>>(kp); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
} // namespace epgx
