// epgx_realjac.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_realjac.cuh"

namespace epgx {
template <> cudaError_t launch_realjac<double>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  switch (slots) {
  case 1: realjac_kernel<double, 1, 3><<<grid, threads, smem, st>>>(kp); break;
  case 2: realjac_kernel<double, 2, 3><<<grid, threads, smem, st>>>(kp); break;
  case 4: realjac_kernel<double, 4, 3><<<grid, threads, smem, st>>>(kp); break;
  case 8: realjac_kernel<double, 8, 3><<<grid, threads, smem, st>>>(kp); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <> cudaError_t launch_realjac<float>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  switch (slots) {
  case 1: realjac_kernel<float, 1, 3><<<grid, threads, smem, st>>>(kp); break;
  case 2: realjac_kernel<float, 2, 3><<<grid, threads, smem, st>>>(kp); break;
  case 4: realjac_kernel<float, 4, 3><<<grid, threads, smem, st>>>(kp); break;
  case 8: realjac_kernel<float, 8, 3><<<grid, threads, smem, st>>>(kp); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
} // namespace epgx
