// epgx_realjac_f64.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_realjac.cuh"

namespace epgx {
template <> cudaError_t launch_realjac<double>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  const bool mw = kp.G > 32; // several warps per atom
  switch (slots) {
  case 1: mw ? realjac_kernel<double, 1, 3, true><<<grid, threads, smem, st>>>(kp) : realjac_kernel<double, 1, 3, false><<<grid, threads, smem, st>>>(kp); break;
  case 2: mw ? realjac_kernel<double, 2, 3, true><<<grid, threads, smem, st>>>(kp) : realjac_kernel<double, 2, 3, false><<<grid, threads, smem, st>>>(kp); break;
  case 4:
    if (threads <= 128) mw ? realjac_kernel<double, 4, 3, true, 128><<<grid, threads, smem, st>>>(kp) : realjac_kernel<double, 4, 3, false, 128><<<grid, threads, smem, st>>>(kp);
    else mw ? realjac_kernel<double, 4, 3, true><<<grid, threads, smem, st>>>(kp) : realjac_kernel<double, 4, 3, false><<<grid, threads, smem, st>>>(kp);
    break;
  case 8: mw ? realjac_kernel<double, 8, 3, true><<<grid, threads, smem, st>>>(kp) : realjac_kernel<double, 8, 3, false><<<grid, threads, smem, st>>>(kp); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
} // namespace epgx
