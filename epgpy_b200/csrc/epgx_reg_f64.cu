// epgx_reg_f64.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_reg.cuh"

namespace epgx {
static_assert(TAPE_CHUNK == kTapeChunk && TRC_PER_WINDOW == kTrcPerWindow && TRC_REALS == kTrcReals, "epgx_launch.h constants");
template <> cudaError_t launch_reg<double>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  switch (slots) {
  case 1: reg_kernel<double, 1><<<grid, threads, smem, st>>>(kp); break;
  case 2: reg_kernel<double, 2><<<grid, threads, smem, st>>>(kp); break;
  case 4:
    if (threads <= 128) reg_kernel<double, 4, 128><<<grid, threads, smem, st>>>(kp);
    else reg_kernel<double, 4><<<grid, threads, smem, st>>>(kp);
    break;
  case 8:
    if (threads <= 128) reg_kernel<double, 8, 128><<<grid, threads, smem, st>>>(kp);
    else reg_kernel<double, 8><<<grid, threads, smem, st>>>(kp);
    break;
  case 16:
    if (threads <= 128) reg_kernel<double, 16, 128><<<grid, threads, smem, st>>>(kp);
    else reg_kernel<double, 16><<<grid, threads, smem, st>>>(kp);
    break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
} // namespace epgx
