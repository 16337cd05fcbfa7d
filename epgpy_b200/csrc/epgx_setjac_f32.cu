// epgx_setjac_f32.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_setjac.cuh"

namespace epgx {
static_assert(SJ_ROW == kSjRow, "staging row");
template <> cudaError_t launch_setjac<float>(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  switch (slots) {
  case 2: setjac_kernel<float, 2><<<grid, threads, smem, st>>>(kp); break;
  case 4: setjac_kernel<float, 4><<<grid, threads, smem, st>>>(kp); break;
  case 8: setjac_kernel<float, 8><<<grid, threads, smem, st>>>(kp); break;
  case 16: setjac_kernel<float, 16><<<grid, threads, smem, st>>>(kp); break;
  default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
} // namespace epgx
