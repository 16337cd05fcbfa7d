// epgx_ring_f64.cu -- instantiations of one kernel family (see epgx_launch.h)
#include "epgx_launch.h"
#include "epgx_ring.cuh"

namespace epgx {
template <typename real, int NP, int NVT>
static cudaError_t go(const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
  auto kern = ring_kernel<real, NP, NVT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, threads, smem, st>>>(kp);
  return cudaGetLastError();
}

template <> cudaError_t launch_ring<double>(int np, int nvt, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st) {
#define CASE(NP_, NVT_) if (np == NP_ && nvt == NVT_) return go<double, NP_, NVT_>(kp, grid, threads, smem, st);
  CASE(1, 0) CASE(1, 1) CASE(1, 3) CASE(2, 0) CASE(2, 1) CASE(2, 3)
  CASE(3, 0) CASE(3, 1) CASE(4, 0) CASE(4, 1) // three / four exchange pools: at most one resident partial state (registers)
#undef CASE
  return cudaErrorInvalidValue;
}
} // namespace epgx
