// epgx_launch.h -- internal: kernel launchers, one translation unit per kernel family and precision so that the
// library builds in parallel.  Each returns cudaErrorInvalidValue when no instance matches the keys.
#pragma once
#include <cuda_runtime.h>

#include "epgx_common.cuh"

namespace epgx {
template <typename real> cudaError_t launch_ring(int npool, int nvt, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st);
template <typename real> cudaError_t launch_reg(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st);
template <typename real> cudaError_t launch_real(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st);
template <typename real> cudaError_t launch_realjac(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st);
template <typename real> cudaError_t launch_setjac(int slots, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st);
template <typename real> cudaError_t launch_pulsejac(int orders, const KParams &kp, dim3 grid, int threads, int smem, cudaStream_t st);
constexpr int kTapeChunk = 64;      // == TAPE_CHUNK of epgx_reg.cuh (checked there)
#ifndef EPGX_REAL_WINDOWS
#define EPGX_REAL_WINDOWS 2 // real kernel: tape windows joined into one kernel window (one coefficient-staging pass for 64 TRs)
#endif
constexpr int kRealWindows = EPGX_REAL_WINDOWS;
constexpr int kTrcPerWindow = 21;   // == TRC_PER_WINDOW
constexpr int kTrcReals = 14;       // == TRC_REALS
constexpr int kTrjPerWindow = 12;   // == TRJ_PER_WINDOW of epgx_realjac.cuh
constexpr int kTrjReals = 32;       // == TRJ_REALS: (1 + 3) state sets x 8
constexpr int kSjRow = 16;          // == SJ_ROW of epgx_setjac.cuh
} // namespace epgx
