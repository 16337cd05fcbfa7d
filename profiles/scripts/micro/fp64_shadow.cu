// fp64_shadow.cu -- can other pipes issue in the shadow of the half-rate FP64 pipe on B200 (sm_100a)?
// Per warp: NF independent DFMA chains, interleaved 1:1 (or 2:1) with instructions of another kind.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_shadow fp64_shadow.cu ; run: ./fp64_shadow
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(128) k(double *out, int iters, double a, double b, int lane_src) {
  double x[8];
  float f[8];
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-3f + i; q[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE != 9) x[i] = fma(x[i], a, b);
        if (MODE == 1) f[i] = fmaf(f[i], 1.0001f, 0.5f);                   // + FFMA
        if (MODE == 2) q[i] = q[i] * 3 + 7;                                 // + IMAD
        if (MODE == 3) f[i] = __shfl_sync(0xffffffffu, f[i], lane_src);     // + SHFL (32-bit)
        if (MODE == 4) f[i] = (q[0] & (1 << i)) ? f[i] : f[(i + 1) & 7];    // + FSEL-ish
        if (MODE == 5) { f[i] = fmaf(f[i], 1.0001f, 0.5f); q[i] = q[i] * 3 + 7; } // + FFMA + IMAD (2 per DFMA)
        if (MODE == 9) { f[i] = fmaf(f[i], 1.0001f, 0.5f); q[i] = q[i] * 3 + 7; } // no DFMA
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + f[i] + q[i];
  if (s == 123.456) out[0] = s;
}

template <int MODE> float run(int blocks, int threads, int iters) {
  double *buf; cudaMalloc(&buf, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(buf, iters, 1.0000001, 1e-9, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  cudaFree(buf);
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int iters = 20000;
  printf("%s, %d SMs; cycles per DFMA warp-instruction per SMSP (32 DFMA per iteration per warp)\n", p.name, p.multiProcessorCount);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int wps = 1; wps <= 4; ++wps) { // warps per SMSP: one CTA of 128 threads per SM = 1 warp per SMSP
    const int blocks = p.multiProcessorCount * wps, threads = 128;
    float t0 = run<0>(blocks, threads, iters), t1 = run<1>(blocks, threads, iters), t2 = run<2>(blocks, threads, iters),
          t3 = run<3>(blocks, threads, iters), t4 = run<4>(blocks, threads, iters), t5 = run<5>(blocks, threads, iters),
          t9 = run<9>(blocks, threads, iters);
    const double per = 1.0 / (iters * 32.0 * wps) * 1e-3 * clk * 1e3; // ms -> cycles per DFMA per SMSP (at the nominal clock)
    printf("warps/SMSP %d: DFMA only %.2f | +FFMA %.2f | +IMAD %.2f | +SHFL %.2f | +FSEL %.2f | +FFMA+IMAD %.2f | FFMA+IMAD alone %.2f\n", wps,
           t0 * per, t1 * per, t2 * per, t3 * per, t4 * per, t5 * per, t9 * per);
  }
  return 0;
}
