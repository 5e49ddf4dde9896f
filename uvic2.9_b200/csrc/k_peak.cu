// FP64 issue-rate micro-benchmark: the roof the MOBI and flux kernels are measured against.
// Dependent chains of DFMA / DADD / DMUL, 8 independent chains per thread (enough ILP to cover the FP64 pipe latency),
// 8 warps per scheduler partition resident, one CTA wave over all SMs; timed with CUDA events.  No memory traffic.
#include <cuda_runtime.h>
#include <cstdio>
#include "../../include/uvic_b200.h"

template <int OP>
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int q = 0; q < 8; q++) x[q] = (double)(threadIdx.x + q) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int q = 0; q < 8; q++) {
        if (OP == 0) x[q] = __fma_rn(x[q], a, b);
        if (OP == 1) x[q] = __dadd_rn(x[q], b);
        if (OP == 2) x[q] = __dmul_rn(x[q], a);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) s += x[q];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

extern "C" int uvic_b200_measure_fp64_peak(int device, double *dfma_per_s, double *dadd_per_s, double *dmul_per_s, double *sm_clock_mhz) {
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 1;
  const int ctas = p.multiProcessorCount * 4, threads = 256, iters = 4096;
  double *d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * (size_t)ctas * threads) != cudaSuccess) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double *outs[3] = {dfma_per_s, dadd_per_s, dmul_per_s};
  for (int op = 0; op < 3; op++) {
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (op == 0) k_fp64_peak<0><<<ctas, threads>>>(d, iters, 0.9999999, 1e-7);
      if (op == 1) k_fp64_peak<1><<<ctas, threads>>>(d, iters, 0.9999999, 1e-7);
      if (op == 2) k_fp64_peak<2><<<ctas, threads>>>(d, iters, 0.9999999, 1e-7);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return 1; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      double n = (double)ctas * threads * iters * 16.0 * 8.0;   // thread-level instructions
      double rate = n / (ms * 1e-3);
      if (rep > 0 && rate > best) best = rate;
    }
    if (outs[op]) *outs[op] = best;
  }
  if (sm_clock_mhz) *sm_clock_mhz = p.clockRate * 1e-3;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}
