// FFMA issue rate by the kind of the coefficient operand: immediate, uniform register (kernel parameter), vector register.
// 8-sample x 8-tap x 3 FFMA blocks like the FIR stage of ofdm_link_fast_kernel; 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
struct Taps { float h[24]; };

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, const __grid_constant__ Taps tp, const float* __restrict__ gmem) {
  float x[16], acc[24];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = 0.f;
  float hr[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) hr[i] = MODE == 2 ? gmem[threadIdx.x + 512 * i] : 0.f;   // per-thread values: vector registers
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        const float xv = x[(i - l) & 15];
        float h0, h1, h2;
        if (MODE == 0) { h0 = 0.11f + 0.01f * l; h1 = 0.07f - 0.01f * l; h2 = 0.05f + 0.02f * l; }
        else if (MODE == 1) { h0 = tp.h[3 * l]; h1 = tp.h[3 * l + 1]; h2 = tp.h[3 * l + 2]; }
        else { h0 = hr[3 * l]; h1 = hr[3 * l + 1]; h2 = hr[3 * l + 2]; }
        acc[3 * i] = fmaf(h0, xv, acc[3 * i]);
        acc[3 * i + 1] = fmaf(h1, xv, acc[3 * i + 1]);
        acc[3 * i + 2] = fmaf(h2, xv, acc[3 * i + 2]);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = acc[i] * 0.125f + 0.01f;   // 16 FFMA: feedback keeps everything live
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float *d, *g; cudaMalloc(&d, 4); cudaMalloc(&g, 512 * 24 * 4); cudaMemset(g, 0, 512 * 24 * 4);
  Taps tp; for (int i = 0; i < 24; ++i) tp.h[i] = 0.05f + 0.003f * i;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  const char* names[3] = {"immediate", "uniform register (kernel parameter)", "vector register"};
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<p.multiProcessorCount, 512>>>(d, iters, tp, g);
      if (mode == 1) k<1><<<p.multiProcessorCount, 512>>>(d, iters, tp, g);
      if (mode == 2) k<2><<<p.multiProcessorCount, 512>>>(d, iters, tp, g);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double instr = 208.0;   // 192 + 16 FFMA per iteration
    printf("%-38s %.3f FFMA per clock per scheduler\n", names[mode], instr * iters * 4 / (best * 1e-3 * 1.92e9));
  }
  return 0;
}
