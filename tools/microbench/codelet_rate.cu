// Issue rate of the radix-32 register codelet alone (and of the 8-tap Gauss FIR chunk alone): 16 warps per SM, no memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_regs.cuh"
using namespace ofdm;

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float a) {
  float2 v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, a * i);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      fft_dit_inplace<32, -1>(v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i].x *= 0.03125f;   // keep the values bounded (32 FMUL per 388-instruction codelet)
    } else {
      // FIR chunk: 8 samples x 8 taps x 3 FFMA with the taps in registers-as-constants (a, a+1, ...)
      float2 y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float k1 = 0.f, k2 = 0.f, k3 = 0.f;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          const float2 x = v[(i - l) & 31];
          k1 = fmaf(0.11f + 0.01f * l, x.x + x.y, k1);
          k2 = fmaf(0.07f - 0.01f * l, x.x, k2);
          k3 = fmaf(0.05f + 0.02f * l, x.y, k3);
        }
        y[i] = make_float2(k1 - k3, k1 + k2);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i + 8 * (it & 3)] = y[i];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i].x + v[i].y;
  if (s == 123.456f) out[0] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 1; ++mode) {
    for (int w : {4, 8, 16}) {
      float best = 1e30f;
      const int iters = 20000;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<0><<<p.multiProcessorCount, 32 * w>>>(d, iters, 0.5f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
      }
      // cycles per codelet per scheduler-warp: w/4 warps per scheduler
      printf("codelet: %2d warps/SM: %.1f cycles per codelet iteration per scheduler (420 instructions incl. 32 FMUL) -> %.2f instr/clk\n", w,
             best * 1e-3 * 1.92e9 / iters / (w / 4), 420.0 / (best * 1e-3 * 1.92e9 / iters / (w / 4)));
    }
  }
  return 0;
}
