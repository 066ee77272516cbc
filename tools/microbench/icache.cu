// Instruction-cache probe, sm_100a: a straight-line loop body of K FFMAs (16 B each) executed by 16 warps per SM,
// either in lock step or staggered by 1/16 of the body (free-running warps of a long kernel).  Prints cycles per
// warp-instruction per scheduler (1.0 = full issue rate) against the body size.
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void __launch_bounds__(512) body(float* out, int iters, int stagger, float a, float b) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
  if (stagger) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)(threadIdx.x >> 5) * stagger) {}
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i & 7] = fmaf(acc[i & 7], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

template <int K>
void run(float* d, int sms) {
  const int iters = (1 << 22) / K;
  for (int stag = 0; stag < 2; ++stag) {
    const int stagger = stag ? K / 4 : 0;   // 16 warps, 4 per scheduler: a body takes ~4K cycles per warp at full rate
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      body<K><<<sms, 512>>>(d, iters, stagger, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double inst_per_sched = double(iters) * K * 4;   // 4 warps per scheduler
    printf("body %6.1f KB  %-9s  %.3f cycles per warp-instruction per scheduler\n", K * 16 / 1024.0, stag ? "staggered" : "lock-step",
           best * 1e-3 * 1.92e9 / inst_per_sched);
  }
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* d; cudaMalloc(&d, 4);
  run<512>(d, p.multiProcessorCount);
  run<1024>(d, p.multiProcessorCount);
  run<1536>(d, p.multiProcessorCount);
  run<2048>(d, p.multiProcessorCount);
  run<2560>(d, p.multiProcessorCount);
  run<3072>(d, p.multiProcessorCount);
  run<3584>(d, p.multiProcessorCount);
  run<4096>(d, p.multiProcessorCount);
  run<5120>(d, p.multiProcessorCount);
  run<6144>(d, p.multiProcessorCount);
  run<8192>(d, p.multiProcessorCount);
  return 0;
}
