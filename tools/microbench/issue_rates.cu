// Issue-rate probes for sm_100a: warp-instructions per clock per SM sub-partition for the instruction
// shapes the link kernel is made of.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_rates issue_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters, float a, float b, const float* __restrict__ src) {
  float x[16], y[16], z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = src[threadIdx.x + 32 * i];
    y[i] = src[threadIdx.x + 32 * i + 512] ;
    z[i] = src[threadIdx.x + 32 * i + 1024];
  }
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(x[i]) + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) x[i] = fmaf(x[i], a, b);                       // reg, const, const
        if (MODE == 1) x[i] = fmaf(x[i], y[i], z[i]);                 // 3 distinct registers
        if (MODE == 2) x[i] = fmaf(x[i], y[i], b);                    // reg, reg, const
        if (MODE == 3) x[i] = x[i] + y[i];                            // FADD reg reg
        if (MODE == 4) x[i] = fmaf(y[i], z[i], x[i]);                 // accumulate form (a*b + acc)
        if (MODE == 5) x[i] = fmaf(y[(i + 1) & 15], z[(i + 3) & 15], x[i]);
        if (MODE == 6) { x[i] = fmaf(y[i], z[i], x[i]); if ((i & 3) == 0) u[i >> 2] = (u[i >> 2] ^ u[(i >> 2) + 4]) + 0x9E3779B9u; }   // 4 FFMA : 1.? ALU
        if (MODE == 7) { x[i] = fmaf(y[i], z[i], x[i]); if ((i & 1) == 0) u[i >> 1] = (u[i >> 1] * 0xD2511F53u) ^ 0x12345u; }  // IMAD mix
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float(u[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm, float* d, const float* src, int sms, double ghz) {
  const int iters = 4096, threads = 32 * warps_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<sms, threads>>>(d, iters, 0.999f, 0.001f, src);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double fp_insts = double(iters) * 64 * warps_per_sm;     // FP warp-instructions per SM
  const double cycles = best * 1e-3 * ghz * 1e9;
  printf("%-44s warps/SM=%2d  FP warp-inst/clk/SMSP = %.3f\n", name, warps_per_sm, fp_insts / cycles / 4);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  printf("%s  SMs=%d  clock=%.3f GHz (nominal max; rates assume it)\n", p.name, p.multiProcessorCount, ghz);
  float *d, *src; cudaMalloc(&d, 4); cudaMalloc(&src, 4096 * 4); cudaMemset(src, 0, 4096 * 4);
  for (int w : {4, 8, 16, 32}) {
    run<0>("FFMA reg,const,const", w, d, src, p.multiProcessorCount, ghz);
    run<1>("FFMA x=x*y+z (3 regs)", w, d, src, p.multiProcessorCount, ghz);
    run<2>("FFMA x=x*y+c", w, d, src, p.multiProcessorCount, ghz);
    run<3>("FADD x=x+y", w, d, src, p.multiProcessorCount, ghz);
    run<4>("FFMA x=y*z+x", w, d, src, p.multiProcessorCount, ghz);
    run<5>("FFMA x=y[i+1]*z[i+3]+x", w, d, src, p.multiProcessorCount, ghz);
    run<6>("FFMA + 1/4 (LOP3,IADD)", w, d, src, p.multiProcessorCount, ghz);
    run<7>("FFMA + 1/2 (IMAD,LOP3)", w, d, src, p.multiProcessorCount, ghz);
  }
  return 0;
}
