// Second set of issue-rate probes (sm_100a): integer / MUFU / conversion shapes used by the link kernel.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters, float a, float b, unsigned ka, const float* __restrict__ src) {
  float x[16], y[16];
  unsigned u[16], w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    x[i] = src[threadIdx.x + 32 * i] + 1.5f;
    y[i] = src[threadIdx.x + 32 * i + 512] + 0.25f;
    u[i] = __float_as_uint(x[i]) + i * 77u;
    w[i] = __float_as_uint(y[i]) + i * 131u;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) { unsigned long long p = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(p >> 32) ^ (unsigned)p; }  // IMAD.WIDE + LOP3
        if (MODE == 1) { u[i] = __umulhi(u[i], 0xD2511F53u); }                                // IMAD.HI
        if (MODE == 2) { u[i] = u[i] * 0xD2511F53u + 12345u; }                                // IMAD lo
        if (MODE == 3) { u[i] = u[i] ^ w[i] ^ ka; }                                           // LOP3 reg reg const
        if (MODE == 4) { u[i] = u[i] ^ w[i] ^ w[(i + 5) & 15]; }                              // LOP3 3 regs
        if (MODE == 5) { x[i] = __log2f(x[i]) ; }                                             // MUFU.LG2 (+ maybe range fix)
        if (MODE == 6) { float v; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(x[i])); x[i] = v; }
        if (MODE == 7) { x[i] = (float)u[i]; u[i] += 3u; }                                    // I2FP + IADD
        if (MODE == 8) { x[i] = fmaf(x[i], a, y[i]); }                                        // FFMA reg, const, reg
        if (MODE == 9) { x[i] = fmaf(x[i], 0.70710678f, y[i]); }                              // FFMA reg, imm, reg
        if (MODE == 10) { u[i] = __byte_perm(u[i], w[i], 0x7610); }                           // PRMT
        if (MODE == 11) { u[i] = __funnelshift_l(u[i], w[i], 7); }                            // SHF
        if (MODE == 12) { x[i] = __saturatef(fmaf(x[i], a, 0.5f)); }                          // FFMA.SAT
        if (MODE == 13) { x[i] = fmaf(x[i], y[i], x[(i + 1) & 15]); }                         // FFMA 3 regs (2 distinct + chain)
        if (MODE == 14) { x[i] = fmaf(y[i], y[i], x[i]); }                                    // FFMA y*y+x (2 distinct regs)
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i] + __uint_as_float(u[i]);
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm, float* d, const float* src, int sms, double ghz) {
  const int iters = 2048, threads = 32 * warps_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<sms, threads>>>(d, iters, 0.999f, 0.001f, 0x9E3779B9u, src);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double ops = double(iters) * 64 * warps_per_sm;
  const double cycles = best * 1e-3 * ghz * 1e9;
  printf("%-44s warps/SM=%2d  cycles per op per SMSP = %.3f\n", name, warps_per_sm, cycles * 4 / ops);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const double ghz = 1.92;
  float *d, *src; cudaMalloc(&d, 4); cudaMalloc(&src, 4096 * 4); cudaMemset(src, 0, 4096 * 4);
  const int w = 16, sms = p.multiProcessorCount;
  run<0>("IMAD.WIDE.U32 + xor(hi,lo)", w, d, src, sms, ghz);
  run<1>("IMAD.HI.U32", w, d, src, sms, ghz);
  run<2>("IMAD lo (reg*imm+imm)", w, d, src, sms, ghz);
  run<3>("LOP3 reg,reg,const", w, d, src, sms, ghz);
  run<4>("LOP3 3 regs", w, d, src, sms, ghz);
  run<5>("__log2f", w, d, src, sms, ghz);
  run<6>("lg2.approx.ftz", w, d, src, sms, ghz);
  run<7>("I2FP + IADD", w, d, src, sms, ghz);
  run<8>("FFMA reg,const,reg", w, d, src, sms, ghz);
  run<9>("FFMA reg,imm,reg", w, d, src, sms, ghz);
  run<10>("PRMT reg,reg", w, d, src, sms, ghz);
  run<11>("SHF.L.W reg,reg,imm", w, d, src, sms, ghz);
  run<12>("FFMA + SAT", w, d, src, sms, ghz);
  run<13>("FFMA x*y+x' (3 regs)", w, d, src, sms, ghz);
  run<14>("FFMA y*y+x (2 distinct)", w, d, src, sms, ghz);
  return 0;
}
