// A/B timing harness for the headline instantiation of ofdm_link_fast_kernel (N = 1024, 64-QAM, MMSE, 8 taps, CP = 7):
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -I ofdm-based-systems_b200/csrc -I include \
//        tools/microbench/fast_variants.cu -o tools/microbench/fast_variants
//   ./fast_variants [symbols] [reps]
// One process, every variant timed with CUDA events after warm-up; prints ms per launch and algorithmic TFLOP/s
// (197 084 flop per OFDM symbol).  The tables are synthetic but sane (unit-energy decaying taps, MMSE table of their
// spectrum); BER is printed so that a broken variant is visible.  Not part of the product.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "link_fast.cuh"
#ifdef WITH_STREAM_EXPERIMENT
#include "experiments/link_stream2_kernel.cuh"
#endif

using namespace ofdm;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static double g_fsym = 197084.0;
static std::vector<float2> twiddles(int E, bool fused) {
  const int RS = E + 2;
  std::vector<float2> tw(size_t(E) * RS, make_float2(0.f, 0.f));
  for (int k = 0; k < E; ++k)
    for (int r = fused ? 0 : 1; r < E; ++r) {
      const double ang = -2.0 * M_PI * double(k * r) / double(E * E);
      // kOptFusedTwiddle: pairs {W^(k n), W^(k (n + E/2))}
      const size_t at = fused ? size_t(k) * RS + 2 * (r % (E / 2)) + r / (E / 2) : size_t(k) * RS + r - 1;
      tw[at] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  return tw;
}

template <typename K>
static void run(const char* name, K kern, int block, size_t smem, FastParams p, unsigned long long* d_cnt, int reps, int points) {
  if (getenv("OCC1") && smem < 120 * 1024) smem = 120 * 1024;   // one block per SM
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  p.n_points = points;
  dim3 grid(148 * occ, points);
  for (int i = 0; i < 3; ++i) kern<<<grid, block, smem>>>(p);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(d_cnt, 0, 10 * 8 * kMaxSweepPoints));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) { p.seed = 1000 + i; kern<<<grid, block, smem>>>(p); }
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  unsigned long long h[10];
  CK(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
  const double flops = g_fsym * double(p.sym_count) * points;
  printf("%-44s occ %d  %8.4f ms  %6.2f TFLOP/s (%4.1f %% of 72.6)  BER %.5f\n", name, occ, ms, flops / ms / 1e9,
         100.0 * flops / ms / 1e9 / 72.6, h[1] ? double(h[0]) / double(h[1]) : 0.0);
}

int main(int argc, char** argv) {
  const unsigned long long nsym = argc > 1 ? strtoull(argv[1], nullptr, 10) : 162761ull;
  const int reps = argc > 2 ? atoi(argv[2]) : 20;
  const int points = argc > 3 ? atoi(argv[3]) : 1;
#ifndef HE
#define HE 32
#define HT 32
#endif
  constexpr int E = HE, T = HT, N = E * T, P = 7, L = 8;
  // unit-energy decaying taps, H = fft(taps, N)
  std::complex<double> h[L];
  double en = 0;
  for (int l = 0; l < L; ++l) { h[l] = std::polar(std::exp(-0.35 * l), 0.9 * l * l); en += std::norm(h[l]); }
  for (int l = 0; l < L; ++l) h[l] /= std::sqrt(en);
  const double knorm = std::sqrt(2.0 * 63.0 / 3.0), s = 8.0;   // 64-QAM
  FastParams p{};
  for (int l = 0; l < L; ++l) {
    const double re = h[l].real() / (knorm * std::sqrt(double(N))), im = h[l].imag() / (knorm * std::sqrt(double(N)));
    p.taps[l] = make_float2((float)re, (float)im);
    p.taps3[l] = make_float4((float)re, (float)(im - re), (float)(re + im), 0.f);
  }
  std::vector<float4> eq(N);
  for (int k = 0; k < N; ++k) {
    std::complex<double> H = 0;
    for (int l = 0; l < L; ++l) H += h[l] * std::polar(1.0, -2.0 * M_PI * k * l / N);
    // Y~ = sqrt(N) * H * X / knorm' ...: decision = sat(Re(Y~ conj A) / (G + sigma2) + 0.5) * (s - 1); level units are 2c-(s-1)
    const double c = 2.0 * (s - 1.0);          // full slicer span in level units
    const std::complex<double> A = H * c;      // so that Re(Y conj A) / |A|^2 * ... = level / c
    eq[k] = make_float4((float)A.real(), (float)A.imag(), (float)std::norm(A), (float)(s - 1.0));
  }
  auto tw = twiddles(E, false), twf = twiddles(E, true);
  { int lg = 0; while ((1 << lg) < N) ++lg; g_fsym = 10.0 * N * lg + 8.0 * L * (N + P) + 4.0 * (N + P) + N * 24.0; }
  static_assert(HE == HT, "harness: two-pass shapes only (no pass-3 twiddles)");
  float4* d_eq; float2* d_tw; float2* d_twf; unsigned long long* d_cnt;
  CK(cudaMalloc(&d_eq, N * sizeof(float4))); CK(cudaMalloc(&d_tw, tw.size() * sizeof(float2)));
  CK(cudaMalloc(&d_cnt, 10 * 8 * kMaxSweepPoints));
  CK(cudaMemcpy(d_eq, eq.data(), N * sizeof(float4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_twf, twf.size() * sizeof(float2)));
  CK(cudaMemcpy(d_twf, twf.data(), twf.size() * sizeof(float2), cudaMemcpyHostToDevice));
  p.eq_tab = d_eq; p.tw = d_tw;
  p.slice_top = float(s - 1.0); p.tx_scale2 = 1.f; p.z_unscale = 1.f; p.prefix_len = P; p.zero_prefix = 0; p.equalizer = 2;
  p.half_bits = 3; p.field_mask = 0x0E0E0E0Eu; p.seed = 1; p.point = 0; p.n_points = 1;
  const float sigma = (float)(std::sqrt(0.5 / 100.0) / knorm);    // 20 dB on the unit-power stream, in kernel units
  for (int i = 0; i < kMaxSweepPoints; ++i) p.point_tab[i] = SweepPoint{sigma, 1e-6f};
  p.sym_begin = 0; p.sym_count = nsym * (1024 / N); p.counters = d_cnt; p.y_scale = 1.f / 32.f;
  const char* only = getenv("ONLY");
#define VARIANT(name, BLOCK, SYNC, A, F, S, I, K, NR, FU, TAPS, OPT, ...)                                     \
  if (!only || strstr(name, only)) {                                                                    \
    p.tw = ((OPT) & kOptFusedTwiddle) ? d_twf : d_tw;                                                   \
    run(name, ofdm_link_fast_kernel<E, T, false, true, false, BLOCK, SYNC, A, F, S, I, K, NR, FU, TAPS, OPT, ##__VA_ARGS__>, BLOCK, \
        FastGeometry<E, T, BLOCK>::SMEM_BYTES, p, d_cnt, reps, points);                                 \
  }
#ifdef WITH_STREAM_EXPERIMENT
#define VARIANT2(name, BLOCK, TAPS, TRIG)                                                                \
  if ((p.tw = d_tw, !only) || strstr(name, only))                                                        \
    run(name, ofdm_link_stream_kernel<E, T, BLOCK, false, false, TAPS, TRIG>, BLOCK, StreamGeometry<E, T, BLOCK>::SMEM_BYTES, p, d_cnt, reps, points);
#else
#define VARIANT2(name, BLOCK, TAPS, TRIG)
#endif
#ifdef WITH_STREAM_EXPERIMENT
#define VARIANT3(name, MODE)                                                                             \
  if ((p.tw = d_tw, !only) || strstr(name, only))                                                        \
    run(name, ofdm_link_stream2_kernel<E, T, 512, MODE, 8>, 512, StreamGeometry<E, T, 512>::SMEM_BYTES + (((MODE) & 10) == 10 ? 32768 : 0), p, d_cnt, reps, points);
#else
#define VARIANT3(name, MODE)
#endif
#include "fast_variants.inc"
  return 0;
}
