// In-register radix-R FFT codelets with compile-time twiddles, plus the constexpr trigonometry
// they need.  These are the butterflies of the shared-memory Stockham transform in link_kernel.cuh,
// which replaces np.fft.ifft / np.fft.fft(norm="ortho") of the reference (modulation/models.py:32,46).
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace ofdm {

// ---------------------------------------------------------------- constexpr sin / cos (double)
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double cx_sin_small(double x) {  // |x| <= pi/4, Taylor to x^19
  double x2 = x * x, term = x, sum = x;
  for (int k = 1; k <= 10; ++k) {
    term *= -x2 / double((2 * k) * (2 * k + 1));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double cx_cos_small(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int k = 1; k <= 10; ++k) {
    term *= -x2 / double((2 * k - 1) * (2 * k));
    sum += term;
  }
  return sum;
}
// cos(2*pi*i/n), sin(2*pi*i/n) for 0 <= i < n, exact symmetries, argument reduced to [0, pi/4]
__host__ __device__ constexpr double cx_cos_frac(int i, int n) {
  i %= n;
  if (8 * i <= n) return cx_cos_small(2 * kPi * i / n);
  if (4 * i <= n) return cx_sin_small(2 * kPi * (n - 4 * i) / (4.0 * n));   // cos(x) = sin(pi/2 - x)
  if (2 * i <= n) return -cx_cos_frac(n - 2 * i, 2 * n) ;                   // cos(x) = -cos(pi - x)
  return cx_cos_frac(n - i, n);                                             // cos(x) = cos(2pi - x)
}
__host__ __device__ constexpr double cx_sin_frac(int i, int n) {
  i %= n;
  if (8 * i <= n) return cx_sin_small(2 * kPi * i / n);
  if (4 * i <= n) return cx_cos_small(2 * kPi * (n - 4 * i) / (4.0 * n));   // sin(x) = cos(pi/2 - x)
  if (2 * i <= n) return cx_sin_frac(n - 2 * i, 2 * n);                     // sin(x) = sin(pi - x)
  return -cx_sin_frac(n - i, n);                                            // sin(x) = -sin(2pi - x)
}

// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

template <int... I, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, I...>, F&& f) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

__host__ __device__ constexpr int cx_log2(int n) { return n <= 1 ? 0 : 1 + cx_log2(n / 2); }
__host__ __device__ constexpr int cx_brev(int x, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((x >> b) & 1) << (bits - 1 - b);
  return r;
}

// multiply d by W_n^i = exp(-/+ j 2 pi i / n)   (DIR = -1 forward, +1 inverse), i, n compile-time
template <int I, int NN, int DIR>
__device__ __forceinline__ float2 mul_const_twiddle(float2 d) {
  static_assert(I >= 0 && I < NN, "twiddle index");
  if constexpr (I == 0) {
    return d;
  } else if constexpr (4 * I == NN) {          // -j (forward) / +j (inverse)
    return DIR < 0 ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
  } else if constexpr (2 * I == NN) {
    return make_float2(-d.x, -d.y);
  } else if constexpr (8 * I == NN) {          // (1 -/+ j)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    return DIR < 0 ? make_float2((d.x + d.y) * h, (d.y - d.x) * h) : make_float2((d.x - d.y) * h, (d.y + d.x) * h);
  } else if constexpr (8 * I == 3 * NN) {      // (-1 -/+ j)/sqrt2
    constexpr float h = 0.70710678118654752440f;
    return DIR < 0 ? make_float2((d.y - d.x) * h, -(d.x + d.y) * h) : make_float2(-(d.x + d.y) * h, (d.x - d.y) * h);
  } else {
    constexpr float wr = (float)cx_cos_frac(I, NN);
    constexpr float wi = (float)(DIR < 0 ? -cx_sin_frac(I, NN) : cx_sin_frac(I, NN));
    return make_float2(fmaf(d.x, wr, -d.y * wi), fmaf(d.x, wi, d.y * wr));
  }
}

// In-place decimation-in-frequency FFT of R points held in registers.
// Output element k ends up in v[cx_brev(k, log2 R)]  (read it back with fft_out_index<R>(k)).
template <int R, int DIR>
__device__ __forceinline__ void fft_dif_inplace(float2 (&v)[R]) {
  static_assert((R & (R - 1)) == 0 && R >= 1, "radix must be a power of two");
  constexpr int LOG = cx_log2(R);
  static_for<LOG>([&](auto S) {
    constexpr int half = R >> (S.value + 1);
    constexpr int span = 2 * half;
    static_for<R / span>([&](auto B) {
      static_for<half>([&](auto I) {
        constexpr int lo = B.value * span + I.value, hi = lo + half;
        const float2 a = v[lo], c = v[hi];
        v[lo] = cadd(a, c);
        v[hi] = mul_const_twiddle<I.value, span, DIR>(csub(a, c));
      });
    });
  });
}
template <int R>
__host__ __device__ constexpr int fft_out_index(int k) { return cx_brev(k, cx_log2(R)); }

// Decimation-in-time butterfly with the compile-time twiddle W = W_NN^I applied to c, FMA-fused:
//   (a, c) <- (a + W c, a - W c).   General twiddles cost 6 instructions (4 FFMA for a + W c, then
//   a - W c = 2 a - (a + W c) as 2 FFMA); trivial ones 4 FADD; the 45-degree ones 2 FADD + 4 FFMA.
template <int I, int NN, int DIR>
__device__ __forceinline__ void dit_butterfly(float2& a, float2& c) {
  static_assert(I >= 0 && 2 * I < NN, "twiddle index");
  if constexpr (I == 0) {
    const float2 lo = cadd(a, c), hi = csub(a, c);
    a = lo;
    c = hi;
  } else if constexpr (4 * I == NN) {  // W = -j (forward) / +j (inverse)
    const float2 wc = DIR < 0 ? make_float2(c.y, -c.x) : make_float2(-c.y, c.x);
    const float2 lo = cadd(a, wc), hi = csub(a, wc);
    a = lo;
    c = hi;
  } else if constexpr (8 * I == NN || 8 * I == 3 * NN) {
    constexpr float h = 0.70710678118654752440f;
    float t1, t2;  // W c = h * (t1, t2)
    if constexpr (8 * I == NN) {
      t1 = DIR < 0 ? c.x + c.y : c.x - c.y;
      t2 = DIR < 0 ? c.y - c.x : c.x + c.y;
    } else {
      t1 = DIR < 0 ? c.y - c.x : -(c.x + c.y);
      t2 = DIR < 0 ? -(c.x + c.y) : c.x - c.y;
    }
    const float2 lo = make_float2(fmaf(h, t1, a.x), fmaf(h, t2, a.y));
    const float2 hi = make_float2(fmaf(-h, t1, a.x), fmaf(-h, t2, a.y));
    a = lo;
    c = hi;
  } else {
    constexpr float wr = (float)cx_cos_frac(I, NN);
    constexpr float wi = (float)(DIR < 0 ? -cx_sin_frac(I, NN) : cx_sin_frac(I, NN));
    const float2 lo = make_float2(fmaf(-wi, c.y, fmaf(wr, c.x, a.x)), fmaf(wi, c.x, fmaf(wr, c.y, a.y)));
    const float2 hi = make_float2(fmaf(2.0f, a.x, -lo.x), fmaf(2.0f, a.y, -lo.y));
    a = lo;
    c = hi;
  }
}

// In-place decimation-in-time FFT of R points in registers.  Same convention as fft_dif_inplace: input
// element n in v[n], output element k in v[fft_out_index<R>(k)] (working slot j lives in v[brev j]).
// FIRST = 1 skips the first stage (the trivial butterflies of elements n and n + R/2, n < R/2, in place): callers that
// can fuse work into those butterflies - a run-time twiddle per element, a common offset - do them themselves.
template <int R, int DIR, int FIRST = 0>
__device__ __forceinline__ void fft_dit_inplace(float2 (&v)[R]) {
  static_assert((R & (R - 1)) == 0 && R >= 1, "radix must be a power of two");
  constexpr int LOG = cx_log2(R);
  static_for<LOG - FIRST>([&](auto S0) {
    constexpr int S = decltype(S0)::value + FIRST;
    constexpr int half = 1 << S;
    constexpr int span = 2 * half;
    static_for<R / span>([&](auto B) {
      static_for<half>([&](auto I) {
        constexpr int lo = decltype(B)::value * span + decltype(I)::value, hi = lo + half;
        dit_butterfly<decltype(I)::value, span, DIR>(v[cx_brev(lo, LOG)], v[cx_brev(hi, LOG)]);
      });
    });
  });
}

}  // namespace ofdm
