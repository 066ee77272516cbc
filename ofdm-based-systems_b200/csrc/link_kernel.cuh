// ofdm_link kernel: the whole per-OFDM-symbol chain of the reference's Simulation.run()
// (simulation/models.py:454-606) in ONE launch.  A "team" of T = N/E threads owns one OFDM symbol at
// a time and keeps its E = N/T samples per thread in registers; shared memory is only used for the
// Stockham exchanges between FFT passes and for the FIR's neighbour reads.
//
//   bits (Philox | replayed byte stream)        bits_generation/models.py:27-55, simulation/models.py:59-69
//   -> QAM / PSK map, per-subcarrier order      constellation/models.py:220-249,382-411; adaptive.py:130-201
//   -> ortho IFFT (OFDM) | identity (SC-OFDM)   modulation/models.py:27-39, 66-72
//   -> CP / ZP / no prefix, PAPR statistics     prefix/models.py:34-69; simulation/models.py:519-524
//   -> causal FIR over the serial stream        channel/models.py:46-55   (inter-symbol tail carried)
//   -> + AWGN (Philox Box-Muller | replayed)    noise/models.py:13-22
//   -> strip / overlap-add, ortho FFT           prefix/models.py:46-52,71-101; modulation/models.py:41-48
//   -> ZF / MMSE (per-symbol sigma^2) / none    equalization/models.py:22-68
//   -> (SC-OFDM: ortho IFFT)                    modulation/models.py:89
//   -> hard demap, bit + symbol error counts    constellation/models.py:19-27,251-295; simulation/models.py:597-606
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fft_regs.cuh"
#include "link_params.h"
#include "philox.cuh"

namespace ofdm {

template <int N, int E>
struct Geometry {
  static_assert(N % E == 0, "E must divide N");
  static constexpr int T = N / E;                     // threads per OFDM symbol
  static constexpr int BLOCK = T > 128 ? T : 128;
  static constexpr int TEAMS = BLOCK / T;             // OFDM symbols in flight per block
  static constexpr int R1 = E;
  static constexpr int REM1 = N / R1;
  static constexpr int R2 = REM1 < E ? REM1 : E;
  static constexpr int R3 = REM1 / R2;
  static_assert(R1 * R2 * R3 == N && R3 <= E, "unsupported FFT plan");
  static constexpr int BUF = N + N / 32 + 1;          // padded exchange buffer (float2)
  static constexpr int TEAM_SMEM = BUF + kMaxTaps;    // + previous-symbol tail
  static constexpr int RED = T > 32 ? T / 32 : 1;     // cross-warp reduction scratch (floats)
  static constexpr size_t SMEM_BYTES = size_t(TEAMS) * (TEAM_SMEM * sizeof(float2) + RED * sizeof(float));
  static constexpr int TW2 = (R2 > 1) ? (R2 - 1) * (N / R2) : 0;  // twiddle table sizes
  static constexpr int TW3 = (R3 > 1) ? (R3 - 1) * (N / R3) : 0;
};

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 5); }

template <int T>
struct Team {
  int t;               // thread index inside the team
  int team_in_block;
  unsigned mask;       // lanes of this team inside its warp (T <= 32)
  float* red;
  __device__ __forceinline__ void sync() const {
    if constexpr (T <= 32) {
      __syncwarp(mask);
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(team_in_block + 1), "n"(T) : "memory");
    }
  }
  __device__ __forceinline__ float sum(float x) const {
    constexpr int W = T < 32 ? T : 32;
#pragma unroll
    for (int off = W / 2; off >= 1; off >>= 1) x += __shfl_xor_sync(mask, x, off);
    if constexpr (T > 32) {
      const int w = t >> 5;
      if ((t & 31) == 0) red[w] = x;
      sync();
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < T / 32; ++i) s += red[i];
      sync();
      x = s;
    }
    return x;
  }
};

// One Stockham pass of radix R over data held as v[m] <-> index t + T*m.
//   leg r of butterfly j = t + T*q is v[q + r*(E/R)];  k = j mod NS
//   u[r] = v[..] * W_{NS*R}^{k r};  U = FFT_R(u);  out[(j-k)*R + k + r*NS] = U[r]
// The last pass (NS*R == N) has k == j, so its output lands back in v[q + r*(E/R)].
template <int N, int E, int R, int NS, int DIR, bool LAST>
__device__ __forceinline__ void stockham_pass(float2 (&v)[E], float2* buf, const float2* __restrict__ tw, int t) {
  constexpr int T = N / E, Q = E / R, NB = N / R;
  static_for<Q>([&](auto Qi) {
    constexpr int q = decltype(Qi)::value;
    const int j = t + T * q;
    float2 u[R];
#pragma unroll
    for (int r = 0; r < R; ++r) u[r] = v[q + r * Q];
    if constexpr (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) {
        const float2 w = __ldg(&tw[(r - 1) * NB + j]);
        u[r] = DIR < 0 ? cmul(u[r], w) : cmul_conj(u[r], w);
      }
    }
    fft_dif_inplace<R, DIR>(u);
    if constexpr (LAST) {
      static_for<R>([&](auto Ri) {
        constexpr int r = decltype(Ri)::value;
        v[q + r * Q] = u[fft_out_index<R>(r)];
      });
    } else {
      const int k = j & (NS - 1);
      const int base = (j - k) * R + k;
      static_for<R>([&](auto Ri) {
        constexpr int r = decltype(Ri)::value;
        buf[pad_idx(base + r * NS)] = u[fft_out_index<R>(r)];
      });
    }
  });
}

template <int N, int E, int T>
__device__ __forceinline__ void reload_strided(float2 (&v)[E], const float2* buf, const Team<T>& team) {
  team.sync();
#pragma unroll
  for (int m = 0; m < E; ++m) v[m] = buf[pad_idx(team.t + T * m)];
  team.sync();
}

// Unscaled N-point DFT (DIR = -1) or inverse DFT (DIR = +1) of v (natural order, strided layout).
template <int N, int E, int DIR>
__device__ __forceinline__ void team_fft(float2 (&v)[E], float2* buf, const float2* __restrict__ tw,
                                         const Team<N / E>& team) {
  using G = Geometry<N, E>;
  constexpr int T = G::T;
  if constexpr (G::R2 == 1) {
    stockham_pass<N, E, G::R1, 1, DIR, true>(v, buf, tw, team.t);
  } else if constexpr (G::R3 == 1) {
    stockham_pass<N, E, G::R1, 1, DIR, false>(v, buf, tw, team.t);
    reload_strided<N, E, T>(v, buf, team);
    stockham_pass<N, E, G::R2, G::R1, DIR, true>(v, buf, tw, team.t);
  } else {
    stockham_pass<N, E, G::R1, 1, DIR, false>(v, buf, tw, team.t);
    reload_strided<N, E, T>(v, buf, team);
    stockham_pass<N, E, G::R2, G::R1, DIR, false>(v, buf, tw, team.t);
    reload_strided<N, E, T>(v, buf, team);
    stockham_pass<N, E, G::R3, G::R1 * G::R2, DIR, true>(v, buf, tw + G::TW2, team.t);
  }
}

__device__ __forceinline__ unsigned inv_gray(unsigned x) {
  x ^= x >> 1;
  x ^= x >> 2;
  x ^= x >> 4;
  x ^= x >> 8;
  return x;
}

// label -> constellation point (closed form of constellation/models.py:180-218 and :356-380)
__device__ __forceinline__ float2 map_label(unsigned lab, int bps, float amp, int scheme) {
  if (bps == 0) return make_float2(0.f, 0.f);
  if (scheme == SCHEME_QAM) {
    const int m2 = bps >> 1, s = 1 << m2;
    const unsigned lo = lab & (s - 1), hi = lab >> m2;
    const int gi = int(lo ^ (lo >> 1)), gq = int(hi ^ (hi >> 1));
    return make_float2(float(2 * gi - (s - 1)) * amp, float((s - 1) - 2 * gq) * amp);
  }
  const unsigned k = inv_gray(lab);
  float sn, cs;
  sincospif(2.0f * float(k) / float(1 << bps), &sn, &cs);
  return make_float2(cs * amp, sn * amp);
}

// nearest constellation point -> label (slicer form of the O(n*M) search of constellation/models.py:19-27)
__device__ __forceinline__ unsigned demap_point(float2 z, int bps, float kslice, int scheme) {
  if (bps == 0) return 0u;
  if (scheme == SCHEME_QAM) {
    const int m2 = bps >> 1, s = 1 << m2;
    const float top = float(s - 1), half = 0.5f * top;
    const float fc = fminf(fmaxf(fmaf(z.x, 0.5f * kslice, half), 0.f), top);
    const float fr = fminf(fmaxf(fmaf(-z.y, 0.5f * kslice, half), 0.f), top);
    const unsigned col = (unsigned)__float2int_rn(fc), row = (unsigned)__float2int_rn(fr);
    return (inv_gray(row) << m2) | inv_gray(col);
  }
  const int M = 1 << bps;
  const float a = atan2f(z.y, z.x) * (float(M) * 0.15915494309189535f);
  const unsigned k = unsigned(__float2int_rn(a)) & unsigned(M - 1);
  return k ^ (k >> 1);
}

template <int N, int E>
__global__ void __launch_bounds__(Geometry<N, E>::BLOCK) ofdm_link_kernel(const LinkParams p) {
  using G = Geometry<N, E>;
  constexpr int T = G::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Team<T> team;
  team.t = threadIdx.x % T;
  team.team_in_block = threadIdx.x / T;
  team.mask = (T >= 32) ? 0xffffffffu : (((1u << (T & 31)) - 1u) << ((threadIdx.x & 31) & ~(T - 1)));
  float2* buf = reinterpret_cast<float2*>(smem_raw) + size_t(team.team_in_block) * G::TEAM_SMEM;
  float2* tail = buf + G::BUF;
  team.red = reinterpret_cast<float*>(reinterpret_cast<float2*>(smem_raw) + size_t(G::TEAMS) * G::TEAM_SMEM) +
             team.team_in_block * G::RED;
  const int t = team.t;

  const int P = p.prefix_len, L = p.n_taps, NP = N + P;
  const int TL = L - 1;  // tail length carried between symbols
  const PhiloxKey key{(uint32_t)p.seed, (uint32_t)(p.seed >> 32)};
  const float inv_sqrt_n = rsqrtf(float(N));
  const bool cp_fast = (p.prefix_type == PREFIX_CYCLIC) && !p.isi;

  // contiguous chunk of OFDM symbols for this team (needed for the ISI chain)
  const unsigned long long n_teams = (unsigned long long)gridDim.x * G::TEAMS;
  const unsigned long long team_id = (unsigned long long)blockIdx.x * G::TEAMS + team.team_in_block;
  const unsigned long long per = (p.sym_count - p.sym_lo + n_teams - 1) / n_teams;
  unsigned long long s0 = p.sym_lo + team_id * per, s1 = s0 + per;
  if (s0 > p.sym_count) s0 = p.sym_count;
  if (s1 > p.sym_count) s1 = p.sym_count;

  unsigned long long acc_bit_err = 0, acc_sym_err = 0, acc_bits = 0, acc_syms = 0;
  double acc_pow = 0.0;
  double acc_zpow = 0.0;
  float acc_max = 0.f;

  float2 v[E];
  unsigned short lab[E];

  // ---- transmitter: labels -> X -> x (time domain, ortho scaled) in v[], strided layout
  auto make_tx = [&](unsigned long long srel, bool count) {
    const unsigned long long gs = p.sym_begin + srel;
    uint4 words = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int n = t + T * m;
      const float4 sc = __ldg(&p.sc_tab[n]);
      const unsigned info = __float_as_uint(sc.z);
      const int bps = int(info & 0xffu);
      unsigned l = 0;
      if (p.bits_src == SRC_PHILOX) {
        if ((m & 7) == 0) {
          words = philox4x32<10>(make_uint4((uint32_t)gs, (uint32_t)(gs >> 32), (0u << 28) | uint32_t((m >> 3) * T + t), p.point), key);
        }
        const unsigned w = (m & 4) ? ((m & 2) ? words.w : words.z) : ((m & 2) ? words.y : words.x);
        l = ((m & 1) ? (w >> 16) : (w & 0xffffu)) & ((1u << bps) - 1u);
      } else if (bps > 0) {
        const unsigned long long bitpos = srel * p.bits_per_ofdm + (info >> 8);
        const unsigned long long byte = bitpos >> 3;
        unsigned w = 0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const unsigned long long bi = byte + b;
          w = (w << 8) | (bi < p.bits_len ? (unsigned)__ldg(&p.bits[bi]) : 0u);
        }
        l = (w >> (24 - int(bitpos & 7) - bps)) & ((1u << bps) - 1u);
      }
      lab[m] = (unsigned short)l;
      v[m] = map_label(l, bps, sc.x, p.scheme);
      if (count && p.dump_tx) p.dump_tx[srel * N + n] = (unsigned short)l;
    }
    if (p.modulator == MOD_OFDM) {
      team_fft<N, E, +1>(v, buf, p.tw, team);
#pragma unroll
      for (int m = 0; m < E; ++m) v[m] = cscale(v[m], inv_sqrt_n);
    }
    if (count) {  // PAPR statistics over every tx sample, prefix included (simulation/models.py:519-522)
      float s = 0.f, mx = 0.f;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int n = t + T * m;
        const float pw = fmaf(v[m].x, v[m].x, v[m].y * v[m].y);
        const float wgt = (p.prefix_type == PREFIX_CYCLIC && n >= N - P) ? 2.f : 1.f;
        s = fmaf(wgt, pw, s);
        mx = fmaxf(mx, pw);
      }
      acc_pow += double(s);
      acc_max = fmaxf(acc_max, mx);
    }
    // publish x for the FIR's neighbour reads
    team.sync();
#pragma unroll
    for (int m = 0; m < E; ++m) buf[pad_idx(t + T * m)] = v[m];
    team.sync();
  };

  auto save_tail = [&]() {  // tail[i] = x[N - TL + i]
    for (int i = t; i < TL; i += T) tail[i] = buf[pad_idx(N - TL + i)];
    team.sync();
  };

  bool have_prev = false;
  if (p.isi && s0 < s1 && (p.sym_begin + s0) > 0) {
    make_tx(s0 - 1, false);  // halo: only its tail is needed (srel may be "-1" -> wraps, used for Philox only)
    save_tail();
    have_prev = true;
  }

  for (unsigned long long s = s0; s < s1; ++s) {
    const unsigned long long gs = p.sym_begin + s;
    make_tx(s, true);

    // tx stream sample at index i in [-(L-1), N+P) of the current symbol (prefix/models.py:34-69)
    auto stream = [&](int i) -> float2 {
      int xi;
      if (i < 0) {
        if (!have_prev) return make_float2(0.f, 0.f);
        const int ip = NP + i;  // index in the previous symbol's stream
        if (p.prefix_type == PREFIX_ZERO) {
          if (ip >= N) return make_float2(0.f, 0.f);
          xi = ip;
        } else {
          xi = N + i;  // CP: (ip - P); NONE: ip
        }
        return tail[xi - (N - TL)];
      }
      if (p.prefix_type == PREFIX_CYCLIC) {
        xi = i - P;
        if (xi < 0) xi += N;
      } else {
        if (i >= N) return make_float2(0.f, 0.f);
        xi = i;
      }
      return buf[pad_idx(xi)];
    };

    // ---- channel: r[n] = sum_l h[l] * stream(i0 - l), then prefix removal folded in
    float2 r[E];
#pragma unroll
    for (int m = 0; m < E; ++m) r[m] = make_float2(0.f, 0.f);
    if (cp_fast) {
      for (int l = 0; l < L; ++l) {
        const float2 h = p.taps[l];
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const float2 x = buf[pad_idx((t + T * m - l) & (N - 1))];
          r[m].x = fmaf(h.x, x.x, fmaf(-h.y, x.y, r[m].x));
          r[m].y = fmaf(h.x, x.y, fmaf(h.y, x.x, r[m].y));
        }
      }
    } else {
      const int off = (p.prefix_type == PREFIX_CYCLIC) ? P : 0;
      for (int l = 0; l < L; ++l) {
        const float2 h = p.taps[l];
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const int n = t + T * m;
          float2 x = stream(n + off - l);
          if (p.prefix_type == PREFIX_ZERO && n < P) x = cadd(x, stream(n + N - l));  // overlap-add (prefix/models.py:87-101)
          r[m].x = fmaf(h.x, x.x, fmaf(-h.y, x.y, r[m].x));
          r[m].y = fmaf(h.x, x.y, fmaf(h.y, x.x, r[m].y));
        }
      }
    }
    team.sync();
    if (p.isi) {
      save_tail();
      have_prev = true;
    }

    // ---- noise on the samples that survive prefix removal (noise/models.py:19-22)
    if (p.noise_src == SRC_PHILOX) {
      uint4 w = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int n = t + T * m;
        if ((m & 1) == 0)
          w = philox4x32<10>(make_uint4((uint32_t)gs, (uint32_t)(gs >> 32), (1u << 28) | uint32_t((m >> 1) * T + t), p.point), key);
        float2 g = (m & 1) ? box_muller(w.z, w.w) : box_muller(w.x, w.y);
        g = cscale(g, p.sigma);
        if (p.dump_noise) p.dump_noise[s * NP + n + (p.prefix_type == PREFIX_CYCLIC ? P : 0)] = g;
        if (p.prefix_type == PREFIX_ZERO && n < P) {
          const uint4 w2 = philox4x32<10>(make_uint4((uint32_t)gs, (uint32_t)(gs >> 32), (2u << 28) | uint32_t(n), p.point), key);
          const float2 g2 = cscale(box_muller(w2.x, w2.y), p.sigma);
          if (p.dump_noise) p.dump_noise[s * NP + n + N] = g2;
          g = cadd(g, g2);
        }
        r[m] = cadd(r[m], g);
      }
    } else if (p.noise_src == SRC_REPLAY_F32 || p.noise_src == SRC_REPLAY_F64) {
      auto load_noise = [&](unsigned long long idx) -> float2 {
        if (p.noise_src == SRC_REPLAY_F32) return __ldg(reinterpret_cast<const float2*>(p.noise) + idx);
        const double2 d = __ldg(reinterpret_cast<const double2*>(p.noise) + idx);
        return make_float2((float)d.x, (float)d.y);
      };
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const int n = t + T * m;
        const unsigned long long base = s * (unsigned long long)NP;
        float2 g = load_noise(base + n + (p.prefix_type == PREFIX_CYCLIC ? P : 0));
        if (p.prefix_type == PREFIX_ZERO && n < P) g = cadd(g, load_noise(base + n + N));
        r[m] = cadd(r[m], g);
      }
    }

    // ---- receiver: ortho FFT
    team_fft<N, E, -1>(r, buf, p.tw, team);
#pragma unroll
    for (int m = 0; m < E; ++m) r[m] = cscale(r[m], inv_sqrt_n);
    if (p.dump_y) {
#pragma unroll
      for (int m = 0; m < E; ++m) p.dump_y[s * N + t + T * m] = r[m];
    }

    // ---- equaliser (equalization/models.py:22-68)
    if (p.equalizer == EQ_MMSE) {
      float ss = 0.f;
#pragma unroll
      for (int m = 0; m < E; ++m) ss = fmaf(r[m].x, r[m].x, fmaf(r[m].y, r[m].y, ss));
      const float sigma2 = team.sum(ss) * p.mmse_c;
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const float4 e = __ldg(&p.eq_tab[t + T * m]);
        const float inv = 1.0f / (e.z + sigma2);
        const float2 yh = cmul_conj(r[m], make_float2(e.x, e.y));
        r[m] = cscale(yh, inv);
      }
    } else if (p.equalizer == EQ_ZF) {
#pragma unroll
      for (int m = 0; m < E; ++m) {
        const float4 e = __ldg(&p.eq_tab[t + T * m]);
        r[m] = cmul(r[m], make_float2(e.x, e.y));
      }
    }
    if (p.post_sigma) {  // coloured noise injected AFTER the equaliser (examples/waterfilling_noise_bump_experiment.py:163-171)
      if (p.post_src == SRC_REPLAY_F64) {
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const double2 g = __ldg(reinterpret_cast<const double2*>(p.post_noise) + s * (unsigned long long)N + t + T * m);
          r[m] = cadd(r[m], make_float2((float)g.x, (float)g.y));
        }
      } else {
        uint4 w = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int m = 0; m < E; ++m) {
          if ((m & 1) == 0)
            w = philox4x32<10>(make_uint4((uint32_t)gs, (uint32_t)(gs >> 32), (3u << 28) | uint32_t((m >> 1) * T + t), p.point), key);
          const float2 g = (m & 1) ? box_muller(w.z, w.w) : box_muller(w.x, w.y);
          r[m] = cadd(r[m], cscale(g, p.post_scale * __ldg(&p.post_sigma[t + T * m])));
        }
      }
    }
    if (p.rx_gain) {  // receiver compensation of the applied power loading (examples/waterfilling_noise_bump_experiment.py:173-176)
#pragma unroll
      for (int m = 0; m < E; ++m) r[m] = cscale(r[m], __ldg(&p.sc_tab[t + T * m]).w);
    }
    if (p.z_power) {     // block-wide mean power of what the demapper would see (:178-179): first pass of the renormalisation
      float zp = 0.f;
#pragma unroll
      for (int m = 0; m < E; ++m) zp = fmaf(r[m].x, r[m].x, fmaf(r[m].y, r[m].y, zp));
      acc_zpow += (double)zp;
    }
    if (p.z_scale != 1.0f) {   // ... and its second pass (:180-181)
#pragma unroll
      for (int m = 0; m < E; ++m) r[m] = cscale(r[m], p.z_scale);
    }
    if (p.modulator == MOD_SC) {  // SC-OFDM: back to the time domain (modulation/models.py:89)
      team_fft<N, E, +1>(r, buf, p.tw, team);
#pragma unroll
      for (int m = 0; m < E; ++m) r[m] = cscale(r[m], inv_sqrt_n);
    }
    if (p.dump_z) {
#pragma unroll
      for (int m = 0; m < E; ++m) p.dump_z[s * N + t + T * m] = r[m];
    }

    // ---- hard decisions + error counting (simulation/models.py:597-606)
    unsigned be = 0, se = 0, nb = 0, ns = 0;
#pragma unroll
    for (int m = 0; m < E; ++m) {
      const int n = t + T * m;
      const float4 sc = __ldg(&p.sc_tab[n]);
      const unsigned info = __float_as_uint(sc.z);
      const int bps = int(info & 0xffu);
      const unsigned rx = demap_point(r[m], bps, sc.y, p.scheme);
      if (p.dump_rx) p.dump_rx[s * N + n] = (unsigned short)rx;
      unsigned diff = rx ^ unsigned(lab[m]);
      se += (diff != 0u);
      ns += 1u;
      int valid = bps;
      if (p.limit_bits) {
        const unsigned long long bitpos = s * p.bits_per_ofdm + (info >> 8);
        const long long room = (long long)p.compare_limit - (long long)bitpos;
        valid = room <= 0 ? 0 : (room < bps ? int(room) : bps);
        diff >>= (bps - valid);
      }
      be += __popc(diff);
      nb += unsigned(valid);
    }
    acc_bit_err += be;
    acc_sym_err += se;
    acc_bits += nb;
    acc_syms += ns;
  }

  // ---- reduce the per-thread counters: warp shuffle, then one atomic per warp
  auto warp_sum64 = [](unsigned long long x) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    return x;
  };
  __syncwarp();
  const unsigned long long b0 = warp_sum64(acc_bit_err), b1 = warp_sum64(acc_bits), b2 = warp_sum64(acc_sym_err),
                           b3 = warp_sum64(acc_syms);
  double pw = acc_pow, zw = acc_zpow;
  float mx = acc_max;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    pw += __shfl_down_sync(0xffffffffu, pw, off);
    zw += __shfl_down_sync(0xffffffffu, zw, off);
    mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) {
    if (b0) atomicAdd(&p.counters[CNT_BIT_ERRORS], b0);
    if (b1) atomicAdd(&p.counters[CNT_BITS], b1);
    if (b2) atomicAdd(&p.counters[CNT_SYM_ERRORS], b2);
    if (b3) atomicAdd(&p.counters[CNT_SYMBOLS], b3);
    if (pw != 0.0) atomicAdd(p.tx_power_sum, pw);
    if (p.z_power) {
      atomicAdd(reinterpret_cast<double*>(&p.counters[CNT_Z_POWER]), zw);
      if (b3) atomicAdd(&p.counters[CNT_Z_VALUES], b3);
    }
    atomicMax(p.tx_power_max_bits, (unsigned long long)__double_as_longlong((double)mx));
  }
  if (t == 0 && s1 > s0) atomicAdd(&p.counters[CNT_OFDM_SYMBOLS], s1 - s0);
}

}  // namespace ofdm
