// ofdm_link_fast kernel: the Monte-Carlo hot loop, <= 8 channel taps, N = E*T subcarriers: a team of T lanes with
//   E samples per lane.  T = E in {8, 16, 32} (N = 64, 256, 1024: two-pass transform); T = 2E, 4E or 8E (N = 128, 512
//   inside a warp; N = 2048, 4096, 8192 with a team of 2, 4 or 8 warps): a third radix-2/4/8 pass follows a second exchange.
//   The headline instantiation is OFDM, one square-QAM order (4 .. 256), cyclic prefix >= channel memory, Philox bits
//   and noise; compile-time flags add, each in its own instantiation so that the headline stream stays untouched:
//     REPLAY  recorded bits and noise streamed from HBM          ADAPT   per-subcarrier orders / applied power loading
//     FRAMES  batches of channel realisations (with ADAPT)       SC      single-carrier OFDM (FFT -> EQ -> IFFT)
//     ISI     short / no prefix: chained symbols, carried tail   PSK     M-ary PSK
//     DUMP    per-symbol Y, Z, labels, noise for the parity tests
//   and run-time fields cover the zero-padding guard interval and prefixes longer than one row of samples.
// Same chain and same reference lines as link_kernel.cuh; what differs is the machine mapping:
//   * a team of T lanes (one warp for N = 1024) owns an OFDM symbol, E samples per lane in registers;
//   * ONE forward-FFT body serves both transforms (the IFFT runs as an FFT on re/im-swapped data) and the
//     FIR + AWGN stage is a rolled loop over 8-sample chunks that works in place in shared memory, so the
//     per-symbol instruction stream stays small; the warps of a block run free (SYNC = 0; SYNC = 2, round 1's choice,
//     keeps the warps that share a scheduler in step so that they share instruction-cache lines);
//   * the kernel is bound by the issue rate, so the formulation removes instructions wherever the algebra allows: the
//     inter-pass twiddles and the mapper's level offset ride in first-stage butterflies, the Gauss-form FIR sums are
//     chained, the error count works on one word per four subcarriers (kOpt* below, profiles/r2_fast_kernel_history.md);
//   * every scale factor (1/sqrt(2(M-1)/3), both 1/sqrt(N), the slicer's k/2 and 1/(s-1)) is folded on the
//     host into the FIR taps and the equaliser table; level <-> index conversions use mantissa tricks and
//     FFMA.SAT (no I2F / F2I / FMNMX);
//   * ZF and "no equaliser" run through the MMSE form conj(A) / (G + sigma2) with sigma2 = 0;
//   * bit errors are counted on packed words: gray^-1 is linear over GF(2), so the error pattern of a
//     subcarrier is gray^-1(col_tx ^ col_rx), evaluated four subcarriers per 32-bit word.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fft_regs.cuh"
#include "link_params.h"
#include "philox.cuh"

namespace ofdm {

constexpr int kFastTaps = 8;

struct FrameHeader {       // per channel realisation (built on the device by frames.cu)
  float2 taps[kFastTaps];  // unit-energy taps / sqrt(N)
  float4 taps3[kFastTaps]; // {h_re, h_im - h_re, h_re + h_im, -} of the same taps (FastParams::taps3)
  float sigma;             // per-component noise standard deviation
  float mmse_c;            // as FastParams::mmse_c
  float pad[2];
};

constexpr int kMaxSweepPoints = 32;   // SNR points per launch (the parameter block carries their table)

struct SweepPoint {          // one SNR point of a launch (grid y index)
  float sigma;              // per-component noise standard deviation
  float mmse_c;             // MMSE: sigma2 = mmse_c * sum_k |Y~_k|^2 (Y~ = unscaled FFT output); else 0
};

struct FastParams {
  float2 taps[kFastTaps];   // unit-energy taps / (sqrt(2(M-1)/3) * sqrt(N))   (levels are 2c-(s-1), IFFT unscaled)
  float4 taps3[kFastTaps];  // the same taps as {h_re, h_im - h_re, h_re + h_im, -}: three real products per complex one
  const float4* eq_tab;     // {Re A, Im A, G, -}: decision = sat(Re/Im(Y~ conj A) / (G + sigma2) + 0.5) * (s-1)
  const float2* tw;         // pass-2 twiddles exp(-2 pi i k r / E^2) at [k*(E+2) + r-1] (kOptFusedTwiddle: r = n at [k*(E+2) + 2n],
                            // r = n + E/2 at [k*(E+2) + 2n + 1], n < E/2), then (T > E) the pass-3 base
                            // twiddles exp(-2 pi i j / N), j < N / (T/E)   (build_fast_twiddles, link_fast.cu)
  float slice_top;          // s-1
  float tx_scale2;          // |tx|^2 = tx_scale2 * |x~|^2 (PAPR statistics)
  float z_unscale;          // DUMP only: Z = (Y~ conj A / (G + sigma2)) * z_unscale   (= 2 (s-1) / k)
  int prefix_len;
  int zero_prefix;          // 1: zero-padding guard interval (prefix/models.py:60-101), folded by overlap-add; 0: cyclic prefix
  int equalizer;
  int half_bits;            // log2(s)
  unsigned int field_mask;  // (s-1) << 1 replicated in every byte
  unsigned int k4b = 0x4B000000u;   // bits of 2^23, read from the constant bank so that PRMT keeps its byte selector as the immediate
  unsigned long long seed;
  unsigned int point;
  // n_points >= 1 SNR points in one grid: the blocks with blockIdx.y = pt run point `point + pt` with point_tab[pt] and
  // add into counters + 10 * pt (one CounterBlock per point)
  unsigned int n_points;
  SweepPoint point_tab[kMaxSweepPoints];
  unsigned long long sym_begin, sym_count;
  unsigned long long* counters;   // CounterBlock(s): 8 counters, power sum (double), power max (double bits)
  float2* dump_z;
  unsigned short* dump_rx;
  unsigned short* dump_tx;
  float2* dump_noise;
  float2* dump_y;           // DUMP only: ortho-scaled FFT output before the equaliser
  float y_scale;            // 1 / sqrt(N)
  // REPLAY instantiation: the reference's own byte stream (MSB first) and complex64 noise over the serial
  // stream, (N + P) samples per OFDM symbol (bits_generation/models.py:27-55, noise/models.py:19-22)
  const unsigned char* bits;
  unsigned long long bits_len;
  // ADAPT instantiation (per-subcarrier QAM orders, constellation/adaptive.py:52-201):
  const unsigned short* bit_offsets; // [N]: bit offset of subcarrier k inside an OFDM symbol (ADAPT + REPLAY)
  unsigned int bits_per_ofdm;        // sum of the per-subcarrier bits (ADAPT + REPLAY)
  const unsigned int* field_masks;   // [(E/4) * T]: word j of lane t = ((s_k - 1) << 1) in byte i for k = t + T (4 j + i)
  const float2* level_tab;           // [N]: {g_k, -(2^23 + s_k - 1)} with g_k = 1 / sqrt(2 (M_k - 1) / 3)  (0 when silent);
                                     // the slicer's s_k - 1 is the 4th component of eq_tab
  // PSK instantiation (constellation/models.py:356-474): labels of psk_bits bits, one per byte of the packed words
  const float2* psk_tab;    // [256]: label -> exp(j 2 pi gray^-1(label) / M)
  int psk_bits;             // log2 M, 1 .. 8
  float psk_scale;          // M / (2 pi)
  // FRAMES instantiation (always with ADAPT): a batch of channel realisations, `frame_syms` OFDM symbols each.
  // eq_tab / level_tab / field_masks hold one table per frame ([F][N], [F][N], [F][N/4]); a (frame, chunk) unit is
  // processed by one block, which loads the frame's tables into shared memory when the frame changes.
  const FrameHeader* frame_hdr;        // [F]
  unsigned long long* frame_counters;  // [F][10]: 8 counters, power sum (double), power max (double bits)
  unsigned long long frame_syms;       // OFDM symbols per frame
  unsigned int n_frames, chunks_per_frame, chunk_syms;
  const void* noise;        // complex64 (noise_f64 = 0) or complex128 (noise_f64 = 1); NULL = noiseless
  int noise_f64;
};

template <int E, int T_ = E, int BLOCK_ = 512>
struct FastGeometry {
  static_assert(T_ % E == 0 && (T_ / E == 1 || T_ / E == 2 || T_ / E == 4 || T_ / E == 8) && (T_ <= 32 || T_ % 32 == 0), "team shape");
  static constexpr int T = T_;                // lanes per OFDM symbol
  static constexpr int N = E * T;
  static constexpr int W = T / E;             // radix of the third pass (1: two-pass transform)
  // the equaliser table (16 N bytes) sits in shared memory beside the teams' buffers up to N = 4096; at N = 8192 (two teams of
  // eight warps, 139 KB of sample buffers) it is read through the read-only data path instead
  static constexpr bool EQ_SMEM = N <= 4096;
  static constexpr int RS = E + 2;            // row stride (complex): conflict-free 128-bit row accesses
  // float2 per team: T rows of E samples; teams narrower than a half-warp are offset by half a bank cycle (64 B) so that
  // the 64-bit column accesses of the two teams of a half-warp fall on different banks (ncu: 2-way conflicts on every
  // column access at N = 64 without the offset, profiles/r2_small_n.md)
  static constexpr int TEAM_F2 = T * RS + ((T < 16 && (T * RS) % 16 == 0) ? 8 : 0);
  static constexpr int BLOCK = BLOCK_;
  static constexpr int TEAMS = BLOCK / T;
  static constexpr int TW2_F2 = E * RS;       // pass-2 twiddles (float2): one padded row per lane column, read as 128-bit pairs
  static constexpr int TW3_F2 = W > 1 ? N / W : 0;
  static constexpr int TW_F2 = TW2_F2 + TW3_F2;
  static constexpr int RED_F = T > 32 ? TEAMS * (T / 32) : 0;   // cross-warp reduction scratch (floats)
  static constexpr int TAIL_F2 = 8;           // ISI: last tx samples of the previous OFDM symbol, per team
  static constexpr int PSK_F2 = 256;          // PSK: point table
  static constexpr size_t SMEM_BYTES = (size_t(TEAMS) * TEAM_F2 + TW_F2) * sizeof(float2) + size_t(EQ_SMEM ? N : 0) * sizeof(float4) +
                                       RED_F * sizeof(float) + (size_t(TEAMS) * TAIL_F2 + PSK_F2) * sizeof(float2);
};

// barrier among the lanes of one team: the warp when the team fits one, else a named barrier (ids 5..12)
template <int T>
__device__ __forceinline__ void team_sync(int team_in_block) {
  if constexpr (T <= 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(5 + team_in_block), "n"(T) : "memory");
  }
}

// Alignment of the warps of a block at the section boundaries of the symbol loop (transmitter, channel, receiver):
// SYNC = 0: warps run free.  SYNC = 1: __syncthreads().  SYNC = 2: named barrier among the warps that share a
// scheduler (warp id mod 4), so that they fetch the same instruction-cache lines.  (Grouping one warp per scheduler
// instead - so that every scheduler mixes FMA-heavy, Philox and MUFU sections - was measured 1.3 % slower.)
template <int SYNC, int BLOCK>
__device__ __forceinline__ void section_sync() {
  if constexpr (SYNC == 1) {
    __syncthreads();
  } else if constexpr (SYNC == 2) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + ((threadIdx.x >> 5) & 3)), "n"(BLOCK / 4) : "memory");
  }
}

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool V>
__device__ __forceinline__ float4 lds128(const float2* ptr) {
  if constexpr (V) {
    float4 r;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "r"((unsigned)__cvta_generic_to_shared(ptr)));
    return r;
  } else {
    return *reinterpret_cast<const float4*>(ptr);
  }
}
__device__ __forceinline__ float2 lds64v(const float2* ptr) {
  float2 r;
  asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"((unsigned)__cvta_generic_to_shared(ptr)));
  return r;
}

// circularly-symmetric sigma * (N(0,1) + j N(0,1)) from two 32-bit words (same distribution as box_muller())
__device__ __forceinline__ float2 fast_box_muller(uint32_t wr, uint32_t wa, float sigma) {
  const float u1 = fmaf((float)wr, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float rad = sigma * fast_sqrt(-1.3862943611198906f * fast_lg2(u1));  // sigma * sqrt(-2 ln u)
  const float ang = (float)(int32_t)wa * 1.4629180792671596e-09f;            // (-pi, pi)
  return make_float2(rad * __cosf(ang), rad * __sinf(ang));
}

// sigma * (N(0,1) + j N(0,1)) from a 32-bit radius word and a 16-bit angle field (bytes picked by `sel`):
//   radius = sqrt(-2 sigma^2 ln u), u = (w + 0.5) 2^-32  -> tail out to 6.6 sigma;
//   angle  = 2 pi a / 65536 - pi via the mantissa of 2^23 + a (noise is rotation invariant, 16 bits suffice)
__device__ __forceinline__ float2 fast_noise(uint32_t wr, uint32_t wa, uint32_t sel, float c2 /* -2 sigma^2 ln 2 */) {
  const float u1 = fmaf((float)wr, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float rad = fast_sqrt(c2 * fast_lg2(u1));
  const float f = __uint_as_float(__byte_perm(wa, 0x4B000000u, sel));           // 2^23 + a
  const float ang = fmaf(f, 9.587379924285257e-05f, -807.3893119735271f);       // (2 pi / 65536) a - pi
  return make_float2(rad * __cosf(ang), rad * __sinf(ang));
}

// Same distribution from ONE 32-bit word: radius field = top 20 bits, angle field = low 12 bits.
//   u = (w | 0xFFF) 2^-32 in (0, 1]: radius^2 = c2 (lg2(w | 0xFFF) - 32), c2 = -2 sigma^2 ln 2, c2m = -32 c2 (biased by
//   2^-18 so that the approximate lg2 cannot produce a negative argument).  A radius field of 0 (probability 2^-20) only
//   says u < 2^-20: noise_refill() then replaces the sample by one whose u comes from 32 fresh bits, so the tail
//   continues to u = 2^-53 (8.6 sigma) like a generator with a 52-bit radius word.
//   angle = 2 pi (a + 0.5) / 4096 - pi via the mantissa of 2^23 + a: 4096 equally spaced rays are invisible after any
//   projection (the radius is continuous) and noise is rotation invariant.
constexpr uint32_t kRefillBelow = 4096u;   // w < 4096  <=>  radius field == 0
__device__ __forceinline__ float noise20_radius(uint32_t w, float c2, float c2m) {
  // the OR is not optional: u straight from the word saves it, but w = 0 (once per 2^32 samples) then gives an infinite
  // radius that poisons its whole OFDM symbol - bench.py's 1e12-bit record caught 29 such symbols
  return fast_sqrt(fmaf(c2, fast_lg2((float)(w | 0xFFFu)), c2m));
}
__device__ __forceinline__ float2 noise20_dir(uint32_t w) {
  const float f = __uint_as_float((w & 0xFFFu) | 0x4B000000u);                                   // 2^23 + a
  const float ang = fmaf(f, 1.5339807878856412e-03f, -12871.104692850697f);                      // (2 pi / 4096)(a + 0.5) - pi (offset for the ROUNDED slope)
  return make_float2(__cosf(ang), __sinf(ang));
}
__device__ __forceinline__ float2 fast_noise20(uint32_t w, float c2, float c2m) {
  const float rad = noise20_radius(w, c2, c2m);
  const float2 d = noise20_dir(w);
  return make_float2(rad * d.x, rad * d.y);
}

// Rare path of the 32-bit-per-sample noise: at least one of this lane's E radius fields was 0.  Regenerates the lane's
// words, and for every sample with a zero field draws 32 fresh bits r: u = (r + 0.5) 2^-52, same direction; the FIR
// output in `row` already holds the coarse sample, so the difference is added.
template <int E, int NROUNDS>
__device__ __noinline__ void noise_refill(float2* row, int t, uint32_t gs_lo, uint32_t gs_hi, uint32_t point, PhiloxKey key,
                                          float c2, float c2m, float2* dump_noise, unsigned long long dump_base) {
#pragma unroll 1
  for (int c = 0; c < E / 8; ++c) {
    const uint32_t q2 = 2u * uint32_t((E / 8) * t + c);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const uint4 w4 = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | (q2 + h), point), key);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t w = j == 0 ? w4.x : j == 1 ? w4.y : j == 2 ? w4.z : w4.w;
        if (w >= kRefillBelow) continue;
        const int i = 8 * c + 4 * h + j;
        const uint4 r = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | (1u << 20) | uint32_t(E * t + i), point), key);
        const float u = fmaf((float)r.x, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (r + 0.5) 2^-32
        const float rad_new = fast_sqrt(c2 * (fast_lg2(u) - 20.0f));
        const float rad_old = noise20_radius(w, c2, c2m);
        const float2 d = noise20_dir(w);
        const float2 o = row[i];
        row[i] = make_float2(fmaf(rad_new - rad_old, d.x, o.x), fmaf(rad_new - rad_old, d.y, o.y));
        if (dump_noise) dump_noise[dump_base + i] = make_float2(rad_new * d.x, rad_new * d.y);
      }
    }
  }
}

// Zero-padding guard interval, overlap-add at the receiver (prefix/models.py:71-101): the noise of tail sample N + n
// lands on sample n < P.  `row` holds this lane's samples E t .. E t + E - 1; one Philox call per folded sample.
template <int E, int NROUNDS>
__device__ __noinline__ void zero_prefix_tail_noise(float2* row, int t, int P, uint32_t gs_lo, uint32_t gs_hi, uint32_t point,
                                                    PhiloxKey key, float noise_c2, float2* dump_noise, unsigned long long dump_base) {
  for (int n = E * t; n < P && n < E * t + E; ++n) {
    const uint4 w = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (2u << 28) | (uint32_t)n, point), key);
    const float2 g = fast_noise(w.x, w.y, 0x7610u, noise_c2);
    if (dump_noise) dump_noise[dump_base + n] = g;
    row[n - E * t] = cadd(row[n - E * t], g);
  }
}

// prefix-XOR inside the 4-bit fields that sit at bits 1..4 of every byte (inverse Gray code)
__device__ __forceinline__ unsigned inv_gray_fields(unsigned x) {
  x ^= x >> 1;
  x ^= x >> 2;
  return x;
}

// OPT bits (measured one by one on the headline shape, profiles/r2_fast_kernel_history.md):
//   1  noise from 32 bits per complex sample (2 Philox calls per 8 samples + rare refill) instead of 48 (3 calls)
//   2  FIR with three real products per complex tap (Gauss) instead of four
//   4  no per-symbol noise estimate: ZF / no equaliser, where sigma2 = 0 (the MMSE form then needs neither the power sum
//      over the spectrum nor its shuffles; -3 instructions per subcarrier)
//   16..240  depth of the twiddle ring of the transform's exchange (bits 4..7; 0 = load each pair at its use)
//   1024 pass-2 twiddles fused into the first butterflies of the second codelet: (a, c) <- (w_a a + w_c c, w_a a - w_c c)
//        costs 10 instructions instead of 12 (table layout: FastParams::tw)
//   2048 level offset -(s-1) folded into the first butterflies of the transmitter's first codelet: the difference of two
//        2^23-based label floats needs no offset, their sum one (6 FADD per butterfly instead of 8)
//   8192 PAPR maximum with 3-input integer maxima on the bit patterns (non-negative floats order like integers)
//   16384 the data-bit Philox calls use NROUNDS like the noise calls
//   32768 Gauss-form FIR without the two final additions: the k3 and k2 sums continue the k1 chain,
//         y_re = k1 - sum (h_re + h_im) x_im, y_im = k1 + sum (h_im - h_re) x_re
//   65536 error count on ONE word per 4 subcarriers: the column and row difference fields side by side in a byte
//         ((dc >> 1) | (dr << 3)), inverse Gray code with nibble-isolating masks
constexpr int kOptNoise32 = 1, kOptGaussFir = 2, kOptNoEstimate = 4, kOptFusedTwiddle = 1024,
              kOptFusedLevels = 2048, kOptIntMax = 8192, kOptDataRounds = 16384, kOptFirChain = 32768, kOptJointGray = 65536;
constexpr int kOptTwRing4 = (4 << 4) | 256;   // twiddle ring of 4 buffers, volatile loads
// what every product instantiation uses (the twiddle table of a link is laid out for kOptFusedTwiddle), the one-tap and the
// multi-tap formulation of the channel; profiles/r2_fast_kernel_history.md section 7 has the measurements, the A/B harness
// (tools/microbench/fast_variants.cu) still builds the formulations without them
constexpr int kOptCommon = kOptFusedTwiddle | kOptFusedLevels | kOptIntMax | kOptDataRounds | kOptFirChain |
                           kOptJointGray | kOptTwRing4;
constexpr int kOptOneTap = kOptNoise32 | kOptCommon, kOptDefault = kOptNoise32 | kOptGaussFir | kOptCommon;

// Philox4x32 rounds of the fast kernel's bit and noise streams.  7 is the fewest rounds with which Philox4x32 passes
// BigCrush (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11, table 2: "crush-resistant");
// 10 is that paper's default with a safety margin (-DOFDM_FAST_PHILOX_ROUNDS=10 builds the library with it; the general
// kernel always uses 10).  Three rounds fewer are 2.5 % of the headline kernel's time.
#ifndef OFDM_FAST_PHILOX_ROUNDS
#define OFDM_FAST_PHILOX_ROUNDS 7
#endif
constexpr int kFastRounds = OFDM_FAST_PHILOX_ROUNDS;

template <int E, int T, bool DUMP, bool PAPR, bool REPLAY = false, int BLOCK = 512, int SYNC = 2, bool ADAPT = false,
          bool FRAMES = false, bool SC = false, bool ISI = false, bool PSK = false, int NROUNDS = kFastRounds, int FIR_UNROLL = 2,
          int TAPS = kFastTaps, int OPT = kOptDefault, int MINB = 1>
__global__ void __launch_bounds__(BLOCK, MINB) ofdm_link_fast_kernel(const __grid_constant__ FastParams p) {
  // MINB: blocks per SM the register allocation must allow (2 for the narrow codelets: 64 registers, 32 warps per SM)
  // TAPS: channel taps the FIR evaluates (the host zero-pads the tap table, so a shorter loop only drops exact zeros)
  static_assert(TAPS >= 1 && TAPS <= kFastTaps && (TAPS == kFastTaps || !FRAMES), "tap count");
  constexpr bool NOISE32 = (OPT & kOptNoise32) != 0, GAUSS = (OPT & kOptGaussFir) != 0;
  constexpr int TWR = (OPT >> 4) & 15;        // depth of the twiddle ring of the exchange (0: load at use)
  constexpr bool FTW = (OPT & kOptFusedTwiddle) != 0;                             // see kOptFusedTwiddle
  constexpr bool FLV = (OPT & kOptFusedLevels) != 0 && !ADAPT && !PSK && !SC;     // see kOptFusedLevels
  constexpr bool IMAX = (OPT & kOptIntMax) != 0;
  constexpr int DROUNDS = (OPT & kOptDataRounds) ? NROUNDS : 10;
  constexpr bool TWV = (OPT & 256) != 0;       // ring and column loads as volatile accesses (keeps their order in SASS)
  // PSK: M-ary phase-shift keying, one order on every subcarrier; labels through a shared-memory point table at the
  // transmitter, angle rounding at the receiver (the equaliser's positive real denominator does not move the angle)
  static_assert(!PSK || (!ADAPT && !FRAMES), "PSK: one order, single link");
  // ISI: cyclic prefix shorter than the channel memory, or no prefix (channel/models.py:52-55 over the serial stream):
  // the FIR reaches into the previous OFDM symbol.  A team then owns a CONTIGUOUS chain of symbols, carries the last 8
  // tx samples of the previous one in shared memory, and recomputes the transmitter of the symbol before its chain
  // (one "halo" pass), so the result does not depend on how the symbol range is partitioned.
  static_assert(!ISI || !FRAMES, "ISI chains: single link");
  // SC: single-carrier OFDM (modulation/models.py:58-91) - the constellation symbols are the time samples; the
  // receiver runs FFT -> equaliser -> IFFT, i.e. the shared transform body serves phases 1 and 2 instead of 0 and 1
  static_assert(!SC || (!ADAPT && !FRAMES), "SC-OFDM: one order on every sample, single link");
  static_assert(!FRAMES || (ADAPT && !DUMP && !REPLAY), "frame batches: fused mode with per-frame tables");
  using G = FastGeometry<E, T, BLOCK>;
  constexpr int N = G::N, RS = G::RS, WORDS = E / 4, W = G::W;
  constexpr int CALLS = (E + 15) / 16;  // Philox calls for E random bytes
  extern __shared__ float4 smem4[];
  const int lane = threadIdx.x & 31;
  const int t = threadIdx.x % T;
  const int team_in_block = threadIdx.x / T;
  // linear sample / subcarrier index i lives at row i / E, column i % E; this lane's strided set is
  // i = t + T m  ->  row (t / E) + W m, column t % E
  const int tcol = t % E, trow = t / E;
  float2* smem2 = reinterpret_cast<float2*>(smem4);
  float2* buf = smem2 + size_t(team_in_block) * G::TEAM_F2;
  float2* row = buf + t * RS;
  float2* s_tw = smem2 + size_t(G::TEAMS) * G::TEAM_F2;
  float4* s_eq = reinterpret_cast<float4*>(s_tw + G::TW_F2);
  float* s_red = reinterpret_cast<float*>(s_eq + (G::EQ_SMEM ? N : 0)) + team_in_block * (T / 32);
  float2* s_tail = reinterpret_cast<float2*>(reinterpret_cast<float*>(s_eq + (G::EQ_SMEM ? N : 0)) + G::RED_F) + team_in_block * G::TAIL_F2;
  float2* s_psk = s_tail - team_in_block * G::TAIL_F2 + G::TEAMS * G::TAIL_F2;
  float2* col = buf + trow * RS + tcol;   // strided set: col[W * RS * m]
  auto tsync = [&]() { team_sync<T>(team_in_block); };

  // block-resident copies of the twiddle and equaliser tables
  for (int i = threadIdx.x; i < G::TW_F2; i += BLOCK) s_tw[i] = __ldg(&p.tw[i]);
  if constexpr (!FRAMES) {
    for (int i = threadIdx.x; i < (G::EQ_SMEM ? N : 0); i += BLOCK) s_eq[i] = __ldg(&p.eq_tab[i]);
  }
  if constexpr (PSK) {
    for (int i = threadIdx.x; i < G::PSK_F2; i += BLOCK) s_psk[i] = __ldg(&p.psk_tab[i]);
  }
  __syncthreads();

  const PhiloxKey key{(uint32_t)p.seed, (uint32_t)(p.seed >> 32)};
  const int P = p.prefix_len;
  // With a guard interval at least as long as the channel memory, a zero-padded symbol folded by overlap-add sees the
  // same circular convolution as a cyclic-prefixed one.  What differs: no prefix power in the PAPR statistics (Pc = 0),
  // the noise of the P folded tail samples adds to the first P samples, and the stream index of sample n is n, not P + n.
  const int Pc = p.zero_prefix ? 0 : P, noise_off = p.zero_prefix ? 0 : P;
  // Narrow teams (E = 8): a cyclic prefix of the usual length covers more than one row of samples (BASELINE config #1:
  // N = 64, CP = 16), so the power weights of this lane's samples - 2 for the ones the prefix repeats, n = t + T m >= N - Pc -
  // live in E registers and the sum is one FFMA per sample, whatever the prefix length
  constexpr bool PWGT = PAPR && E <= 8;
  [[maybe_unused]] float pwgt[PWGT ? E : 1];
  if constexpr (PWGT) {
#pragma unroll
    for (int m = 0; m < E; ++m) pwgt[m] = (t + T * m >= N - Pc) ? 2.f : 1.f;
  }
  const float magic = 8388608.0f;  // 2^23
  // symbols of this pass: s = s_lo + it * s_stride + s_first < s_hi, global index sym_base + s.  One pass over the
  // launch's range, or (FRAMES) one pass per (frame, chunk) unit with the whole block on the same frame.
  unsigned long long s_lo = 0, s_hi = p.sym_count, sym_base = p.sym_begin;
  // the SNR point of this block is the grid's y index (block-uniform: everything derived from it lives in uniform
  // registers / the constant bank); a single-point launch is a sweep of one
  const uint32_t point = p.point + blockIdx.y;
  float mmse_c = FRAMES ? 0.f : p.point_tab[blockIdx.y].mmse_c;
  // 32-bit team indices and iteration counts: the host splits launches so that a team sees fewer than 2^32 symbols
  unsigned s_stride = gridDim.x * G::TEAMS;
  unsigned s_first = blockIdx.x * G::TEAMS + team_in_block;
  float noise_c2 = FRAMES ? 0.f : -1.3862943611198906f * p.point_tab[blockIdx.y].sigma * p.point_tab[blockIdx.y].sigma;   // -2 sigma^2 ln 2
  [[maybe_unused]] float noise_c2m = -32.000003814697266f * noise_c2;   // see noise20_radius()
  const float2* level_tab = p.level_tab;
  [[maybe_unused]] const float4* eq_g = p.eq_tab;   // !EQ_SMEM: this block's equaliser table in global memory
  __shared__ FrameHeader s_hdr;
  [[maybe_unused]] unsigned long long unit = blockIdx.x;
  [[maybe_unused]] long long cur_frame = -1;

  unsigned long long acc_bit_err = 0, acc_sym_err = 0, acc_syms = 0;
  double acc_pow = 0.0;
  float acc_max = 0.f;

  // ADAPT: this lane's packed field masks; bits per OFDM symbol carried by its E subcarriers
  unsigned fmask[ADAPT ? WORDS : 1];
  unsigned lane_bits = PSK ? E * p.psk_bits : E * 2 * p.half_bits;
  auto load_masks = [&](const unsigned* masks) {
    lane_bits = 0;
#pragma unroll
    for (int j = 0; j < WORDS; ++j) {
      fmask[ADAPT ? j : 0] = __ldg(&masks[j * T + t]);
      lane_bits += 2 * __popc(fmask[ADAPT ? j : 0]);
    }
  };
  if constexpr (ADAPT && !FRAMES) load_masks(p.field_masks);

  // recorded bits of the next OFDM symbol, prefetched one symbol ahead: the words [first, first + n_words) that hold its
  // bits (word-aligned symbols with one order; any bit offset with per-subcarrier orders, hence one spare register)
  constexpr int PF = REPLAY ? WORDS + (ADAPT ? 1 : 0) : 1;
  unsigned next_bits[PF];
  auto replay_prefetch = [&](unsigned long long sn) {
    if constexpr (REPLAY) {
      if (sn >= p.sym_count) return;
      unsigned long long first;
      int n_words;
      if constexpr (ADAPT) {
        const unsigned long long bit0 = sn * (unsigned long long)p.bits_per_ofdm;
        first = bit0 >> 5;
        n_words = (int)(((bit0 & 31) + p.bits_per_ofdm + 31) >> 5);
      } else {
        n_words = N * (PSK ? p.psk_bits : 2 * p.half_bits) / 32;
        first = sn * (unsigned long long)n_words;
      }
      const unsigned* src = reinterpret_cast<const unsigned*>(p.bits) + first;
      const unsigned long long avail = (p.bits_len >> 2) > first ? (p.bits_len >> 2) - first : 0;   // whole words in the stream
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int w = t + T * j;
        if (w < n_words) {
          if ((unsigned long long)w < avail) {
            next_bits[j] = __ldg(src + w);
          } else {   // the stream ends inside this word (byte granular)
            unsigned word = 0;
            for (int b = 0; b < 4; ++b) {
              const unsigned long long byte = (first + w) * 4ull + b;
              if (byte < p.bits_len) word |= (unsigned)__ldg(p.bits + byte) << (8 * b);
            }
            next_bits[j] = word;
          }
        }
      }
      if (p.noise) {
        const int nbytes = (p.noise_f64 ? 16 : 8) * (N + P);
        const char* nz = reinterpret_cast<const char*>(p.noise) + sn * (unsigned long long)nbytes;
        for (int l = t * 128; l < nbytes; l += T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nz + l));
      }
    }
  };
  if constexpr (!ISI) replay_prefetch((unsigned long long)s_first);

  auto flush_counters = [&](unsigned long long* counters, double* power_sum, unsigned long long* power_max_bits) {
    // ---- counters: warp shuffle, one atomic per warp
    auto warp_sum64 = [](unsigned long long x) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
      return x;
    };
    const unsigned long long b0 = warp_sum64(acc_bit_err), b2 = warp_sum64(acc_sym_err), b3 = warp_sum64(acc_syms);
    const unsigned long long bits_total = warp_sum64((acc_syms / E) * lane_bits);
    const unsigned long long b4 = warp_sum64(t == 0 ? acc_syms / E : 0ull);   // OFDM symbols: one lane per team counts
    double pw = acc_pow;
    float mx = acc_max;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      pw += __shfl_down_sync(0xffffffffu, pw, off);
      mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) {
      if (b0) atomicAdd(&counters[CNT_BIT_ERRORS], b0);
      if (b3) {
        atomicAdd(&counters[CNT_BITS], bits_total);
        atomicAdd(&counters[CNT_SYMBOLS], b3);
        if (b4) atomicAdd(&counters[CNT_OFDM_SYMBOLS], b4);
      }
      if (b2) atomicAdd(&counters[CNT_SYM_ERRORS], b2);
      if (PAPR) {
        atomicAdd(power_sum, pw * double(p.tx_scale2));
        atomicMax(power_max_bits, (unsigned long long)__double_as_longlong(double(mx) * double(p.tx_scale2)));
      }
    }
  };

  do {
  if constexpr (FRAMES) {
    if (unit >= (unsigned long long)p.n_frames * p.chunks_per_frame) break;
    const unsigned long long f = unit / p.chunks_per_frame, c = unit - f * p.chunks_per_frame;
    if ((long long)f != cur_frame) {
      // the whole block moves to frame f: its equaliser table and header into shared memory, masks into registers
      __syncthreads();
      for (int i = threadIdx.x; i < (G::EQ_SMEM ? N : 0); i += BLOCK) s_eq[i] = __ldg(&p.eq_tab[f * N + i]);
      eq_g = p.eq_tab + f * N;
      if (threadIdx.x < (int)(sizeof(FrameHeader) / sizeof(float)))
        reinterpret_cast<float*>(&s_hdr)[threadIdx.x] = __ldg(reinterpret_cast<const float*>(&p.frame_hdr[f]) + threadIdx.x);
      __syncthreads();
      load_masks(p.field_masks + f * (N / 4));
      level_tab = p.level_tab + f * N;
      noise_c2 = -1.3862943611198906f * s_hdr.sigma * s_hdr.sigma;
      noise_c2m = -32.000003814697266f * noise_c2;
      mmse_c = s_hdr.mmse_c;
      cur_frame = (long long)f;
    }
    s_lo = c * p.chunk_syms;
    s_hi = s_lo + p.chunk_syms < p.frame_syms ? s_lo + p.chunk_syms : p.frame_syms;
    s_stride = G::TEAMS;
    s_first = team_in_block;
    sym_base = p.sym_begin + f * p.frame_syms;
  }
  unsigned iters = s_hi > s_lo ? (unsigned)((s_hi - s_lo + s_stride - 1) / s_stride) : 0u;
  // ISI: contiguous chain [chain_lo, chain_hi) per team, preceded by the halo pass of symbol chain_lo - 1
  [[maybe_unused]] unsigned long long chain_lo = 0, chain_hi = 0;
  [[maybe_unused]] bool has_prev = false;
  if constexpr (ISI) {
    const unsigned long long per = iters;                 // = ceil(sym_count / teams)
    chain_lo = (unsigned long long)s_first * per < s_hi ? (unsigned long long)s_first * per : s_hi;
    chain_hi = chain_lo + per < s_hi ? chain_lo + per : s_hi;
    has_prev = chain_lo < chain_hi && sym_base + chain_lo > 0;
    iters = (unsigned)per + 1u;
    if (t < G::TAIL_F2) s_tail[t] = make_float2(0.f, 0.f);
    replay_prefetch(has_prev ? chain_lo - 1 : chain_lo);
  }
  for (unsigned it = 0; it < iters; ++it) {
    const bool halo = ISI && it == 0;
    const unsigned long long s = ISI ? (halo ? (has_prev ? chain_lo - 1 : chain_lo) : chain_lo + it - 1)
                                     : s_lo + (unsigned long long)it * s_stride + s_first;
    const bool active = ISI ? (!halo && s < chain_hi) : s < s_hi;
    const unsigned long long gs = sym_base + ((active || (halo && has_prev)) ? s : s_lo);
    const uint32_t gs_lo = (uint32_t)gs, gs_hi = (uint32_t)(gs >> 32);

    unsigned txc[WORDS], txr[WORDS];  // transmitted level indices, 2*index at bits 1..4 of each byte
    float2 v[E];

    // ---- transmitter epilogue: PAPR statistics of the time samples x[t + T m] = sample(m) and their publication
    //      in shared memory for the FIR (prefix/models.py:34-44, simulation/models.py:519-524)
    auto tx_epilogue = [&](auto&& sample) {
        float ssum[4] = {0.f, 0.f, 0.f, 0.f}, smax[4] = {0.f, 0.f, 0.f, 0.f};   // 4 chains: latency, not issue
        [[maybe_unused]] unsigned imax[2] = {0u, 0u};
        [[maybe_unused]] float pw_prev = 0.f;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const float2 x = sample(m);
          if constexpr (PAPR) {
            const float pw = fmaf(x.x, x.x, x.y * x.y);
            // the cyclic prefix repeats the last P samples; for P <= T: n = t + T m >= N - P  <=>  m == E-1, t >= T - P
            if constexpr (PWGT) {
              ssum[m & 3] = fmaf(pw, pwgt[PWGT ? m : 0], ssum[m & 3]);
            } else {
              ssum[m & 3] += (m == E - 1 && t >= T - Pc) ? 2.f * pw : pw;
            }
            if constexpr (IMAX) {
              if (m & 1) imax[(m >> 1) & 1] = __vimax3_u32(imax[(m >> 1) & 1], __float_as_uint(pw_prev), __float_as_uint(pw));
              pw_prev = pw;
            } else {
              smax[m & 3] = fmaxf(smax[m & 3], pw);
            }
          }
          col[W * RS * m] = x;
        }
        if constexpr (PAPR && IMAX) smax[0] = __uint_as_float(max(imax[0], imax[1]));
        if (PAPR && !PWGT && Pc > T) {
          // long prefix (more than one row of samples): the rows above the last one that it also repeats; rolled
          // loop over this lane's own samples in shared memory (rare shape, keeps the instruction stream small)
          const int pm = (N - Pc) / T, pt = (N - Pc) % T;
#pragma unroll 1
          for (int m = pm; m < E - 1; ++m) {
            const float2 o = col[W * RS * m];
            if (m > pm || t >= pt) ssum[0] += fmaf(o.x, o.x, o.y * o.y);
          }
        }
        if (PAPR && active) {
          acc_pow += double((ssum[0] + ssum[1]) + (ssum[2] + ssum[3]));
          acc_max = fmaxf(acc_max, fmaxf(fmaxf(smax[0], smax[1]), fmaxf(smax[2], smax[3])));
        }
        tsync();
    };

    // OFDM: rolled, ONE copy of the transform body serves both phases (instruction cache).  SC-OFDM: unrolled, so
    // that the hand-over of the equalised spectrum in registers between phases 1 and 2 has exact live ranges.
#pragma unroll(SC ? 3 : 1)
    for (int phase = 0; phase < (halo ? 1 : (SC ? 3 : 2)); ++phase) {
      section_sync<SYNC, BLOCK>();
      if (phase == 0) {
        // ---- bits -> QAM levels (constellation/models.py:180-249). One random byte per subcarrier:
        //      low nibble -> column (in-phase) index, high nibble -> row (quadrature) index.
        if constexpr (REPLAY) {
          // the symbol's N*bps/8 bytes -> shared memory as big-endian words (coalesced loads); label k = t + T m is
          // the bps bits at bit offset k*bps, MSB first (constellation/models.py:226-243): a funnel shift over two
          // words.  column = gray(label & (s-1)), row = gray(label >> log2 s); gray is applied on packed words.
          const int bps = PSK ? p.psk_bits : 2 * p.half_bits;
          // ADAPT: the symbol starts at any bit of the stream and every subcarrier has its own width and offset
          const int intra = ADAPT ? (int)((s * (unsigned long long)p.bits_per_ofdm) & 31) : 0;
          const int sym_words = ADAPT ? (intra + (int)p.bits_per_ofdm + 31) >> 5 : N * bps / 32;
          unsigned* wscr = reinterpret_cast<unsigned*>(buf);
#pragma unroll
          for (int j = 0; j < PF; ++j)
            if (t + T * j < sym_words) wscr[t + T * j] = __byte_perm(next_bits[j], 0u, 0x0123);
          if (t == 0) wscr[sym_words] = 0u;
          // one OFDM symbol ahead: this team's next recorded bits into registers, its noise towards L2
          replay_prefetch(ISI ? (halo ? chain_lo : s + 1) : s + s_stride);
          tsync();
#pragma unroll
          for (int j = 0; j < WORDS; ++j) txc[j] = txr[j] = 0u;
          const unsigned smask = (1u << p.half_bits) - 1u;
          const int bit0 = t * bps, tb = T * bps;
#pragma unroll
          for (int m = 0; m < E; ++m) {
            if constexpr (ADAPT) {
              const int hb = __popc((fmask[ADAPT ? m >> 2 : 0] >> (8 * (m & 3))) & 0xffu);   // log2 of the side
              const int bit = intra + (int)__ldg(&p.bit_offsets[t + T * m]), w = bit >> 5;
              const unsigned win = __funnelshift_l(wscr[w + 1], wscr[w], bit & 31);
              const unsigned lab = hb ? win >> (32 - 2 * hb) : 0u;
              txc[m >> 2] |= (lab & ((1u << hb) - 1u)) << (8 * (m & 3) + 1);
              txr[m >> 2] |= (lab >> hb) << (8 * (m & 3) + 1);
              continue;
            }
            const int bit = bit0 + m * tb, w = bit >> 5;
            const unsigned lab = __funnelshift_l(wscr[w + 1], wscr[w], bit & 31) >> (32 - bps);
            if constexpr (PSK) {
              txc[m >> 2] |= lab << (8 * (m & 3));
            } else {
              txc[m >> 2] |= (lab & smask) << (8 * (m & 3) + 1);
              txr[m >> 2] |= (lab >> p.half_bits) << (8 * (m & 3) + 1);
            }
          }
          if constexpr (!PSK) {
#pragma unroll
            for (int j = 0; j < WORDS; ++j) {
              const unsigned fm = ADAPT ? fmask[ADAPT ? j : 0] : p.field_mask;
              txc[j] = (txc[j] ^ (txc[j] >> 1)) & fm;
              txr[j] = (txr[j] ^ (txr[j] >> 1)) & fm;
            }
          }
          tsync();
        } else {
#pragma unroll
          for (int c = 0; c < CALLS; ++c) {
            const uint4 w = philox4x32<DROUNDS>(make_uint4(gs_lo, gs_hi, (0u << 28) | uint32_t(c * T + t), point), key);
            const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (4 * c + j < WORDS) {
                const unsigned fm = ADAPT ? fmask[ADAPT ? 4 * c + j : 0] : p.field_mask;
                if constexpr (PSK) {
                  txc[4 * c + j] = ww[j] & p.field_mask;   // one label per byte (field_mask = M - 1 in every byte)
                  txr[4 * c + j] = 0u;
                } else {
                  txc[4 * c + j] = (ww[j] << 1) & fm;   // bits 0..3 of each byte -> column index
                  txr[4 * c + j] = (ww[j] >> 3) & fm;   // bits 4..7 of each byte -> row index
                }
              }
            }
          }
        }
        // level = 2*index - (s-1) as float via the mantissa of 2^23 + 2*index; re/im swapped so the
        // forward FFT below computes the inverse transform
        const float cen = -(magic + p.slice_top);
        if constexpr (PSK) {
#pragma unroll
          for (int m = 0; m < E; ++m) {
            const float2 pt = s_psk[(txc[m >> 2] >> (8 * (m & 3))) & 0xffu];
            v[m] = make_float2(pt.y, pt.x);
          }
        }
        if constexpr (FLV) {
          // first butterflies of the codelet on the raw label floats f = 2^23 + 2 index (exact sums and differences):
          // (v[n], v[n + E/2]) <- (v[n] + v[n + E/2], v[n] - v[n + E/2]) with v = (-(f_r + cen), f_c + cen)
          // cen2 = -2 (2^23 + s-1): every partial sum below is an even integer below 2^25, hence exact
          const float cen2 = 2.f * cen, ncen2 = -cen2;
          const unsigned k4b = p.k4b;
#pragma unroll
          for (int n = 0; n < E / 2; ++n) {
            const int m2 = n + E / 2;
            const float fca = __uint_as_float(__byte_perm(txc[n >> 2], k4b, 0x7650 + (n & 3)));
            const float fra = __uint_as_float(__byte_perm(txr[n >> 2], k4b, 0x7650 + (n & 3)));
            const float fcc = __uint_as_float(__byte_perm(txc[m2 >> 2], k4b, 0x7650 + (m2 & 3)));
            const float frc = __uint_as_float(__byte_perm(txr[m2 >> 2], k4b, 0x7650 + (m2 & 3)));
            v[n] = make_float2((ncen2 - fra) - frc, (fca + cen2) + fcc);
            v[m2] = make_float2(frc - fra, fca - fcc);
          }
        }
#pragma unroll
        for (int m = 0; m < ((PSK || FLV) ? 0 : E); ++m) {
          const unsigned fc = __byte_perm(txc[m >> 2], 0x4B000000u, 0x7650 + (m & 3));
          const unsigned fr = __byte_perm(txr[m >> 2], 0x4B000000u, 0x7650 + (m & 3));
          if constexpr (ADAPT) {
            // per-subcarrier side s_k and power normalisation: (f - (2^23 + s_k - 1)) is the exact integer level
            const float2 g = __ldg(&level_tab[t + T * m]);
            v[m] = make_float2(-(__uint_as_float(fr) + g.y) * g.x, (__uint_as_float(fc) + g.y) * g.x);
          } else {
            const float li = __uint_as_float(fc) + cen;        // I level:  2*col - (s-1)
            const float lq = -(__uint_as_float(fr) + cen);     // Q level: (s-1) - 2*row
            v[m] = make_float2(lq, li);
          }
        }
      } else if (!SC || phase == 1) {
        // ---- channel + noise, in place in shared memory, 8 samples per iteration
        //      (channel/models.py:52-55 restricted to P >= L-1 -> circular; noise/models.py:19-22)
        float2 prev[8];
        {
          const float2* hrow = buf + ((t + T - 1) % T) * RS + (E - 8);
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float4 q = *reinterpret_cast<const float4*>(hrow + i);
            prev[i] = make_float2(q.x, q.y);
            prev[i + 1] = make_float2(q.z, q.w);
          }
        }
        [[maybe_unused]] float2 new_tail;
        [[maybe_unused]] float2 fold[8];   // zero padding shorter than the channel: what overlap-add folds onto samples 0 .. P-1
        if constexpr (ISI) {
          // sample -j before the body (j = 8 - i): inside the cyclic prefix for j <= P (the wrap-around above), else
          // sample N - (j - P) of the previous OFDM symbol
          if (t == 0) {
            if (p.zero_prefix) {
              // Zero padding shorter than the channel memory (prefix/models.py:71-101 over channel/models.py:52-55): the
              // receiver folds only the P tail samples it keeps, fold[i] = sum_{l > i} h_l x[N - (l - i)] for i < P (lane
              // 0's wrap-around halo IS the symbol's own tail x[N-8 .. N-1]); the rest of the tail leaks into the next
              // symbol, whose samples -j see zeros for j <= P (the padding) and this symbol's body beyond.
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int l = i + 1; l < TAPS; ++l) {
                  const float2 h = p.taps[l], x = prev[8 - (l - i)];
                  acc.x = fmaf(h.x, x.x, fmaf(-h.y, x.y, acc.x));
                  acc.y = fmaf(h.x, x.y, fmaf(h.y, x.x, acc.y));
                }
                fold[i] = i < P ? acc : make_float2(0.f, 0.f);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) prev[i] = (8 - i > P) ? s_tail[i + P] : make_float2(0.f, 0.f);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (8 - i > P) prev[i] = s_tail[i + P];
            }
          }
          if (t < G::TAIL_F2) new_tail = buf[(T - 1) * RS + (E - 8) + t];   // this symbol's tail, before the FIR overwrites it
        }
        tsync();
        if constexpr (ISI) {
          if (t < G::TAIL_F2) s_tail[t] = new_tail;   // read again only in the next symbol's channel phase
        }
        [[maybe_unused]] uint32_t wmin = 0xffffffffu;   // NOISE32: smallest noise word of this lane (refill test)
        [[maybe_unused]] float ps[8];                   // GAUSS: re + im of the halo samples
        if constexpr (GAUSS) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ps[i] = prev[i].x + prev[i].y;
        }
#pragma unroll FIR_UNROLL
        for (int c = 0; c < E / 8; ++c) {
          float2 cur[8], y[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float4 q = *reinterpret_cast<const float4*>(row + 8 * c + i);
            cur[i] = make_float2(q.x, q.y);
            cur[i + 1] = make_float2(q.z, q.w);
          }
          if constexpr (GAUSS) {
            // h x = (k1 - k3) + j (k1 + k2) with k1 = h_re (x_re + x_im), k2 = (h_im - h_re) x_re, k3 = (h_re + h_im) x_im:
            // three sums over the taps, every product a 2-register FFMA with the tap in the constant bank
            float cs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[i] = cur[i].x + cur[i].y;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float k1 = 0.f, k2 = 0.f, k3 = 0.f;
#pragma unroll
              for (int l = 0; l < TAPS; ++l) {
                const float2 x = (i - l >= 0) ? cur[i - l] : prev[8 + i - l];
                const float xs = (i - l >= 0) ? cs[i - l] : ps[8 + i - l];
                const float4 h = FRAMES ? s_hdr.taps3[l] : p.taps3[l];
                k1 = fmaf(h.x, xs, k1);
                if constexpr ((OPT & kOptFirChain) == 0) {
                  k2 = fmaf(h.y, x.x, k2);
                  k3 = fmaf(h.z, x.y, k3);
                }
              }
              if constexpr ((OPT & kOptFirChain) != 0) {
                k2 = k3 = k1;
#pragma unroll
                for (int l = 0; l < TAPS; ++l) {
                  const float2 x = (i - l >= 0) ? cur[i - l] : prev[8 + i - l];
                  const float4 h = FRAMES ? s_hdr.taps3[l] : p.taps3[l];
                  k2 = fmaf(h.y, x.x, k2);
                  k3 = fmaf(-h.z, x.y, k3);
                }
                y[i] = make_float2(k3, k2);
                continue;
              }
              y[i] = make_float2(k1 - k3, k1 + k2);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) ps[i] = cs[i];
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float yr = 0.f, yi = 0.f;
#pragma unroll
              for (int l = 0; l < TAPS; ++l) {
                const float2 x = (i - l >= 0) ? cur[i - l] : prev[8 + i - l];
                const float2 h = FRAMES ? s_hdr.taps[l] : p.taps[l];
                yr = fmaf(h.x, x.x, yr);
                yr = fmaf(-h.y, x.y, yr);
                yi = fmaf(h.x, x.y, yi);
                yi = fmaf(h.y, x.x, yi);
              }
              y[i] = make_float2(yr, yi);
            }
          }
          {  // unconditional (sigma = 0 scales the samples to zero): keeping FIR and noise in ONE basic block
             // lets ptxas interleave the Philox / MUFU chains with the FIR's FFMAs
            if constexpr (!REPLAY && NOISE32) {
              // 8 complex samples from 2 Philox calls: one word per sample (20-bit radius field, 12-bit angle field)
              const uint32_t q2 = 2u * uint32_t((E / 8) * t + c);
              const uint4 wa = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | q2, point), key);
              const uint4 wb = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | (q2 + 1u), point), key);
              const uint32_t w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
              wmin = min(min(wmin, min(w8[0], w8[1])), min(min(w8[2], w8[3]), min(min(w8[4], w8[5]), min(w8[6], w8[7]))));
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // explicit fused add: whether rad * d + y contracts must not depend on the instantiation (the dump-capable
                // kernel also needs the product on its own), the counters of the two are compared for equality
                const float rad = noise20_radius(w8[i], noise_c2, noise_c2m);
                const float2 d = noise20_dir(w8[i]);
                if constexpr (DUMP) {
                  if (active && p.dump_noise)
                    p.dump_noise[s * (unsigned long long)(N + P) + noise_off + E * t + 8 * c + i] = make_float2(__fmul_rn(rad, d.x), __fmul_rn(rad, d.y));
                }
                y[i] = make_float2(__fmaf_rn(rad, d.x, y[i].x), __fmaf_rn(rad, d.y, y[i].y));
              }
            } else if constexpr (!REPLAY) {
              // 8 complex samples from 3 Philox calls: 8 x 32-bit radius words + 8 x 16-bit angle fields
              const uint32_t q3 = 3u * uint32_t((E / 8) * t + c);
              const uint4 wa = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | q3, point), key);
              const uint4 wb = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | (q3 + 1u), point), key);
              const uint4 wc = philox4x32<NROUNDS>(make_uint4(gs_lo, gs_hi, (1u << 28) | (q3 + 2u), point), key);
              const uint32_t rw[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
              const uint32_t aw[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 g = fast_noise(rw[i], aw[i >> 1], (i & 1) ? 0x7632u : 0x7610u, noise_c2);
                if constexpr (DUMP) {
                  if (active && p.dump_noise) p.dump_noise[s * (unsigned long long)(N + P) + noise_off + E * t + 8 * c + i] = g;
                }
                y[i] = cadd(y[i], g);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; i += 2)
            *reinterpret_cast<float4*>(row + 8 * c + i) = make_float4(y[i].x, y[i].y, y[i + 1].x, y[i + 1].y);
#pragma unroll
          for (int i = 0; i < 8; ++i) prev[i] = cur[i];
        }
        if constexpr (ISI) {
          if (p.zero_prefix && t == 0) {   // the folded part of the symbol's own tail (see above)
#pragma unroll
            for (int i = 0; i < 8; ++i) row[i] = cadd(row[i], fold[i]);
          }
        }
        if constexpr (!REPLAY && NOISE32) {
          if (wmin < kRefillBelow)   // probability 2^-20 per sample: out of line
            noise_refill<E, NROUNDS>(row, t, gs_lo, gs_hi, point, key, noise_c2, noise_c2m, (DUMP && active) ? p.dump_noise : nullptr,
                                     s * (unsigned long long)(N + P) + noise_off + E * t);
        }
        if constexpr (!REPLAY) {
          if (p.zero_prefix)   // rare link shape: kept out of line so that the hot instruction stream stays small
            zero_prefix_tail_noise<E, NROUNDS>(row, t, P, gs_lo, gs_hi, point, key, noise_c2,
                                               (DUMP && active) ? p.dump_noise : nullptr, s * (unsigned long long)(N + P) + N);
        }
        tsync();
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = col[W * RS * m];
        if constexpr (REPLAY) {
          // recorded noise, added in the transposed layout: consecutive lanes read consecutive samples (the lines
          // were pulled into L2 one OFDM symbol ago)
          const unsigned long long ni = (active ? s : 0ull) * (unsigned long long)(N + P) + noise_off + t;
          if (p.noise && !p.noise_f64) {
            const float2* nz = reinterpret_cast<const float2*>(p.noise) + ni;
#pragma unroll
            for (int m = 0; m < E; ++m) v[m] = cadd(v[m], __ldg(nz + T * m));
          } else if (p.noise) {
            const double2* nz = reinterpret_cast<const double2*>(p.noise) + ni;
#pragma unroll
            for (int m = 0; m < E; ++m) {
              const double2 g = __ldg(nz + T * m);
              v[m] = cadd(v[m], make_float2((float)g.x, (float)g.y));
            }
          }
          if (p.zero_prefix && p.noise) {
            // overlap-add of the recorded tail noise: stream sample N + n onto sample n = t + T m < P
#pragma unroll
            for (int m = 0; m < E; ++m) {
              if (t + T * m < P) {
                const unsigned long long ti = ni + N + T * m;
                float2 g;
                if (p.noise_f64) {
                  const double2 d = __ldg(reinterpret_cast<const double2*>(p.noise) + ti);
                  g = make_float2((float)d.x, (float)d.y);
                } else {
                  g = __ldg(reinterpret_cast<const float2*>(p.noise) + ti);
                }
                v[m] = cadd(v[m], g);
              }
            }
          }
        }
        if constexpr (FLV) {   // the transform body below starts at the second stage (see the mapper)
#pragma unroll
          for (int n = 0; n < E / 2; ++n) {
            const float2 a = v[n], c = v[n + E / 2];
            v[n] = cadd(a, c);
            v[n + E / 2] = csub(a, c);
          }
        }
        tsync();
        section_sync<SYNC, BLOCK>();
      }

      // ---- forward FFT of N = E*T points: radix-E in registers, row/column exchange, twiddle, radix-E; when the
      //      team is wider than E a second exchange and a radix-W pass follow (Stockham: natural order throughout)
      float2 u[E];
      // element t + T m of the transform: u[oidx(m)]
      auto oidx = [](int m) constexpr { return W > 1 ? m : fft_out_index<E>(m); };
      if constexpr (SC) {
        if (phase == 0) {
          // no transform at the single-carrier transmitter: the levels are the time samples
          tx_epilogue([&](int m) { return make_float2(v[m].y, v[m].x); });
          continue;
        }
      }
      {
      fft_dit_inplace<E, -1, FLV ? 1 : 0>(v);
#pragma unroll
      for (int r = 0; r < E; r += 2) {
        const float2 a = v[fft_out_index<E>(r)], b = v[fft_out_index<E>(r + 1)];
        *reinterpret_cast<float4*>(row + r) = make_float4(a.x, a.y, b.x, b.y);
      }
      // TWR > 0: the pass-2 twiddles travel through a ring of TWR 128-bit buffers that is filled BEFORE the column loads
      // (the table is read-only), so that no twiddle load waits behind the 32 column loads in the shared-memory queue and
      // every later one is requested TWR pairs of legs ahead of its use
      [[maybe_unused]] float4 ring[TWR > 0 ? TWR : 1];
      if constexpr (TWR > 0) {
#pragma unroll
        for (int b = 0; b < TWR && b < E / 2; ++b) ring[b] = lds128<TWV>(s_tw + tcol * RS + 2 * b);
      }
      tsync();
#pragma unroll
      for (int m = 0; m < E; ++m) u[m] = TWV ? lds64v(col + W * RS * m) : col[W * RS * m];
      tsync();
      section_sync<SYNC, BLOCK>();
      if constexpr (FTW) {
        // table row of this lane: {W^(k n), W^(k (n + E/2))} for n < E/2 (k = tcol); the twiddles ride in the first butterflies
#pragma unroll
        for (int n = 0; n < E / 2; ++n) {
          float4 w;
          if constexpr (TWR > 0) {
            w = ring[n % TWR];
          } else {
            w = *reinterpret_cast<const float4*>(s_tw + tcol * RS + 2 * n);
          }
          const float2 a = n == 0 ? u[0] : cmul(u[n], make_float2(w.x, w.y)), c = u[n + E / 2];
          const float2 lo = make_float2(fmaf(-w.w, c.y, fmaf(w.z, c.x, a.x)), fmaf(w.w, c.x, fmaf(w.z, c.y, a.y)));
          u[n] = lo;
          u[n + E / 2] = make_float2(fmaf(2.0f, a.x, -lo.x), fmaf(2.0f, a.y, -lo.y));
          if constexpr (TWR > 0) {
            if (n + TWR < E / 2) ring[n % TWR] = lds128<TWV>(s_tw + tcol * RS + 2 * (n + TWR));
          }
        }
      }
#pragma unroll
      for (int c = 0; c < (FTW ? 0 : E - 1); c += 2) {   // twiddles of legs c + 1 and c + 2 in one 128-bit load (row stride RS: conflict-free)
        float4 w;
        if constexpr (TWR > 0) {
          w = ring[(c / 2) % TWR];
        } else {
          w = *reinterpret_cast<const float4*>(s_tw + tcol * RS + c);
        }
        u[c + 1] = cmul(u[c + 1], make_float2(w.x, w.y));
        if (c + 2 < E) u[c + 2] = cmul(u[c + 2], make_float2(w.z, w.w));
        if constexpr (TWR > 0) {
          if (c / 2 + TWR < E / 2) ring[(c / 2) % TWR] = lds128<TWV>(s_tw + tcol * RS + c + 2 * TWR);
        }
      }
      fft_dit_inplace<E, -1, FTW ? 1 : 0>(u);
      if constexpr (W > 1) {
        // pass-2 output r of butterfly j = t lands at linear index E*E*(t/E) + E*r + (t%E); reload the strided
        // set; radix-W butterflies j = t + T q over the legs u[q + r Q], twiddles W_N^(j r), results in place
        constexpr int Q = E / W;
        float2* blk = buf + (E * trow) * RS + tcol;
#pragma unroll
        for (int r = 0; r < E; ++r) blk[r * RS] = u[fft_out_index<E>(r)];
        tsync();
#pragma unroll
        for (int m = 0; m < E; ++m) u[m] = col[W * RS * m];
        const float2* s_tw3 = s_tw + G::TW2_F2;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float2 w1 = s_tw3[t + T * q];
          // FMA-fused butterflies: a + w c costs 4 FFMA and a - w c = 2 a - (a + w c) two more
          auto fused = [](float2& a, float2& c, float2 w) {
            const float2 lo = make_float2(fmaf(-w.y, c.y, fmaf(w.x, c.x, a.x)), fmaf(w.y, c.x, fmaf(w.x, c.y, a.y)));
            c = make_float2(fmaf(2.0f, a.x, -lo.x), fmaf(2.0f, a.y, -lo.y));
            a = lo;
          };
          if constexpr (W == 2) {
            fused(u[q], u[q + Q], w1);
          } else if constexpr (W == 8) {
            // radix 8 as three radix-2 steps on x_r = u[q + r Q]: X_k = sum_r x_r w^r omega8^(r k).  Step 1 pairs (r, r + 4)
            // with w^4; the even outputs are then a radix-4 transform of the sums with base twiddle w, the odd ones of the
            // differences with base w omega8; steps 2 and 3 are the radix-4 recipe below for both halves.  Results sit in
            // bit-reversed positions and are renamed into natural order.
            constexpr float h = 0.70710678118654752440f;
            const float2 w2 = cmul(w1, w1), w4 = cmul(w2, w2);
            const float2 w8 = make_float2((w1.x + w1.y) * h, (w1.y - w1.x) * h);   // w omega8, omega8 = (1 - j) / sqrt 2
            auto mj = [](float2 z) { return make_float2(z.y, -z.x); };            // -j z
            float2 x[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) x[r] = u[q + r * Q];
#pragma unroll
            for (int r = 0; r < 4; ++r) fused(x[r], x[r + 4], w4);
            fused(x[0], x[2], w2);
            fused(x[1], x[3], w2);
            fused(x[4], x[6], mj(w2));
            fused(x[5], x[7], mj(w2));
            fused(x[0], x[1], w1);
            fused(x[2], x[3], mj(w1));
            fused(x[4], x[5], w8);
            fused(x[6], x[7], mj(w8));
            u[q] = x[0];
            u[q + Q] = x[4];
            u[q + 2 * Q] = x[2];
            u[q + 3 * Q] = x[6];
            u[q + 4 * Q] = x[1];
            u[q + 5 * Q] = x[5];
            u[q + 6 * Q] = x[3];
            u[q + 7 * Q] = x[7];
          } else {
            // radix 4 as two radix-2 steps: with w3 = w1 w2, w1 a1 +- w3 a3 = w1 (a1 +- w2 a3), so the legs need w2 first and
            // w1 (times 1 or -j) in the second step: X0,2 = (a0 + w2 a2) +- w1 (a1 + w2 a3), X1,3 = (a0 - w2 a2) -+ j w1 (a1 - w2 a3)
            const float2 w2 = cmul(w1, w1);
            fused(u[q], u[q + 2 * Q], w2);
            fused(u[q + Q], u[q + 3 * Q], w2);
            fused(u[q], u[q + Q], w1);                                      // X0 -> u[q], X2 -> u[q + Q]
            fused(u[q + 2 * Q], u[q + 3 * Q], make_float2(w1.y, -w1.x));    // X1 -> u[q + 2Q], X3 -> u[q + 3Q]
            const float2 x2 = u[q + Q];
            u[q + Q] = u[q + 2 * Q];
            u[q + 2 * Q] = x2;
          }
        }
      }
      }

      if (phase == 0) {
        // ---- x~[t + T m] = swap(u[brev m])
        tx_epilogue([&](int m) { const float2 o = u[oidx(m)]; return make_float2(o.y, o.x); });
      } else if (SC && phase == 1) {
        // ---- SC-OFDM: equalise in the frequency domain and hand swap(Z) to the inverse transform of phase 2
        float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < E; ++m) sq[m & 3] = fmaf(u[m].x, u[m].x, fmaf(u[m].y, u[m].y, sq[m & 3]));
        float ss = (sq[0] + sq[1]) + (sq[2] + sq[3]);
#pragma unroll
        for (int off = (T < 32 ? T : 32) / 2; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        if constexpr (T > 32) {
          if (lane == 0) s_red[t / 32] = ss;
          tsync();
          ss = 0.f;
#pragma unroll
          for (int i = 0; i < T / 32; ++i) ss += s_red[i];
        }
        const float sigma2 = ss * mmse_c;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const float2 yv = u[oidx(m)];
          const int k = t + T * m;
          const float4 e = G::EQ_SMEM ? s_eq[k] : __ldg(&eq_g[k]);
          const float a = fmaf(yv.x, e.x, yv.y * e.y);    //  Re(Y conj A)
          const float b = fmaf(yv.x, e.y, -yv.y * e.x);   // -Im(Y conj A)
          const float inv = fast_rcp(e.z + sigma2);
          if constexpr (DUMP) {
            if (active && p.dump_y) p.dump_y[s * N + k] = make_float2(yv.x * p.y_scale, yv.y * p.y_scale);
          }
          v[m] = make_float2(-b * inv, a * inv);          // swap(Z~): the forward transform then computes the inverse
        }
      } else {
        // ---- equaliser + slicer + error count (equalization/models.py:22-63, constellation/models.py:19-27,
        //      simulation/models.py:597-606)
        // per-symbol MMSE noise estimate; branch-free (mmse_c = 0 for ZF / none) so that the shuffle latency
        // overlaps the per-subcarrier products below
        float sigma2 = 0.f;
        if constexpr (!SC && (OPT & kOptNoEstimate) == 0) {
          float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int m = 0; m < E; ++m) sq[m & 3] = fmaf(u[m].x, u[m].x, fmaf(u[m].y, u[m].y, sq[m & 3]));
          float ss = (sq[0] + sq[1]) + (sq[2] + sq[3]);
#pragma unroll
          for (int off = (T < 32 ? T : 32) / 2; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
          if constexpr (T > 32) {   // the team spans T / 32 warps
            if (lane == 0) s_red[t / 32] = ss;
            tsync();
            ss = 0.f;
#pragma unroll
            for (int i = 0; i < T / 32; ++i) ss += s_red[i];
          }
          sigma2 = ss * mmse_c;
        }
        unsigned rxc[WORDS], rxr[WORDS];
#pragma unroll
        for (int j = 0; j < WORDS; ++j) rxc[j] = rxr[j] = 0u;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const float2 yv = u[oidx(m)];
          const int k = t + T * m;
          const float4 e = G::EQ_SMEM ? s_eq[k] : __ldg(&eq_g[k]);
          // SC: yv = swap(z~) of time sample k, already equalised and scaled for the slicer
          const float a = SC ? yv.y : fmaf(yv.x, e.x, yv.y * e.y);    //  Re(Y conj A)
          const float b = SC ? -yv.x : fmaf(yv.x, e.y, -yv.y * e.x);  // -Im(Y conj A)
          const float inv = SC ? 1.0f : fast_rcp((OPT & kOptNoEstimate) ? e.z : e.z + sigma2);
          if constexpr (DUMP) {
            if (!SC && active && p.dump_y) p.dump_y[s * N + k] = make_float2(yv.x * p.y_scale, yv.y * p.y_scale);
            // 2 (s_k - 1) / knorm_k with knorm_k^2 = 2 (M_k - 1) / 3, M_k = (top + 1)^2
            const float zu = ADAPT ? 2.f * e.w * rsqrtf(fmaxf((e.w * e.w + 2.f * e.w) * (2.f / 3.f), 1e-30f)) : p.z_unscale;
            if (active && p.dump_z) p.dump_z[s * N + k] = make_float2(a * inv * zu, -b * inv * zu);
          }
          if constexpr (PSK) {
            // nearest point = nearest angle: k = rint(arg(z) M / 2 pi) mod M, label = gray(k)
            const unsigned kh = (unsigned)__float2int_rn(atan2f(-b, a) * p.psk_scale) & ((1u << p.psk_bits) - 1u);
            rxc[m >> 2] |= (kh ^ (kh >> 1)) << (8 * (m & 3));
            continue;
          }
          // sat() clamps to the outermost levels, the 2^23 trick rounds to the nearest level index
          const float top = ADAPT ? e.w : p.slice_top;
          const float tc = fmaf(__saturatef(fmaf(a, inv, 0.5f)), top, magic);
          const float tr = fmaf(__saturatef(fmaf(b, inv, 0.5f)), top, magic);
          // accumulate 2*index into byte (m & 3) of the packed word; the 0x4B000000 parts cancel below
          rxc[m >> 2] += __float_as_uint(tc) << (8 * (m & 3) + 1);
          rxr[m >> 2] += __float_as_uint(tr) << (8 * (m & 3) + 1);
        }
        unsigned be = 0, se = 0;
#pragma unroll
        for (int j = 0; j < WORDS; ++j) {
          // sum over the 4 bytes of (0x4B000000 << (8i+1)) mod 2^32: only i = 0 survives
          constexpr unsigned K = (0x4B000000u << 1);
          if constexpr (PSK) {
            const unsigned d = rxc[j] ^ txc[j];
            be += __popc(d);
            unsigned any = d | (d >> 4);
            any |= any >> 2;
            any |= any >> 1;
            se += __popc(any & 0x01010101u);
            if constexpr (DUMP) {
              if (active) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int k = t + T * (4 * j + i);
                  if (p.dump_tx) p.dump_tx[s * N + k] = (unsigned short)((txc[j] >> (8 * i)) & 0xffu);
                  if (p.dump_rx) p.dump_rx[s * N + k] = (unsigned short)((rxc[j] >> (8 * i)) & 0xffu);
                }
              }
            }
            continue;
          }
          const unsigned dc = ((rxc[j] - K) ^ txc[j]) & 0x1E1E1E1Eu;
          const unsigned dr = ((rxr[j] - K) ^ txr[j]) & 0x1E1E1E1Eu;
          if constexpr (!DUMP && (OPT & kOptJointGray) != 0) {
            const unsigned d = (dc >> 1) + (dr << 3);                 // column field in the low nibble, row field in the high one
            unsigned g = d ^ ((d >> 1) & 0x77777777u);               // inverse Gray code inside every nibble
            g ^= (g >> 2) & 0x33333333u;
            be += __popc(g);
            se += __popc((((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u);   // bit 7 of every non-zero byte
            continue;
          }
          be += __popc(inv_gray_fields(dc) & 0x1E1E1E1Eu) + __popc(inv_gray_fields(dr) & 0x1E1E1E1Eu);
          // a byte of dc | dr is at most 0x1E: adding 0x7F sets its bit 7 exactly when it is non-zero, without carries
          se += __popc(((dc | dr) + 0x7F7F7F7Fu) & 0x80808080u);
          if constexpr (DUMP) {
            if (active) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int m = 4 * j + i, k = t + T * m;
                const unsigned ct = (txc[j] >> (8 * i + 1)) & 15u, rt = (txr[j] >> (8 * i + 1)) & 15u;
                const unsigned cr = ((rxc[j] - K) >> (8 * i + 1)) & 15u, rr = ((rxr[j] - K) >> (8 * i + 1)) & 15u;
                auto ig = [](unsigned x) { x ^= x >> 1; x ^= x >> 2; return x & 15u; };
                const int hb = ADAPT ? __popc((fmask[ADAPT ? j : 0] >> (8 * i)) & 0xffu) : p.half_bits;
                if (p.dump_tx) p.dump_tx[s * N + k] = (unsigned short)((ig(rt) << hb) | ig(ct));
                if (p.dump_rx) p.dump_rx[s * N + k] = (unsigned short)((ig(rr) << hb) | ig(cr));
              }
            }
          }
        }
        if (active) {
          acc_bit_err += be;
          acc_sym_err += se;
          acc_syms += E;
        }
      }
    }
    if constexpr (ISI) {
      if (halo) {   // only the transmitter ran: keep the tail of the symbol before the chain
        if (has_prev && t < G::TAIL_F2) s_tail[t] = buf[(T - 1) * RS + (E - 8) + t];
        tsync();
      }
    }
  }

  if constexpr (FRAMES) {
    unsigned long long* fc = p.frame_counters + (unsigned long long)cur_frame * 10;
    flush_counters(fc, reinterpret_cast<double*>(fc + 8), fc + 9);
    acc_bit_err = acc_sym_err = acc_syms = 0;
    acc_pow = 0.0;
    acc_max = 0.f;
    unit += gridDim.x;
  }
  } while (FRAMES);
  if constexpr (!FRAMES) {
    unsigned long long* cb = p.counters + 10ull * blockIdx.y;
    flush_counters(cb, reinterpret_cast<double*>(cb + 8), cb + 9);
  }
}

}  // namespace ofdm
