// ofdm_link_stream2_kernel: the counters-only Monte-Carlo kernel of the fused mode (Philox bits and noise, no dumps) for
// OFDM links with a guard interval at least as long as the channel memory - one square-QAM order or per-subcarrier
// orders / applied power loading (ADAPT), single links and batches of channel realisations (FRAMES).  Same arithmetic,
// same Philox counters and same results, bit for bit, as ofdm_link_fast_kernel (link_fast.cuh), which keeps the dump /
// recorded-stream / SC-OFDM / short-prefix / PSK instantiations and is the kernel the oracle is compared with; the
// parity tests check that both kernels return identical counters (tests/test_link_edges_gpu.py).
//
// What differs is the shape of the instruction stream.  An SM fetches from a 32 KB instruction cache: a symbol loop
// that is larger is re-fetched from L2 on every pass (cyclic access, LRU) at about half an instruction per clock and
// scheduler (tools/microbench/icache.cu), and the 51 KB loop of link_fast.cuh spends 0.5 stall cycles per instruction
// waiting for instructions (profiles/r2_fast_kernel_history.md).  Here the symbol loop is 28 KB:
//   * ONE copy of the radix-E codelet: the four passes of the two transforms are iterations of a rolled loop; nothing
//     lives in registers across an iteration (every pass loads its input from shared memory and stores its output
//     there), so the loop costs no register moves;
//   * the mapper and the slicer are rolled loops over the lane's label words (4 subcarriers each): the mapper writes the
//     levels and the label words into the lane's own row of shared memory, the slicer reads the spectrum and the label
//     words back from there;
//   * the FIR / noise stage is a rolled loop over 8-sample chunks;
//   * the noise direction (cos, sin) comes from a 2048-entry table in shared memory instead of two MUFU evaluations.
#pragma once
#include "link_stream_kernel.cuh"

namespace ofdm {

// TRIG = false keeps the noise of ofdm_link_fast_kernel (two MUFU evaluations per direction): then the counters of the
// two kernels are identical integers for the same seed, which is how the parity tests pin this kernel.
// MODE bits: 1 direction table; 2 FIR as a rolled loop running BACKWARDS over the chunks (no register carry: the halo of a
// chunk is still unmodified in the row; the first chunk takes it from a side buffer); 4 slicer in two rolled halves (register
// moves); 8 slicer rolled over label words through the lane's row (excludes 4)
template <int E, int T, int BLOCK = 512, int MODE = 0, int TAPS = kFastTaps>
__global__ void __launch_bounds__(BLOCK) ofdm_link_stream2_kernel(const __grid_constant__ FastParams p) {
  static_assert(TAPS >= 1 && TAPS <= kFastTaps, "tap count");
  constexpr bool ADAPT = false, FRAMES = false, TRIG = (MODE & 1) != 0, FIRBACK = (MODE & 2) != 0, HALVES = (MODE & 4) != 0, ROLL = (MODE & 8) != 0;
  static_assert(!(HALVES && ROLL), "one slicer form");
  using G = StreamGeometry<E, T, BLOCK>;
  constexpr int N = G::N, RS = G::RS, WORDS = E / 4, W = G::W;
  constexpr int CALLS = (E + 15) / 16;  // Philox calls for E random bytes
  extern __shared__ float4 smem4[];
  const int lane = threadIdx.x & 31;
  const int t = threadIdx.x % T;
  const int team_in_block = threadIdx.x / T;
  const int tcol = t % E, trow = t / E;
  float2* smem2 = reinterpret_cast<float2*>(smem4);
  float2* buf = smem2 + size_t(team_in_block) * G::TEAM_F2;
  float2* row = buf + t * RS;
  float2* s_tw = smem2 + size_t(G::TEAMS) * G::TEAM_F2;
  float4* s_eq = reinterpret_cast<float4*>(s_tw + G::TW_F2);
  float* s_red = reinterpret_cast<float*>(s_eq + N) + team_in_block * (T / 32);
  float2* col = buf + trow * RS + tcol;   // strided set: col[W * RS * m]
  const char* s_trig = reinterpret_cast<const char*>(smem4) + G::BASE_BYTES;
  uint2* s_stash = reinterpret_cast<uint2*>(reinterpret_cast<char*>(smem4) + G::BASE_BYTES + G::TRIG_BYTES) + threadIdx.x;
  // FIRBACK: the last 8 samples of every lane's row, saved before the lane overwrites them (aliases the stash unless ROLL)
  float4* s_halo = reinterpret_cast<float4*>(reinterpret_cast<char*>(smem4) + G::BASE_BYTES + G::TRIG_BYTES + (ROLL ? G::STASH_BYTES : 0));
  float4* my_halo = s_halo + 4 * threadIdx.x;
  const float4* left_halo = s_halo + 4 * (team_in_block * T + (t + T - 1) % T);
  auto tsync = [&]() { team_sync<T>(team_in_block); };

  for (int i = threadIdx.x; i < G::TW_F2; i += BLOCK) s_tw[i] = __ldg(&p.tw[i]);
  if constexpr (!FRAMES) {
    for (int i = threadIdx.x; i < N; i += BLOCK) s_eq[i] = __ldg(&p.eq_tab[i]);
  }
  if constexpr (TRIG) {   // direction a: angle 2 pi (a + 0.5) / kTrigEntries - pi
    for (int i = threadIdx.x; i < kTrigEntries; i += BLOCK) {
      float sn, cs;
      sincospif((float(i) + 0.5f) * (2.0f / kTrigEntries) - 1.0f, &sn, &cs);
      reinterpret_cast<float2*>(const_cast<char*>(s_trig))[i] = make_float2(cs, sn);
    }
  }
  __syncthreads();

  const PhiloxKey key{(uint32_t)p.seed, (uint32_t)(p.seed >> 32)};
  const int P = p.prefix_len;
  // zero-padded guard interval >= channel memory: same circular convolution, no prefix power, folded tail noise
  const int Pc = p.zero_prefix ? 0 : P;
  const float magic = 8388608.0f;  // 2^23
  unsigned long long s_lo = 0, s_hi = p.sym_count, sym_base = p.sym_begin;
  const uint32_t point = p.point + blockIdx.y;
  float mmse_c = FRAMES ? 0.f : p.point_tab[blockIdx.y].mmse_c;
  unsigned s_stride = gridDim.x * G::TEAMS;
  unsigned s_first = blockIdx.x * G::TEAMS + team_in_block;
  float noise_c2 = FRAMES ? 0.f : -1.3862943611198906f * p.point_tab[blockIdx.y].sigma * p.point_tab[blockIdx.y].sigma;   // -2 sigma^2 ln 2
  float noise_c2m = -32.000003814697266f * noise_c2;
  const float2* level_tab = p.level_tab;
  const unsigned* field_masks = p.field_masks;
  __shared__ FrameHeader s_hdr;
  [[maybe_unused]] unsigned long long unit = blockIdx.x;
  [[maybe_unused]] long long cur_frame = -1;

  unsigned long long acc_bit_err = 0, acc_sym_err = 0, acc_syms = 0;
  double acc_pow = 0.0;
  float acc_max = 0.f;

  // bits per OFDM symbol carried by this lane's E subcarriers
  unsigned lane_bits = E * 2 * p.half_bits;
  auto count_lane_bits = [&]() {
    lane_bits = 0;
#pragma unroll
    for (int j = 0; j < WORDS; ++j) lane_bits += 2 * __popc(__ldg(&field_masks[j * T + t]));
  };
  if constexpr (ADAPT && !FRAMES) count_lane_bits();

  auto flush_counters = [&](unsigned long long* counters, double* power_sum, unsigned long long* power_max_bits) {
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
      atomicAdd(power_sum, pw * double(p.tx_scale2));
      atomicMax(power_max_bits, (unsigned long long)__double_as_longlong(double(mx) * double(p.tx_scale2)));
    }
  };

  do {
  if constexpr (FRAMES) {
    if (unit >= (unsigned long long)p.n_frames * p.chunks_per_frame) break;
    const unsigned long long f = unit / p.chunks_per_frame, c = unit - f * p.chunks_per_frame;
    if ((long long)f != cur_frame) {
      // the whole block moves to frame f: its equaliser table and header into shared memory
      __syncthreads();
      for (int i = threadIdx.x; i < N; i += BLOCK) s_eq[i] = __ldg(&p.eq_tab[f * N + i]);
      if (threadIdx.x < (int)(sizeof(FrameHeader) / sizeof(float)))
        reinterpret_cast<float*>(&s_hdr)[threadIdx.x] = __ldg(reinterpret_cast<const float*>(&p.frame_hdr[f]) + threadIdx.x);
      __syncthreads();
      field_masks = p.field_masks + f * (N / 4);
      count_lane_bits();
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
  const unsigned iters = s_hi > s_lo ? (unsigned)((s_hi - s_lo + s_stride - 1) / s_stride) : 0u;
  for (unsigned it = 0; it < iters; ++it) {
    const unsigned long long s = s_lo + (unsigned long long)it * s_stride + s_first;
    const bool active = s < s_hi;
    const unsigned long long gs = sym_base + (active ? s : s_lo);
    const uint32_t gs_lo = (uint32_t)gs, gs_hi = (uint32_t)(gs >> 32);

    // ---- the four radix-E passes of the symbol: 0-1 inverse transform at the transmitter (a forward transform of
    //      re/im-swapped data), 2-3 forward transform at the receiver.  Nothing is live in registers across a pass.
    [[maybe_unused]] unsigned txc[WORDS], txr[WORDS];
#pragma unroll 1
    for (int step = 0; step < 4; ++step) {
      float2 v[E];
      if (step == 0) {
        // ---- bits -> QAM levels (constellation/models.py:180-249), one random byte per subcarrier: low nibble ->
        //      column (in-phase) index, high nibble -> row (quadrature) index; word j of the lane = subcarriers
        //      k = t + T (4 j + i).  Label words to the stash, levels (re/im swapped) to the lane's own row.
#pragma unroll(CALLS > 2 ? 1 : CALLS)
        for (int c = 0; c < CALLS; ++c) {
          const uint4 w = philox4x32<10>(make_uint4(gs_lo, gs_hi, (0u << 28) | uint32_t(c * T + t), point), key);
          const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * c + jj;
            if (jj < WORDS) {   // (E = 8: two words per lane)
              const unsigned fm = ADAPT ? __ldg(&field_masks[j * T + t]) : p.field_mask;
              const unsigned wc = (ww[jj] << 1) & fm;   // bits 0..3 of each byte -> column index
              const unsigned wr = (ww[jj] >> 3) & fm;   // bits 4..7 of each byte -> row index
              if constexpr (ROLL) s_stash[j * BLOCK] = make_uint2(wc, wr);
              else { txc[j] = wc; txr[j] = wr; }
              // level = 2*index - (s-1) as float via the mantissa of 2^23 + 2*index
              const float cen = -(magic + p.slice_top);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const unsigned fc = __byte_perm(wc, 0x4B000000u, 0x7650 + i);
                const unsigned fr = __byte_perm(wr, 0x4B000000u, 0x7650 + i);
                const float li = __uint_as_float(fc) + cen;        // I level:  2*col - (s-1)
                const float lq = -(__uint_as_float(fr) + cen);     // Q level: (s-1) - 2*row
                v[4 * j + i] = make_float2(lq, li);
              }
            }
          }
        }
      } else {
        if (step == 2) {
          // ---- channel + noise, in place in shared memory, 8 samples per iteration
          //      (channel/models.py:52-55 with P >= L-1 -> circular; noise/models.py:19-22)
          uint32_t wmin = 0xffffffffu;   // smallest noise word of this lane (refill test)
          // one 8-sample chunk: cur = the chunk, prev = the 8 samples before it
          auto fir_chunk = [&](int c, const float2 (&cur)[8], const float2 (&prev)[8], const float (&cs)[8], const float (&ps)[8]) {
            float2 y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float k1 = 0.f, k2 = 0.f, k3 = 0.f;
#pragma unroll
              for (int l = 0; l < TAPS; ++l) {
                const float2 x = (i - l >= 0) ? cur[i - l] : prev[8 + i - l];
                const float xs = (i - l >= 0) ? cs[i - l] : ps[8 + i - l];
                const float4 h = p.taps3[l];
                k1 = fmaf(h.x, xs, k1);
                k2 = fmaf(h.y, x.x, k2);
                k3 = fmaf(h.z, x.y, k3);
              }
              y[i] = make_float2(k1 - k3, k1 + k2);
            }
            {  // 8 complex samples from 2 Philox calls: one word per sample
              const uint32_t q2 = 2u * uint32_t((E / 8) * t + c);
              const uint4 wa = philox4x32<10>(make_uint4(gs_lo, gs_hi, (1u << 28) | q2, point), key);
              const uint4 wb = philox4x32<10>(make_uint4(gs_lo, gs_hi, (1u << 28) | (q2 + 1u), point), key);
              const uint32_t w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
              wmin = min(min(wmin, min(w8[0], w8[1])), min(min(w8[2], w8[3]), min(min(w8[4], w8[5]), min(w8[6], w8[7]))));
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if constexpr (TRIG) {
                  const float rad = stream_radius(w8[i], noise_c2, noise_c2m);
                  const float2 d = stream_dir(w8[i], s_trig);
                  y[i] = make_float2(fmaf(rad, d.x, y[i].x), fmaf(rad, d.y, y[i].y));
                } else {
                  y[i] = cadd(y[i], fast_noise20(w8[i], noise_c2, noise_c2m));
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2)
              *reinterpret_cast<float4*>(row + 8 * c + i) = make_float4(y[i].x, y[i].y, y[i + 1].x, y[i + 1].y);
          };
          auto load8 = [&](const float2* src, float2 (&dst)[8], float (&sum)[8]) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const float4 q = *reinterpret_cast<const float4*>(src + i);
              dst[i] = make_float2(q.x, q.y);
              dst[i + 1] = make_float2(q.z, q.w);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sum[i] = dst[i].x + dst[i].y;
          };
          if constexpr (FIRBACK) {
            // backwards over the chunks: the 8 samples before a chunk are still the transmitter's; the chunk before the
            // first one is the left neighbour's tail, which that lane parks in the side buffer before overwriting it
#pragma unroll 1
            for (int c = E / 8 - 1; c >= 0; --c) {
              float2 cur[8], prev[8];
              float cs[8], ps[8];
              load8(row + 8 * c, cur, cs);
              if (c == E / 8 - 1) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) my_halo[i / 2] = make_float4(cur[i].x, cur[i].y, cur[i + 1].x, cur[i + 1].y);
                tsync();
              }
              load8(c == 0 ? reinterpret_cast<const float2*>(left_halo) : row + 8 * (c - 1), prev, ps);
              fir_chunk(c, cur, prev, cs, ps);
            }
          } else {
            float2 prev[8];
            float ps[8];
            load8(buf + ((t + T - 1) % T) * RS + (E - 8), prev, ps);
            tsync();
#pragma unroll 2
            for (int c = 0; c < E / 8; ++c) {
              float2 cur[8];
              float cs[8];
              load8(row + 8 * c, cur, cs);
              fir_chunk(c, cur, prev, cs, ps);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                prev[i] = cur[i];
                ps[i] = cs[i];
              }
            }
          }
          if constexpr (TRIG) {
            if (wmin < kStreamRefillBelow)   // probability 2^-18 per sample: out of line
              stream_refill<E>(row, t, gs_lo, gs_hi, point, key, noise_c2, noise_c2m, s_trig);
          } else {
            if (wmin < kRefillBelow)
              noise_refill<E, 10>(row, t, gs_lo, gs_hi, point, key, noise_c2, noise_c2m, nullptr, 0ull);
          }
          if (p.zero_prefix)   // rare link shape, out of line
            zero_prefix_tail_noise<E, 10>(row, t, P, gs_lo, gs_hi, point, key, noise_c2, nullptr, 0ull);
          tsync();
        }
        // ---- the strided set of the team's buffer: element t + T m
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = col[W * RS * m];
        tsync();
        if (step & 1) {
#pragma unroll
          for (int c = 0; c < E - 1; c += 2) {   // twiddles of legs c + 1 and c + 2 in one 128-bit load (row stride RS: conflict-free)
            const float4 w = *reinterpret_cast<const float4*>(s_tw + tcol * RS + c);
            v[c + 1] = cmul(v[c + 1], make_float2(w.x, w.y));
            if (c + 2 < E) v[c + 2] = cmul(v[c + 2], make_float2(w.z, w.w));
          }
        }
      }

      fft_dit_inplace<E, -1>(v);

      if (!(step & 1)) {
#pragma unroll
        for (int r = 0; r < E; r += 2) {
          const float2 a = v[fft_out_index<E>(r)], b = v[fft_out_index<E>(r + 1)];
          *reinterpret_cast<float4*>(row + r) = make_float4(a.x, a.y, b.x, b.y);
        }
        tsync();
        continue;
      }
      // element t + T m of the finished transform: v[oidx(m)]
      auto oidx = [](int m) constexpr { return W > 1 ? m : fft_out_index<E>(m); };
      if constexpr (W > 1) {
        // pass-2 output r of butterfly j = t lands at linear index E*E*(t/E) + E*r + (t%E); reload the strided
        // set; radix-W butterflies j = t + T q over the legs v[q + r Q], twiddles W_N^(j r), results in place
        constexpr int Q = E / W;
        float2* blk = buf + (E * trow) * RS + tcol;
#pragma unroll
        for (int r = 0; r < E; ++r) blk[r * RS] = v[fft_out_index<E>(r)];
        tsync();
#pragma unroll
        for (int m = 0; m < E; ++m) v[m] = col[W * RS * m];
        const float2* s_tw3 = s_tw + G::TW2_F2;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float2 w1 = s_tw3[t + T * q];
          if constexpr (W == 2) {
            const float2 a = v[q], b = cmul(v[q + Q], w1);
            v[q] = cadd(a, b);
            v[q + Q] = csub(a, b);
          } else {
            const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            const float2 a0 = v[q], a1 = cmul(v[q + Q], w1), a2 = cmul(v[q + 2 * Q], w2), a3 = cmul(v[q + 3 * Q], w3);
            const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
            const float2 jd = make_float2(d13.y, -d13.x);   // -j (a1 - a3)
            v[q] = cadd(s02, s13);
            v[q + Q] = cadd(d02, jd);
            v[q + 2 * Q] = csub(s02, s13);
            v[q + 3 * Q] = csub(d02, jd);
          }
        }
      }

      if (step == 1) {
        // ---- transmitter epilogue: time samples x~[t + T m] = swap(v[oidx(m)]) to the buffer for the FIR, PAPR
        //      statistics (prefix/models.py:34-44, simulation/models.py:519-524)
        float ssum[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned umax[4] = {0u, 0u, 0u, 0u};
        float pw_even = 0.f;
#pragma unroll
        for (int m = 0; m < E; ++m) {
          const float2 o = v[oidx(m)];
          const float pw = fmaf(o.x, o.x, o.y * o.y);
          // the cyclic prefix repeats the last P samples; for P <= T: n = t + T m >= N - P  <=>  m == E-1, t >= T - Pc
          ssum[m & 3] += (m == E - 1 && t >= T - Pc) ? 2.f * pw : pw;
          // two samples per instruction: non-negative floats order like their bit patterns
          if (m & 1) umax[(m >> 1) & 3] = __vimax3_u32(umax[(m >> 1) & 3], __float_as_uint(pw_even), __float_as_uint(pw));
          else pw_even = pw;
          col[W * RS * m] = make_float2(o.y, o.x);
        }
        if (Pc > T) {
          // long prefix (more than one row of samples): the rows above the last one that it also repeats
          const int pm = (N - Pc) / T, pt = (N - Pc) % T;
#pragma unroll 1
          for (int m = pm; m < E - 1; ++m) {
            const float2 o = col[W * RS * m];
            if (m > pm || t >= pt) ssum[0] += fmaf(o.x, o.x, o.y * o.y);
          }
        }
        if (active) {
          acc_pow += double((ssum[0] + ssum[1]) + (ssum[2] + ssum[3]));
          acc_max = fmaxf(acc_max, __uint_as_float(max(max(umax[0], umax[1]), max(umax[2], umax[3]))));
        }
        tsync();
        continue;
      }

      // ---- step 3: equaliser + slicer + error count (equalization/models.py:22-63, constellation/models.py:19-27,
      //      simulation/models.py:597-606).  Per-symbol MMSE noise estimate (mmse_c = 0 for ZF / none); then the
      //      spectrum to the lane's own row and a rolled loop over the label words.
      float sq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int m = 0; m < E; ++m) sq[m & 3] = fmaf(v[m].x, v[m].x, fmaf(v[m].y, v[m].y, sq[m & 3]));
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
      const float sigma2 = ss * mmse_c;
      unsigned be = 0, se = 0;
      // one label word = 4 subcarriers m = 4 j + i of this lane
      auto slice_word = [&](int j, const float2 (&y)[4], unsigned wc, unsigned wr) {
        unsigned rc = 0u, rr = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 e = s_eq[t + T * (4 * j + i)];
          const float a = fmaf(y[i].x, e.x, y[i].y * e.y);    //  Re(Y conj A)
          const float b = fmaf(y[i].x, e.y, -y[i].y * e.x);   // -Im(Y conj A)
          const float inv = fast_rcp(e.z + sigma2);
          const float top = p.slice_top;
          const float tc = fmaf(__saturatef(fmaf(a, inv, 0.5f)), top, magic);
          const float tr = fmaf(__saturatef(fmaf(b, inv, 0.5f)), top, magic);
          rc += __float_as_uint(tc) << (8 * i + 1);
          rr += __float_as_uint(tr) << (8 * i + 1);
        }
        constexpr unsigned K = (0x4B000000u << 1);
        const unsigned dc = ((rc - K) ^ wc) & 0x1E1E1E1Eu;
        const unsigned dr = ((rr - K) ^ wr) & 0x1E1E1E1Eu;
        be += __popc(inv_gray_fields(dc) & 0x1E1E1E1Eu) + __popc(inv_gray_fields(dr) & 0x1E1E1E1Eu);
        se += __popc(((dc | dr) + 0x7F7F7F7Fu) & 0x80808080u);
      };
      if constexpr (ROLL) {
        if constexpr (W > 1 && T <= 32) tsync();   // every lane has its strided set of the last pass: the rows may be rewritten
#pragma unroll
        for (int m = 0; m < E; m += 2) {
          const float2 a = v[oidx(m)], b = v[oidx(m + 1)];
          *reinterpret_cast<float4*>(row + m) = make_float4(a.x, a.y, b.x, b.y);
        }
#pragma unroll(WORDS >= 4 ? 2 : 1)
        for (int j = 0; j < WORDS; ++j) {
          const float4 q0 = *reinterpret_cast<const float4*>(row + 4 * j), q1 = *reinterpret_cast<const float4*>(row + 4 * j + 2);
          const uint2 wtx = s_stash[j * BLOCK];
          const float2 y[4] = {make_float2(q0.x, q0.y), make_float2(q0.z, q0.w), make_float2(q1.x, q1.y), make_float2(q1.z, q1.w)};
          slice_word(j, y, wtx.x, wtx.y);
        }
      } else if constexpr (HALVES) {
        // two rolled halves: after the first, the second half of the spectrum and of the label words moves into the
        // registers of the first (E / 2 complex + WORDS moves)
        constexpr int EP = E / 2, WP = WORDS / 2;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int jj = 0; jj < WP; ++jj) {
            const float2 y[4] = {v[oidx(4 * jj)], v[oidx(4 * jj + 1)], v[oidx(4 * jj + 2)], v[oidx(4 * jj + 3)]};
            // s_eq index: the subcarriers of half h
            unsigned rc = 0u, rr = 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 e = s_eq[t + T * (4 * jj + i) + T * EP * h];
              const float a = fmaf(y[i].x, e.x, y[i].y * e.y);
              const float b = fmaf(y[i].x, e.y, -y[i].y * e.x);
              const float inv = fast_rcp(e.z + sigma2);
              const float top = p.slice_top;
              const float tc = fmaf(__saturatef(fmaf(a, inv, 0.5f)), top, magic);
              const float tr = fmaf(__saturatef(fmaf(b, inv, 0.5f)), top, magic);
              rc += __float_as_uint(tc) << (8 * i + 1);
              rr += __float_as_uint(tr) << (8 * i + 1);
            }
            constexpr unsigned K = (0x4B000000u << 1);
            const unsigned dc = ((rc - K) ^ txc[jj]) & 0x1E1E1E1Eu;
            const unsigned dr = ((rr - K) ^ txr[jj]) & 0x1E1E1E1Eu;
            be += __popc(inv_gray_fields(dc) & 0x1E1E1E1Eu) + __popc(inv_gray_fields(dr) & 0x1E1E1E1Eu);
            se += __popc(((dc | dr) + 0x7F7F7F7Fu) & 0x80808080u);
          }
          if (h == 0) {
#pragma unroll
            for (int i = 0; i < EP; ++i) v[oidx(i)] = v[oidx(i + EP)];
#pragma unroll
            for (int jj = 0; jj < WP; ++jj) {
              txc[jj] = txc[jj + WP];
              txr[jj] = txr[jj + WP];
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < WORDS; ++j) {
          const float2 y[4] = {v[oidx(4 * j)], v[oidx(4 * j + 1)], v[oidx(4 * j + 2)], v[oidx(4 * j + 3)]};
          slice_word(j, y, txc[j], txr[j]);
        }
      }
      if (active) {
        acc_bit_err += be;
        acc_sym_err += se;
        acc_syms += E;
      }
      // the next writer of this row is the lane itself (the mapper of the next symbol): program order
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
