// Kernel parameter block of ofdm_link_kernel (plain data, usable from host-only translation units).
#pragma once
#include <cuda_runtime.h>

namespace ofdm {

constexpr int kMaxTaps = 32;

enum : int { PREFIX_NONE = 0, PREFIX_CYCLIC = 1, PREFIX_ZERO = 2 };
enum : int { MOD_OFDM = 0, MOD_SC = 1 };
enum : int { EQ_NONE = 0, EQ_ZF = 1, EQ_MMSE = 2 };
enum : int { SCHEME_QAM = 0, SCHEME_PSK = 1 };
enum : int { SRC_NONE = 0, SRC_PHILOX = 1, SRC_REPLAY_F32 = 2, SRC_REPLAY_F64 = 3 };
enum : int { CNT_BIT_ERRORS = 0, CNT_BITS = 1, CNT_SYM_ERRORS = 2, CNT_SYMBOLS = 3, CNT_OFDM_SYMBOLS = 4,
              CNT_Z_VALUES = 5 /* equalised values whose power was summed */, CNT_Z_POWER = 6 /* sum |z|^2, a double */, CNT_WORDS = 8 };

struct LinkParams {
  // ---- link shape
  int prefix_len, prefix_type, modulator, equalizer, scheme, n_taps;
  int bits_src;   // SRC_PHILOX | SRC_REPLAY_F32 (byte stream, MSB first)
  int noise_src;  // SRC_NONE | SRC_PHILOX | SRC_REPLAY_F32 | SRC_REPLAY_F64
  int isi;        // 1: the FIR reaches into the previous OFDM symbol -> chained processing with a halo
  int limit_bits; // 1: bit positions >= compare_limit are not compared (ragged tail of a replayed stream)
  int rx_gain;    // 1: the equalised subcarrier k is multiplied by sc_tab[k].w before the demapper
  float2 taps[kMaxTaps];  // unit-energy channel taps (channel/models.py:14-16)
  const float4* sc_tab;   // per subcarrier {amp, slicer scale k, bits(bps | bit_offset << 8), receiver gain}
  const float4* eq_tab;   // per subcarrier ZF {Re 1/H, Im 1/H, -, -} | MMSE {Re H, Im H, |H|^2, -}
  const float2* tw;       // inter-pass twiddles (forward sign), see LinkPlan::build_twiddles
  float sigma;            // fused mode: per-component noise standard deviation
  float mmse_c;           // 1 / (N * snr_lin * mean|H|^2)          (equalization/models.py:43-49)
  unsigned long long seed;
  unsigned int point;     // SNR-point index, part of the Philox counter
  unsigned int bits_per_ofdm;
  unsigned long long sym_begin, sym_count;  // global index of this launch's first OFDM symbol, count
  unsigned long long sym_lo;                // first symbol (relative to sym_begin and to the replay buffers) this launch processes
  // ---- replay inputs (indexed by symbol - sym_begin)
  const unsigned char* bits;
  unsigned long long bits_len;
  const void* noise;  // complex64 or complex128 over the serial stream, (N+P) per symbol
  unsigned long long compare_limit;
  // ---- post-equaliser stage (examples/waterfilling_noise_bump_experiment.py:163-183)
  const float* post_sigma;  // [N] per-component standard deviation of the coloured noise for unit noise power, or NULL
  float post_scale;         // sqrt(noise power) = 10^(-snr_db/20)
  int post_src;             // SRC_PHILOX (stream tag 3) | SRC_REPLAY_F64 (recorded matrix [sym][N] complex128)
  const void* post_noise;
  float z_scale;            // multiplies every equalised value before the demapper (1/sqrt(avg_power), :178-181)
  int z_power;              // 1: accumulate sum |z|^2 (before z_scale) and the count into the counter block
  // ---- outputs
  unsigned long long* counters;  // [CNT_WORDS]
  double* tx_power_sum;          // sum |tx|^2 over every tx sample (prefix included)
  unsigned long long* tx_power_max_bits;  // max |tx|^2 as the bit pattern of a non-negative double
  float2* dump_y;                // [sym][N] pre-equaliser FFT output (ortho scaled)
  float2* dump_z;                // [sym][N] what the demapper sees (received_symbols)
  unsigned short* dump_rx;       // [sym][N] decided labels
  unsigned short* dump_tx;       // [sym][N] transmitted labels
  float2* dump_noise;            // [sym][N+P] the noise that was added (fused mode -> replayable)
};

}  // namespace ofdm
