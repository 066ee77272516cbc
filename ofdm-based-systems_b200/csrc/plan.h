// Shared between the C-ABI translation unit (ofdm_b200.cu) and the per-size kernel instantiation
// units (link_inst.cu, compiled once per supported N so the build parallelises).
#pragma once
#include "../../include/ofdm_b200.h"

#include <cstdint>
#include <cuda_runtime.h>

#include "link_params.h"

namespace ofdm {

constexpr int elements_per_thread(int n) {
  return n <= 8 ? 8 : n == 16 ? 4 : n <= 64 ? 8 : n <= 256 ? 16 : 32;
}

struct CounterBlock {  // device-resident accumulator, 80 bytes
  unsigned long long cnt[CNT_WORDS];
  double power_sum;
  unsigned long long power_max_bits;
};

int fail(int code, const char* fmt, ...);
void count_launch();

int blocks_per_sm_cached(const void* kernel, int block, size_t smem, int* out);

#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) return ::ofdm::fail(OFDM_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// Resident blocks per SM of `kernel` on the CURRENT device, with the opt-in for more than 48 KB of dynamic shared
// memory done on first use.  Both are per kernel and per device (the attribute lives in the device's primary context),
// so the cache is keyed by (kernel, device ordinal); any host thread.
template <class Kern>
int blocks_per_sm(Kern kern, int block, size_t smem, int* out) {
  return blocks_per_sm_cached(reinterpret_cast<const void*>(kern), block, smem, out);
}

}  // namespace ofdm

struct ofdm_link {
  ofdm_link_desc d;
  int E = 0, T = 0, block = 0, teams = 0;
  int bits_per_ofdm = 0;
  int isi = 0;
  int rx_gain = 0;              // 1: a receiver gain per subcarrier was given (applied power loading)
  double mean_h2 = 0.0;
  float2 taps[ofdm::kMaxTaps];
  float4* d_sc = nullptr;
  float4* d_eq = nullptr;
  float2* d_tw = nullptr;
  ofdm::CounterBlock* d_cnt = nullptr;
  ofdm::CounterBlock* d_sweep = nullptr;  // one counter block per SNR point of the last sweep launch (own allocation)
  int sweep_cap = 0, sweep_points = 0;
  unsigned char* arena = nullptr;  // single device allocation behind all the pointers above
  int device = 0, sms = 0, occ = 1;
  size_t smem = 0;
  size_t table_bytes = 0;
  // fast path (link_fast.cuh): eligible link shapes only
  int fast = 0;                 // fast kernel: 0 no, 1 one QAM order on all subcarriers, 2 per-subcarrier orders / loading, 3 PSK
  int fixed_order = 0;          // the single QAM order when fast
  float4* d_eq_fast = nullptr;  // decision-domain equaliser table
  float2* d_level = nullptr;    // fast = 2: {1/knorm_k, -(2^23 + s_k)}
  unsigned* d_mask = nullptr;   // fast = 2: packed field masks
  float2* d_psk = nullptr;      // fast = 3: PSK point table [256]
  unsigned short* d_bitoff = nullptr;   // fast = 2: bit offset of every subcarrier inside an OFDM symbol
  float2* d_tw_fast = nullptr;  // pass-2 twiddles [(r-1)*E + k], then the pass-3 base twiddles exp(-2 pi i j / N)
  float2 taps_fast[8];
  double knorm = 1.0;
  // post-equaliser stage (ofdm_link_set_post): links with one run on the general kernel
  int fast_shape = 0;           // what `fast` was before a post stage switched the link to the general kernel
  float* d_post = nullptr;      // [N] sqrt(noise_profile / 2), own allocation
  const void* post_recorded = nullptr;   // recorded noise matrix for the replay entry points (caller's memory)
  double z_scale = 1.0;
  int z_power = 0;
};

namespace ofdm {
// defined in link_inst.cu, one explicit instantiation per supported size
template <int N> int configure_kernel(ofdm_link* L);
template <int N> int launch_kernel(const ofdm_link* L, const LinkParams& p, cudaStream_t stream);
}  // namespace ofdm

namespace ofdm {
struct FastParams;
bool fast_supports_n(int n);
bool fast_supports_combo(int n, bool adapt, bool sc, bool isi, bool psk);   // link_fast.cu
int fast_samples_per_lane(int n);
}  // namespace ofdm
#include <vector>
namespace ofdm {
std::vector<float2> build_fast_twiddles(int n);
int launch_fast(const ofdm_link* L, const FastParams& p, bool dump, bool replay, bool adapt, bool sc, bool isi, bool psk, cudaStream_t stream);
}  // namespace ofdm

#define OFDM_FOR_EACH_N(X) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)
