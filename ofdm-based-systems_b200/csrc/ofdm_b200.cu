// C ABI of libofdm_b200.so (see include/ofdm_b200.h): host-side plan building in fp64, kernel
// dispatch, host<->device staging for the host-buffer entry points.
#include "../../include/ofdm_b200.h"

#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cstdlib>

#include "link_fast.cuh"
#include "plan.h"

using namespace ofdm;

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

// Link arenas are a few tens of KB and links are created per channel realisation / per sweep: recycle them instead
// of paying cudaMalloc + cudaFree (which synchronises the device) every time.  Small bounded cache, any thread.
struct ArenaCache {
  struct Entry { unsigned char* ptr; size_t bytes; int device; };
  std::mutex mu;
  std::vector<Entry> free_list;
  unsigned char* acquire(size_t bytes, int device) {
    {
      std::lock_guard<std::mutex> lock(mu);
      for (size_t i = 0; i < free_list.size(); ++i)
        if (free_list[i].device == device && free_list[i].bytes >= bytes && free_list[i].bytes <= 2 * bytes + 4096) {
          unsigned char* p = free_list[i].ptr;
          free_list.erase(free_list.begin() + i);
          return p;
        }
    }
    unsigned char* p = nullptr;
    return cudaMalloc(&p, bytes) == cudaSuccess ? p : nullptr;
  }
  void release(unsigned char* ptr, size_t bytes, int device) {
    if (!ptr) return;
    {
      std::lock_guard<std::mutex> lock(mu);
      if (free_list.size() < 16) {
        free_list.push_back({ptr, bytes, device});
        return;
      }
    }
    cudaFree(ptr);
  }
  ~ArenaCache() { for (auto& e : free_list) cudaFree(e.ptr); }
};
ArenaCache g_arenas;

// Pinned host staging for the table upload of ofdm_link_create: one buffer per host thread, grown on demand.
struct PinnedStage {
  unsigned char* ptr = nullptr;
  size_t bytes = 0;
  unsigned char* reserve(size_t need) {
    if (bytes >= need) return ptr;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    bytes = 0;
    const size_t want = need < (1u << 20) ? (1u << 20) : need;
    if (cudaHostAlloc(reinterpret_cast<void**>(&ptr), want, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      ptr = nullptr;
      return nullptr;
    }
    bytes = want;
    return ptr;
  }
  ~PinnedStage() { if (ptr) cudaFreeHost(ptr); }
};
thread_local PinnedStage g_stage;

}  // namespace

namespace ofdm {
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int blocks_per_sm_cached(const void* kernel, int block, size_t smem, int* out) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> cache;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({kernel, dev});
  if (it == cache.end()) {
    int occ = 0;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, smem));
    it = cache.emplace(std::make_pair(kernel, dev), occ > 0 ? occ : 1).first;
  }
  *out = it->second;
  return OFDM_OK;
}
}  // namespace ofdm

namespace {

bool uniform_orders(const int32_t* orders, int n) {
  for (int k = 1; k < n; ++k)
    if (orders[k] != orders[0]) return false;
  return orders[0] >= 4;
}

int configure(ofdm_link* L) {
  switch (L->d.n_subcarriers) {
#define X(n) case n: return configure_kernel<n>(L);
    OFDM_FOR_EACH_N(X)
#undef X
    default: return fail(OFDM_EUNSUPPORTED, "n_subcarriers=%d: only powers of two in 8..8192 have a CUDA plan", L->d.n_subcarriers);
  }
}

int launch(const ofdm_link* L, const LinkParams& p, cudaStream_t stream) {
  switch (L->d.n_subcarriers) {
#define X(n) case n: return launch_kernel<n>(L, p, stream);
    OFDM_FOR_EACH_N(X)
#undef X
    default: return fail(OFDM_EUNSUPPORTED, "n_subcarriers=%d unsupported", L->d.n_subcarriers);
  }
}

// Inter-pass twiddles of the Stockham plan in link_kernel.cuh (forward sign):
//   pass 2 (radix R2, NS = R1):     tw[(r-1)*(N/R2) + j]       = exp(-2 pi i (j mod NS) r / (NS R2))
//   pass 3 (radix R3, NS = R1*R2):  tw[TW2 + (r-1)*(N/R3) + j] likewise
std::vector<float2> compute_twiddles(int N, int E) {
  const int R1 = E, rem = N / R1, R2 = rem < E ? rem : E, R3 = rem / R2;
  std::vector<float2> tw;
  auto add_pass = [&](int R, int NS) {
    if (R <= 1) return;
    const int NB = N / R;
    const size_t base = tw.size();
    tw.resize(base + size_t(R - 1) * NB);
    for (int r = 1; r < R; ++r)
      for (int j = 0; j < NB; ++j) {
        const int k = j % NS;
        const double ang = -2.0 * M_PI * double((long long)k * r % (long long)(NS * R)) / double(NS * R);
        tw[base + size_t(r - 1) * NB + j] = make_float2((float)std::cos(ang), (float)std::sin(ang));
      }
  };
  add_pass(R2, R1);
  add_pass(R3, R1 * R2);
  if (tw.empty()) tw.push_back(make_float2(1.f, 0.f));
  return tw;
}

// The twiddle tables depend on the transform size only: computed once per size and process (a link is created per
// channel realisation / per sweep, and ~2000 sincos per creation were a third of its host cost).
template <class Build>
const std::vector<float2>& cached_twiddles(int key, Build&& build) {
  static std::mutex mu;
  static std::vector<std::pair<int, std::vector<float2>>> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (auto& e : cache)
    if (e.first == key) return e.second;
  cache.reserve(64);            // references stay valid: at most 2 tables per supported size
  cache.emplace_back(key, build());
  return cache.back().second;
}

void fill_params(const ofdm_link* L, LinkParams& p, double snr_db) {
  std::memset(&p, 0, sizeof(p));
  p.prefix_len = L->d.prefix_len;
  p.prefix_type = L->d.prefix_type;
  p.modulator = L->d.modulator;
  p.equalizer = L->d.equalizer;
  p.scheme = L->d.scheme;
  p.n_taps = L->d.n_taps;
  p.isi = L->isi;
  p.rx_gain = L->rx_gain;
  std::memcpy(p.taps, L->taps, sizeof(p.taps));
  p.sc_tab = L->d_sc;
  p.eq_tab = L->d_eq;
  p.tw = L->d_tw;
  const double snr_lin = std::pow(10.0, snr_db / 10.0);
  // equalization/models.py:43-49: sigma2 = mean|Y|^2 / snr_lin / mean|H|^2, inf if the gain is zero
  p.mmse_c = L->mean_h2 == 0.0 ? INFINITY : (float)(1.0 / (double(L->d.n_subcarriers) * snr_lin * L->mean_h2));
  p.bits_per_ofdm = (unsigned)L->bits_per_ofdm;
  p.counters = L->d_cnt->cnt;
  p.tx_power_sum = &L->d_cnt->power_sum;
  p.tx_power_max_bits = &L->d_cnt->power_max_bits;
  // post-equaliser stage (ofdm_link_set_post); the replay entry points switch post_src to the recorded matrix
  p.post_sigma = L->d_post;
  p.post_scale = (float)std::pow(10.0, -snr_db / 20.0);
  p.post_src = SRC_PHILOX;
  p.z_scale = (float)L->z_scale;
  p.z_power = L->z_power;
}

void set_dump(LinkParams& p, const ofdm_link_dump* d) {
  if (!d) return;
  p.dump_y = reinterpret_cast<float2*>(d->y);
  p.dump_z = reinterpret_cast<float2*>(d->z);
  p.dump_rx = d->rx_labels;
  p.dump_tx = d->tx_labels;
  p.dump_noise = reinterpret_cast<float2*>(d->noise);
}

// parameter block of the fast kernel (link_fast.cuh) for one SNR point
void fill_fast(const ofdm_link* L, FastParams& f, double snr_db, const ofdm_link_dump* dump_dev) {
  const int N = L->d.n_subcarriers, M = L->fast == 1 ? L->fixed_order : 1;
  int half_bits = 0;
  while ((1 << (2 * half_bits)) < M) ++half_bits;
  const int side = 1 << half_bits;
  std::memset(&f, 0, sizeof(f));
  f.k4b = 0x4B000000u;
  std::memcpy(f.taps, L->taps_fast, sizeof(f.taps));
  for (int l = 0; l < kFastTaps; ++l)   // Gauss form of the complex product (link_fast.cuh, FIR stage)
    f.taps3[l] = make_float4(L->taps_fast[l].x, L->taps_fast[l].y - L->taps_fast[l].x, L->taps_fast[l].x + L->taps_fast[l].y, 0.f);
  f.eq_tab = L->d_eq_fast;
  f.tw = L->d_tw_fast;
  f.field_masks = L->d_mask;
  f.bit_offsets = L->d_bitoff;
  f.bits_per_ofdm = (unsigned)L->bits_per_ofdm;
  f.level_tab = L->d_level;
  const double snr_lin = std::pow(10.0, snr_db / 10.0);
  // equalization/models.py:43-49 on the unscaled FFT output Y~ = sqrt(N) Y
  f.point_tab[0].mmse_c = L->d.equalizer != OFDM_EQ_MMSE ? 0.f
                          : L->mean_h2 == 0.0            ? INFINITY
                                                         : (float)(1.0 / (double(N) * double(N) * snr_lin * L->mean_h2));
  f.slice_top = float(side - 1);
  f.n_points = 1;
  const double tap_scale = L->d.modulator == OFDM_MOD_SC_OFDM ? 1.0 / L->knorm : 1.0 / (L->knorm * std::sqrt((double)N));
  f.tx_scale2 = (float)(tap_scale * tap_scale);
  f.z_unscale = (float)(2.0 * (side - 1) / L->knorm);
  f.y_scale = (float)(1.0 / std::sqrt((double)N));
  f.prefix_len = L->d.prefix_len;
  f.zero_prefix = L->d.prefix_type == OFDM_PREFIX_ZERO;
  f.equalizer = L->d.equalizer;
  f.half_bits = half_bits;
  f.field_mask = 0x01010101u * (unsigned)((side - 1) << 1);
  if (L->fast == 3) {
    int bps = 0;
    while ((1 << bps) < L->fixed_order) ++bps;
    f.psk_tab = L->d_psk;
    f.psk_bits = bps;
    f.psk_scale = (float)(double(L->fixed_order) / (2.0 * M_PI));
    f.field_mask = 0x01010101u * (unsigned)(L->fixed_order - 1);
    f.z_unscale = 1.f;
  }
  f.counters = L->d_cnt->cnt;
  if (dump_dev) {
    f.dump_y = reinterpret_cast<float2*>(dump_dev->y);
    f.dump_z = reinterpret_cast<float2*>(dump_dev->z);
    f.dump_rx = dump_dev->rx_labels;
    f.dump_tx = dump_dev->tx_labels;
    f.dump_noise = reinterpret_cast<float2*>(dump_dev->noise);
  }
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != prev && prev >= 0) cudaSetDevice(prev);
  }
};

// host-side mirror of the dump block with device buffers behind it
struct DumpStage {
  ofdm_link_dump dev{};
  const ofdm_link_dump* host = nullptr;
  size_t n_sym = 0;
  int N = 0, NP = 0;
  int alloc(const ofdm_link_dump* h, size_t n_symbols, int n, int np) {
    host = h;
    n_sym = n_symbols;
    N = n;
    NP = np;
    if (!h) return OFDM_OK;
    if (h->y) CUDA_TRY(cudaMalloc(&dev.y, n_sym * N * sizeof(float2)));
    if (h->z) CUDA_TRY(cudaMalloc(&dev.z, n_sym * N * sizeof(float2)));
    if (h->rx_labels) CUDA_TRY(cudaMalloc(&dev.rx_labels, n_sym * N * sizeof(uint16_t)));
    if (h->tx_labels) CUDA_TRY(cudaMalloc(&dev.tx_labels, n_sym * N * sizeof(uint16_t)));
    if (h->noise) {
      CUDA_TRY(cudaMalloc(&dev.noise, n_sym * NP * sizeof(float2)));
      CUDA_TRY(cudaMemset(dev.noise, 0, n_sym * NP * sizeof(float2)));
    }
    return OFDM_OK;
  }
  int fetch() {
    if (!host) return OFDM_OK;
    if (host->y) CUDA_TRY(cudaMemcpy(host->y, dev.y, n_sym * N * sizeof(float2), cudaMemcpyDeviceToHost));
    if (host->z) CUDA_TRY(cudaMemcpy(host->z, dev.z, n_sym * N * sizeof(float2), cudaMemcpyDeviceToHost));
    if (host->rx_labels)
      CUDA_TRY(cudaMemcpy(host->rx_labels, dev.rx_labels, n_sym * N * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    if (host->tx_labels)
      CUDA_TRY(cudaMemcpy(host->tx_labels, dev.tx_labels, n_sym * N * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    if (host->noise) CUDA_TRY(cudaMemcpy(host->noise, dev.noise, n_sym * NP * sizeof(float2), cudaMemcpyDeviceToHost));
    return OFDM_OK;
  }
  ~DumpStage() {
    cudaFree(dev.y);
    cudaFree(dev.z);
    cudaFree(dev.rx_labels);
    cudaFree(dev.tx_labels);
    cudaFree(dev.noise);
  }
};

void to_result(const ofdm_link* L, const CounterBlock& h, ofdm_link_result* out) {
  out->bit_errors = h.cnt[CNT_BIT_ERRORS];
  out->bits = h.cnt[CNT_BITS];
  out->symbol_errors = h.cnt[CNT_SYM_ERRORS];
  out->symbols = h.cnt[CNT_SYMBOLS];
  out->ofdm_symbols = h.cnt[CNT_OFDM_SYMBOLS];
  out->tx_samples = h.cnt[CNT_OFDM_SYMBOLS] * (uint64_t)(L->d.n_subcarriers + L->d.prefix_len);
  out->tx_power_sum = h.power_sum;
  double mx;
  std::memcpy(&mx, &h.power_max_bits, sizeof(mx));
  out->tx_power_max = mx;
}

// device block of `n_points` counter blocks for sweep launches, grown on demand
int reserve_sweep(ofdm_link* L, int n_points) {
  if (L->sweep_cap >= n_points) return OFDM_OK;
  if (L->d_sweep) cudaFree(L->d_sweep);
  L->d_sweep = nullptr;
  L->sweep_cap = 0;
  const int cap = n_points < 64 ? 64 : n_points;
  CUDA_TRY(cudaMalloc(&L->d_sweep, size_t(cap) * sizeof(CounterBlock)));
  L->sweep_cap = cap;
  return OFDM_OK;
}

int read_counters(ofdm_link* L, cudaStream_t stream, ofdm_link_result* out) {
  CounterBlock h;
  CUDA_TRY(cudaMemcpyAsync(&h, L->d_cnt, sizeof(h), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  to_result(L, h, out);
  return OFDM_OK;
}

// one block per counter block (SNR point): row b of the all-reduce payload
__global__ void pack_counters_kernel(const CounterBlock* blocks, double* payload, int rank, int world) {
  const CounterBlock* c = blocks + blockIdx.x;
  double* row = payload + (size_t)blockIdx.x * (9 + world);
  const int i = threadIdx.x;
  if (i < 8) row[i] = (double)c->cnt[i];
  else if (i == 8) row[8] = c->power_sum;
  else if (i < 9 + world) row[i] = (i - 9 == rank) ? __longlong_as_double((long long)c->power_max_bits) : 0.0;
}

__global__ void ffma_chain_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) out[0] = s;
}

}  // namespace

extern "C" {

const char* ofdm_b200_last_error(void) { return g_err; }
int ofdm_b200_abi_version(void) { return OFDM_B200_ABI_VERSION; }
uint64_t ofdm_b200_launch_count(void) { return g_launches.load(); }

int ofdm_b200_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    fail(OFDM_ECUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return OFDM_ECUDA;
  }
  return n;
}

int ofdm_link_create(const ofdm_link_desc* desc, const double* taps_chan, const double* h_eq, const int32_t* orders,
                     const double* amp, ofdm_link** out) {
  return ofdm_link_create_loaded(desc, taps_chan, h_eq, orders, amp, nullptr, out);
}

int ofdm_link_create_loaded(const ofdm_link_desc* desc, const double* taps_chan, const double* h_eq, const int32_t* orders,
                            const double* amp, const double* rx_gain, ofdm_link** out) {
  if (!desc || !taps_chan || !h_eq || !orders || !out) return fail(OFDM_EINVAL, "null argument");
  const int N = desc->n_subcarriers, Lt = desc->n_taps, P = desc->prefix_len;
  if (N < 8 || N > 8192 || (N & (N - 1))) return fail(OFDM_EUNSUPPORTED, "n_subcarriers=%d: need a power of two in 8..8192", N);
  if (Lt < 1 || Lt > kMaxTaps) return fail(OFDM_EUNSUPPORTED, "n_taps=%d: need 1..%d", Lt, kMaxTaps);
  if (Lt - 1 > N) return fail(OFDM_EUNSUPPORTED, "channel longer than one OFDM symbol");
  if (P < 0 || P > N) return fail(OFDM_EINVAL, "prefix_len=%d: need 0..N", P);
  if (desc->prefix_type < 0 || desc->prefix_type > 2 || desc->modulator < 0 || desc->modulator > 1 ||
      desc->equalizer < 0 || desc->equalizer > 2 || desc->scheme < 0 || desc->scheme > 1)
    return fail(OFDM_EINVAL, "enum value out of range");
  if (desc->prefix_type == OFDM_PREFIX_NONE && P != 0) return fail(OFDM_EINVAL, "prefix NONE needs prefix_len 0");

  int dev = desc->device;
  if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
  DeviceGuard guard(dev);

  // owns the link (and its arena) until the last step succeeded
  struct LinkOwner {
    ofdm_link* L;
    ~LinkOwner() {
      if (!L) return;
      g_arenas.release(L->arena, L->table_bytes, L->device);
      delete L;
    }
  } owner{new ofdm_link()};
  ofdm_link* L = owner.L;
  L->d = *desc;
  L->device = dev;
  // consecutive OFDM symbols interact when the prefix is shorter than the channel memory
  L->isi = (Lt - 1 > P) ? 1 : 0;
  L->rx_gain = rx_gain != nullptr;
  for (int l = 0; l < kMaxTaps; ++l)
    L->taps[l] = l < Lt ? make_float2((float)taps_chan[2 * l], (float)taps_chan[2 * l + 1]) : make_float2(0.f, 0.f);

  std::vector<float4> sc(N), eq(N);
  int bit_off = 0;
  double sum_h2 = 0.0;
  for (int k = 0; k < N; ++k) {
    const int M = orders[k];
    int bps = 0;
    if (M < 0 || (M & (M - 1)) || M > 65536) return fail(OFDM_EINVAL, "orders[%d]=%d is not a power of two", k, M);
    while ((1 << bps) < M) ++bps;
    if (M <= 1) bps = 0;
    if (desc->scheme == OFDM_SCHEME_QAM && (bps & 1)) return fail(OFDM_EINVAL, "orders[%d]=%d: QAM order must be a perfect square", k, M);
    const double knorm = desc->scheme == OFDM_SCHEME_QAM && M > 1 ? std::sqrt(2.0 * (M - 1) / 3.0) : 1.0;
    const double a = (amp ? amp[k] : 1.0);
    const unsigned info = (unsigned)bps | ((unsigned)bit_off << 8);
    float info_f;
    std::memcpy(&info_f, &info, 4);
    sc[k] = make_float4((float)(a / knorm), (float)knorm, info_f, rx_gain ? (float)rx_gain[k] : 1.f);
    bit_off += bps;
    const std::complex<double> H(h_eq[2 * k], h_eq[2 * k + 1]);
    sum_h2 += std::norm(H);
    if (desc->equalizer == OFDM_EQ_ZF) {
      // equalization/models.py:33-35: h = where(H == 0, 1e-10, H); Z = Y / h
      const std::complex<double> h = (H == std::complex<double>(0.0, 0.0)) ? std::complex<double>(1e-10, 0.0) : H;
      const std::complex<double> inv = 1.0 / h;
      eq[k] = make_float4((float)inv.real(), (float)inv.imag(), 0.f, 0.f);
    } else {
      eq[k] = make_float4((float)H.real(), (float)H.imag(), (float)std::norm(H), 0.f);
    }
  }
  L->bits_per_ofdm = bit_off;
  L->mean_h2 = sum_h2 / N;

  // ---- fast-path eligibility (link_fast.cuh) and its folded tables
  //      fast = 1: one QAM order on every subcarrier; fast = 2: per-subcarrier orders (adaptive loading)
  std::vector<float4> eq_fast_host;
  std::vector<float2> tw_fast_host, level_host, psk_host;
  std::vector<unsigned> mask_host;
  std::vector<unsigned short> bitoff_host;
  {
    bool uniform = true, loadable = true;
    for (int k = 0; k < N; ++k) {
      const int M = orders[k];
      uniform = uniform && (M == orders[0]);
      loadable = loadable && (M == 0 || M == 1 || M == 4 || M == 16 || M == 64 || M == 256);
    }
    const char* force = std::getenv("OFDM_B200_FORCE_GENERAL");
    const bool single_carrier = desc->modulator == OFDM_MOD_SC_OFDM;   // one order, no loading tables
    const bool shape_ok = desc->scheme == OFDM_SCHEME_QAM && (!single_carrier || (uniform_orders(orders, N) && !amp && !rx_gain)) &&
                          Lt <= kFastTaps && fast_supports_n(N) &&
                          // any guard interval: at least as long as the channel memory (cyclic or zero-padded), or shorter / none
                          // (inter-symbol interference: chained symbols)
                          P < N && !(force && force[0] == '1');
    L->fast = !shape_ok || !loadable ? 0 : (uniform && orders[0] >= 4 && !amp && !rx_gain) ? 1 : 2;
    // PSK: one order M = 2 .. 256 on every subcarrier (OFDM or SC-OFDM, any guard interval: the chained-symbol and the
    // single-carrier paths do not care how labels become points)
    const bool psk_ok = desc->scheme == OFDM_SCHEME_PSK && uniform && orders[0] >= 2 && orders[0] <= 256 && !amp && !rx_gain &&
                        P < N && Lt <= kFastTaps && fast_supports_n(N) && !(force && force[0] == '1') &&
                        fast_supports_combo(N, false, single_carrier, Lt - 1 > P, true);
    if (psk_ok) L->fast = 3;
    if (L->fast) {
      const double sqn = std::sqrt((double)N);
      L->fixed_order = (L->fast == 1 || L->fast == 3) ? orders[0] : 0;
      // levels are 2c-(s-1) = knorm * point and the IFFT is unnormalised: with one order the taps absorb
      // 1/(knorm sqrt N); with per-subcarrier orders 1/knorm_k is applied at the mapper (level_tab)
      L->knorm = L->fast == 1 ? std::sqrt(2.0 * (orders[0] - 1) / 3.0) : 1.0;
      // SC-OFDM: no transmitter transform, so the taps absorb 1/knorm only and the receiver's two transforms put
      // 1/N into the decision table
      const double tap_scale = single_carrier ? 1.0 / L->knorm : 1.0 / (L->knorm * sqn);
      for (int l = 0; l < kFastTaps; ++l)
        L->taps_fast[l] = l < Lt ? make_float2((float)(taps_chan[2 * l] * tap_scale), (float)(taps_chan[2 * l + 1] * tap_scale))
                                 : make_float2(0.f, 0.f);
      // decision = sat(Re/Im(Y~ conj A) / (G + sigma2) + 0.5) * (s-1):  k/2 (slicer), 1/sqrt(N) (receiver FFT)
      // and 1/(s-1) (unit interval for FFMA.SAT) folded into A; 4th component = s-1
      std::vector<float4> eqf(N);
      const int E = fast_samples_per_lane(N), T = N / E, Wd = T / E;
      level_host.assign(N, make_float2(0.f, -8388608.0f));
      mask_host.assign(size_t(E / 4) * T, 0u);
      bitoff_host.assign(N, 0);
      {
        unsigned off = 0;   // constellation/adaptive.py:178-198: bps_k bits per subcarrier, subcarrier-minor
        for (int k = 0; k < N; ++k) {
          bitoff_host[k] = (unsigned short)off;
          int b = 0;
          while ((1 << b) < orders[k]) ++b;
          off += orders[k] >= 4 ? (unsigned)b : 0u;
        }
      }
      if (L->fast == 3) {
        // PSK tables: A = H / sqrt(N) (receiver transform; H / N for SC-OFDM, whose second transform is unnormalised too),
        // label -> exp(j 2 pi gray^-1(label) / M)
        const int M = orders[0];
        const double rxn = single_carrier ? double(N) : sqn;
        for (int k = 0; k < N; ++k) {
          const std::complex<double> H(h_eq[2 * k], h_eq[2 * k + 1]);
          if (desc->equalizer == OFDM_EQ_NONE) eqf[k] = make_float4((float)(1.0 / rxn), 0.f, 1.f, 0.f);
          else if (desc->equalizer == OFDM_EQ_ZF && H == std::complex<double>(0.0, 0.0)) eqf[k] = make_float4((float)(1e10 / rxn), 0.f, 1.f, 0.f);
          else eqf[k] = make_float4((float)(H.real() / rxn), (float)(H.imag() / rxn), (float)std::norm(H), 0.f);
        }
        psk_host.assign(256, make_float2(1.f, 0.f));
        for (int lab = 0; lab < M; ++lab) {
          int kk = lab;
          for (int sh = 1; sh < 8; sh <<= 1) kk ^= kk >> sh;          // inverse Gray code
          const double ang = 2.0 * M_PI * double(kk) / double(M);
          psk_host[lab] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
      }
      for (int k = 0; k < N && L->fast != 3; ++k) {
        const int M = orders[k] < 4 ? 1 : orders[k];
        int side = 1;
        while (side * side < M) side *= 2;
        const double knorm_k = std::sqrt(2.0 * (M - 1) / 3.0);
        const float top = float(side - 1);
        const std::complex<double> H(h_eq[2 * k], h_eq[2 * k + 1]);
        if (side == 1) {
          eqf[k] = make_float4(0.f, 0.f, 1.f, 0.f);   // silent subcarrier: decision index 0, no errors counted
        } else {
          // applied power loading: tx amplitude in the level table, receiver gain in the decision-domain table
          const double dec = knorm_k / (2.0 * (single_carrier ? double(N) : sqn) * (side - 1)) * (rx_gain ? rx_gain[k] : 1.0);
          if (desc->equalizer == OFDM_EQ_NONE) {
            eqf[k] = make_float4((float)dec, 0.f, 1.f, top);
          } else if (desc->equalizer == OFDM_EQ_ZF && H == std::complex<double>(0.0, 0.0)) {
            eqf[k] = make_float4((float)(dec * 1e10), 0.f, 1.f, top);   // equalization/models.py:33-35: h := 1e-10
          } else {
            eqf[k] = make_float4((float)(H.real() * dec), (float)(H.imag() * dec), (float)std::norm(H), top);
          }
          level_host[k] = make_float2((float)((amp ? amp[k] : 1.0) / knorm_k), -(8388608.0f + float(side - 1)));
          const int t = k % T, m = k / T;
          mask_host[size_t(m / 4) * T + t] |= (unsigned)((side - 1) << 1) << (8 * (m % 4));
        }
      }
      eq_fast_host.swap(eqf);
      tw_fast_host = cached_twiddles(-N, [&] { return build_fast_twiddles(N); });
      if (L->fast != 2) { level_host.clear(); mask_host.clear(); bitoff_host.clear(); }
    }
  }

  int rc = configure(L);
  if (rc != OFDM_OK) return rc;
  CUDA_TRY(cudaDeviceGetAttribute(&L->sms, cudaDevAttrMultiProcessorCount, dev));

  // one device arena, one host->device copy: [counters | sc | eq | eq_fast | twiddles]
  const std::vector<float2>& tw = cached_twiddles(N, [&] { return compute_twiddles(N, L->E); });
  const size_t off_sc = 256, off_eq = off_sc + N * sizeof(float4), off_eqf = off_eq + N * sizeof(float4),
               off_tw = off_eqf + eq_fast_host.size() * sizeof(float4), off_twf = off_tw + tw.size() * sizeof(float2),
               off_lvl = off_twf + tw_fast_host.size() * sizeof(float2), off_msk = off_lvl + level_host.size() * sizeof(float2),
               off_bo = off_msk + mask_host.size() * sizeof(unsigned),
               off_psk = (off_bo + bitoff_host.size() * sizeof(unsigned short) + 15) & ~size_t(15),
               total = off_psk + psk_host.size() * sizeof(float2);
  // one pinned staging buffer -> one host->device copy
  std::vector<unsigned char> pageable;
  unsigned char* stage_ptr = g_stage.reserve(total);
  if (!stage_ptr) {            // no pinned memory to be had: pageable copy
    pageable.resize(total);
    stage_ptr = pageable.data();
  }
  std::memset(stage_ptr, 0, off_sc);
  std::memcpy(stage_ptr + off_sc, sc.data(), N * sizeof(float4));
  std::memcpy(stage_ptr + off_eq, eq.data(), N * sizeof(float4));
  if (!eq_fast_host.empty()) std::memcpy(stage_ptr + off_eqf, eq_fast_host.data(), eq_fast_host.size() * sizeof(float4));
  std::memcpy(stage_ptr + off_tw, tw.data(), tw.size() * sizeof(float2));
  if (!tw_fast_host.empty()) std::memcpy(stage_ptr + off_twf, tw_fast_host.data(), tw_fast_host.size() * sizeof(float2));
  if (!level_host.empty()) std::memcpy(stage_ptr + off_lvl, level_host.data(), level_host.size() * sizeof(float2));
  if (!mask_host.empty()) std::memcpy(stage_ptr + off_msk, mask_host.data(), mask_host.size() * sizeof(unsigned));
  if (!bitoff_host.empty()) std::memcpy(stage_ptr + off_bo, bitoff_host.data(), bitoff_host.size() * sizeof(unsigned short));
  if (!psk_host.empty()) std::memcpy(stage_ptr + off_psk, psk_host.data(), psk_host.size() * sizeof(float2));
  unsigned char* arena = g_arenas.acquire(total, dev);
  if (!arena) { cudaGetLastError(); return fail(OFDM_ENOMEM, "cudaMalloc(%zu bytes of link tables) failed", total); }
  L->arena = arena;
  L->table_bytes = total;
  CUDA_TRY(cudaMemcpyAsync(arena, stage_ptr, total, cudaMemcpyHostToDevice, nullptr));
  CUDA_TRY(cudaStreamSynchronize(nullptr));   // the staging buffer is reused by the next creation on this thread
  L->d_cnt = reinterpret_cast<CounterBlock*>(arena);
  L->d_sc = reinterpret_cast<float4*>(arena + off_sc);
  L->d_eq = reinterpret_cast<float4*>(arena + off_eq);
  L->d_eq_fast = eq_fast_host.empty() ? nullptr : reinterpret_cast<float4*>(arena + off_eqf);
  L->d_tw = reinterpret_cast<float2*>(arena + off_tw);
  L->d_tw_fast = tw_fast_host.empty() ? nullptr : reinterpret_cast<float2*>(arena + off_twf);
  L->d_level = level_host.empty() ? nullptr : reinterpret_cast<float2*>(arena + off_lvl);
  L->d_mask = mask_host.empty() ? nullptr : reinterpret_cast<unsigned*>(arena + off_msk);
  L->d_psk = psk_host.empty() ? nullptr : reinterpret_cast<float2*>(arena + off_psk);
  L->d_bitoff = bitoff_host.empty() ? nullptr : reinterpret_cast<unsigned short*>(arena + off_bo);
  owner.L = nullptr;
  *out = L;
  return OFDM_OK;
}

void ofdm_link_destroy(ofdm_link* L) {
  if (!L) return;
  DeviceGuard guard(L->device);
  // the cached arena may be handed to the next link at once: everything this link queued on the device must be done
  cudaDeviceSynchronize();
  if (L->d_sweep) cudaFree(L->d_sweep);
  if (L->d_post) cudaFree(L->d_post);
  g_arenas.release(L->arena, L->table_bytes, L->device);
  delete L;
}

int ofdm_link_bits_per_ofdm_symbol(const ofdm_link* L) { return L ? L->bits_per_ofdm : OFDM_EINVAL; }
int ofdm_link_uses_fast_kernel(const ofdm_link* L) { return L ? (L->fast != 0) : OFDM_EINVAL; }
uint64_t ofdm_link_table_bytes(const ofdm_link* L) { return L ? L->table_bytes : 0; }
void* ofdm_link_counters_device_ptr(ofdm_link* L) { return L ? (void*)L->d_cnt : nullptr; }

int ofdm_link_debug_tables(const ofdm_link* L, float* eq, float* level, uint32_t* masks, float* taps, float* taps3) {
  if (!L) return fail(OFDM_EINVAL, "null link");
  if (!L->fast) return fail(OFDM_EUNSUPPORTED, "this link runs on the general kernel: no folded tables");
  DeviceGuard guard(L->device);
  const int N = L->d.n_subcarriers;
  if (eq) CUDA_TRY(cudaMemcpy(eq, L->d_eq_fast, N * sizeof(float4), cudaMemcpyDeviceToHost));
  if (level) {
    if (!L->d_level) return fail(OFDM_EINVAL, "no level table: the link has one order on every subcarrier");
    CUDA_TRY(cudaMemcpy(level, L->d_level, N * sizeof(float2), cudaMemcpyDeviceToHost));
  }
  if (masks) {
    if (!L->d_mask) return fail(OFDM_EINVAL, "no field masks: the link has one order on every subcarrier");
    CUDA_TRY(cudaMemcpy(masks, L->d_mask, (N / 4) * sizeof(unsigned), cudaMemcpyDeviceToHost));
  }
  FastParams f;
  fill_fast(L, f, 0.0, nullptr);
  if (taps) std::memcpy(taps, f.taps, sizeof(f.taps));
  if (taps3) std::memcpy(taps3, f.taps3, sizeof(f.taps3));
  return OFDM_OK;
}

int ofdm_link_set_post(ofdm_link* L, const ofdm_link_post* post) {
  if (!L) return fail(OFDM_EINVAL, "null link");
  DeviceGuard guard(L->device);
  const int N = L->d.n_subcarriers;
  const bool active = post && (post->noise_profile || post->z_scale != 1.0 || post->measure_power);
  if (!active) {   // back to the plain chain (and to the fast kernel, if the shape has one)
    if (L->d_post) cudaFree(L->d_post);
    L->d_post = nullptr;
    L->post_recorded = nullptr;
    L->z_scale = 1.0;
    L->z_power = 0;
    if (L->fast_shape) L->fast = L->fast_shape;
    L->fast_shape = 0;
    return OFDM_OK;
  }
  if (!(post->z_scale > 0.0) || !std::isfinite(post->z_scale)) return fail(OFDM_EINVAL, "z_scale must be positive and finite");
  if (post->noise_profile) {
    std::vector<float> sig(N);
    for (int k = 0; k < N; ++k) {
      const double v = post->noise_profile[k];
      if (!(v >= 0.0) || !std::isfinite(v)) return fail(OFDM_EINVAL, "noise_profile[%d] must be finite and non-negative", k);
      sig[k] = (float)std::sqrt(v / 2.0);
    }
    if (!L->d_post) CUDA_TRY(cudaMalloc(&L->d_post, N * sizeof(float)));
    CUDA_TRY(cudaMemcpy(L->d_post, sig.data(), N * sizeof(float), cudaMemcpyHostToDevice));
  } else if (L->d_post) {
    cudaFree(L->d_post);
    L->d_post = nullptr;
  }
  L->post_recorded = post->noise_profile ? post->recorded_noise : nullptr;
  L->z_scale = post->z_scale;
  L->z_power = post->measure_power ? 1 : 0;
  if (L->fast) {   // the stage lives in the general kernel
    L->fast_shape = L->fast;
    L->fast = 0;
  }
  return OFDM_OK;
}

int ofdm_link_read_z_power(ofdm_link* L, void* stream, double* sum_abs2, uint64_t* n_values) {
  if (!L || !sum_abs2 || !n_values) return fail(OFDM_EINVAL, "null argument");
  DeviceGuard guard(L->device);
  CounterBlock h;
  CUDA_TRY(cudaMemcpyAsync(&h, L->d_cnt, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  std::memcpy(sum_abs2, &h.cnt[CNT_Z_POWER], sizeof(double));
  *n_values = h.cnt[CNT_Z_VALUES];
  return OFDM_OK;
}

int ofdm_link_reset_counters(ofdm_link* L, void* stream) {
  if (!L) return fail(OFDM_EINVAL, "null link");
  DeviceGuard guard(L->device);
  CUDA_TRY(cudaMemsetAsync(L->d_cnt, 0, sizeof(CounterBlock), (cudaStream_t)stream));
  return OFDM_OK;
}

int ofdm_link_pack_counters(ofdm_link* L, double* payload_row_dev, int32_t rank, int32_t world, void* stream) {
  if (!L || !payload_row_dev) return fail(OFDM_EINVAL, "null argument");
  if (world < 1 || world > 1000 || rank < 0 || rank >= world) return fail(OFDM_EINVAL, "rank %d of %d", rank, world);
  DeviceGuard guard(L->device);
  pack_counters_kernel<<<1, (9 + world + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(L->d_cnt, payload_row_dev, rank, world);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return OFDM_OK;
}

int ofdm_link_read_result(ofdm_link* L, void* stream, ofdm_link_result* out) {
  if (!L || !out) return fail(OFDM_EINVAL, "null argument");
  DeviceGuard guard(L->device);
  return read_counters(L, (cudaStream_t)stream, out);
}

int ofdm_link_launch_fused(ofdm_link* L, double snr_db, double noise_sigma, uint64_t seed, uint32_t point,
                           uint64_t first_symbol, uint64_t n_symbols, const ofdm_link_dump* dump_dev, void* stream) {
  if (!L) return fail(OFDM_EINVAL, "null link");
  if (L->bits_per_ofdm == 0) return fail(OFDM_EINVAL, "No active subcarriers (all orders are zero)");
  DeviceGuard guard(L->device);
  if (L->fast) {
    // common link shape: the fast kernel (link_fast.cuh); its Philox streams are its own
    FastParams f;
    fill_fast(L, f, snr_db, dump_dev);
    f.point_tab[0].sigma = (float)noise_sigma;
    f.seed = seed;
    f.point = point;
    f.sym_begin = first_symbol;
    f.sym_count = n_symbols;
    return launch_fast(L, f, dump_dev != nullptr, false, L->fast == 2, L->d.modulator == OFDM_MOD_SC_OFDM, L->isi != 0, L->fast == 3, (cudaStream_t)stream);
  }
  LinkParams p;
  fill_params(L, p, snr_db);
  p.bits_src = SRC_PHILOX;
  p.noise_src = noise_sigma > 0.0 ? SRC_PHILOX : SRC_NONE;
  p.sigma = (float)noise_sigma;
  p.seed = seed;
  p.point = point;
  p.sym_begin = first_symbol;
  p.sym_count = n_symbols;
  set_dump(p, dump_dev);
  return launch(L, p, (cudaStream_t)stream);
}

int ofdm_link_launch_replay(ofdm_link* L, double snr_db, const uint8_t* bits_dev, uint64_t n_bytes, const void* noise_dev,
                            int32_t noise_dtype, uint64_t n_symbols, uint64_t compare_limit_bits,
                            const ofdm_link_dump* dump_dev, void* stream) {
  if (!L || !bits_dev) return fail(OFDM_EINVAL, "null argument");
  if (noise_dtype != OFDM_NOISE_NONE && noise_dtype != OFDM_NOISE_C64 && noise_dtype != OFDM_NOISE_C128)
    return fail(OFDM_EINVAL, "noise_dtype=%d", noise_dtype);
  if (noise_dtype != OFDM_NOISE_NONE && !noise_dev) return fail(OFDM_EINVAL, "noise buffer missing");
  if (L->bits_per_ofdm == 0) return fail(OFDM_EINVAL, "No active subcarriers (all orders are zero)");
  DeviceGuard guard(L->device);
  const uint64_t whole = n_symbols * (uint64_t)L->bits_per_ofdm;
  if (L->fast != 0 && (compare_limit_bits == 0 || compare_limit_bits >= whole) && n_bytes * 8 >= whole &&
      (reinterpret_cast<uintptr_t>(bits_dev) & 3) == 0 && (reinterpret_cast<uintptr_t>(noise_dev) & 15) == 0) {
    // common link shape, whole OFDM symbols: the fast kernel streams the recorded bits and noise
    FastParams f;
    fill_fast(L, f, snr_db, dump_dev);
    f.bits = bits_dev;
    f.bits_len = n_bytes;
    f.noise = noise_dtype == OFDM_NOISE_NONE ? nullptr : noise_dev;
    f.noise_f64 = noise_dtype == OFDM_NOISE_C128;
    f.sym_count = n_symbols;
    return launch_fast(L, f, dump_dev != nullptr, true, L->fast == 2, L->d.modulator == OFDM_MOD_SC_OFDM, L->isi != 0, L->fast == 3,
                       (cudaStream_t)stream);
  }
  // Ragged recorded stream on a fast-kernel link (the comparison ends inside the LAST OFDM symbol, zip() truncation of
  // simulation/models.py:597): the whole symbols before it stream through the fast kernel, the last one runs on the
  // general kernel, which masks the bits past the limit; both add into the link's counter block.
  unsigned long long general_from = 0;
  if (L->fast != 0 && !L->d_post && n_symbols >= 2 && compare_limit_bits != 0 && compare_limit_bits < whole &&
      compare_limit_bits > whole - (uint64_t)L->bits_per_ofdm && n_bytes * 8 >= whole - (uint64_t)L->bits_per_ofdm &&
      (reinterpret_cast<uintptr_t>(bits_dev) & 3) == 0 && (reinterpret_cast<uintptr_t>(noise_dev) & 15) == 0) {
    FastParams f;
    fill_fast(L, f, snr_db, dump_dev);
    f.bits = bits_dev;
    f.bits_len = n_bytes;
    f.noise = noise_dtype == OFDM_NOISE_NONE ? nullptr : noise_dev;
    f.noise_f64 = noise_dtype == OFDM_NOISE_C128;
    f.sym_count = n_symbols - 1;
    const int rc_fast = launch_fast(L, f, dump_dev != nullptr, true, L->fast == 2, L->d.modulator == OFDM_MOD_SC_OFDM, L->isi != 0,
                                    L->fast == 3, (cudaStream_t)stream);
    if (rc_fast) return rc_fast;
    general_from = n_symbols - 1;
  }
  LinkParams p;
  fill_params(L, p, snr_db);
  p.sym_lo = general_from;
  p.bits_src = SRC_REPLAY_F32;
  p.noise_src = noise_dtype;
  p.bits = bits_dev;
  p.bits_len = n_bytes;
  p.noise = noise_dev;
  p.sym_begin = 0;
  p.sym_count = n_symbols;
  p.limit_bits = compare_limit_bits != 0;
  p.compare_limit = compare_limit_bits;
  if (L->d_post) {   // recorded post-equaliser noise, or none (a replayed run never draws from Philox)
    p.post_src = SRC_REPLAY_F64;
    p.post_noise = L->post_recorded;
    if (!L->post_recorded) p.post_sigma = nullptr;
  }
  set_dump(p, dump_dev);
  return launch(L, p, (cudaStream_t)stream);
}

int ofdm_link_launch_sweep(ofdm_link* L, int32_t n_points, const double* snr_db, const double* noise_sigma, uint64_t seed,
                           uint32_t first_point, uint64_t first_symbol, uint64_t n_symbols, void* stream_) {
  if (!L || !snr_db || !noise_sigma) return fail(OFDM_EINVAL, "null argument");
  if (n_points < 1 || n_points > 65536) return fail(OFDM_EINVAL, "n_points=%d: need 1..65536", n_points);
  if (L->bits_per_ofdm == 0) return fail(OFDM_EINVAL, "No active subcarriers (all orders are zero)");
  DeviceGuard guard(L->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = reserve_sweep(L, n_points);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(L->d_sweep, 0, size_t(n_points) * sizeof(CounterBlock), stream));
  L->sweep_points = n_points;
  if (L->fast) {
    // up to kMaxSweepPoints SNR points per launch: a point is a slice of the grid, its table rides in the parameter block
    for (int base = 0; base < n_points; base += kMaxSweepPoints) {
      const int k = n_points - base < kMaxSweepPoints ? n_points - base : kMaxSweepPoints;
      FastParams f;
      fill_fast(L, f, snr_db[base], nullptr);
      for (int i = 0; i < k; ++i) {
        FastParams g;
        fill_fast(L, g, snr_db[base + i], nullptr);    // the MMSE constant of this point
        f.point_tab[i].sigma = (float)noise_sigma[base + i];
        f.point_tab[i].mmse_c = g.point_tab[0].mmse_c;
      }
      f.n_points = (unsigned)k;
      f.counters = reinterpret_cast<unsigned long long*>(L->d_sweep + base);
      f.seed = seed;
      f.point = first_point + (uint32_t)base;
      f.sym_begin = first_symbol;
      f.sym_count = n_symbols;
      rc = launch_fast(L, f, false, false, L->fast == 2, L->d.modulator == OFDM_MOD_SC_OFDM, L->isi != 0, L->fast == 3, stream);
      if (rc) return rc;
    }
    return OFDM_OK;
  }
  for (int i = 0; i < n_points; ++i) {   // general kernel: one launch per point into the point's counter block
    LinkParams p;
    fill_params(L, p, snr_db[i]);
    p.bits_src = SRC_PHILOX;
    p.noise_src = noise_sigma[i] > 0.0 ? SRC_PHILOX : SRC_NONE;
    p.sigma = (float)noise_sigma[i];
    p.seed = seed;
    p.point = first_point + (uint32_t)i;
    p.sym_begin = first_symbol;
    p.sym_count = n_symbols;
    p.counters = L->d_sweep[i].cnt;
    p.tx_power_sum = &L->d_sweep[i].power_sum;
    p.tx_power_max_bits = &L->d_sweep[i].power_max_bits;
    rc = launch(L, p, stream);
    if (rc) return rc;
  }
  return OFDM_OK;
}

int ofdm_link_read_sweep(ofdm_link* L, void* stream, int32_t n_points, ofdm_link_result* out) {
  if (!L || !out) return fail(OFDM_EINVAL, "null argument");
  if (n_points < 1 || n_points > L->sweep_points) return fail(OFDM_EINVAL, "n_points=%d: the last sweep launch had %d", n_points, L->sweep_points);
  DeviceGuard guard(L->device);
  std::vector<CounterBlock> h(n_points);
  CUDA_TRY(cudaMemcpyAsync(h.data(), L->d_sweep, size_t(n_points) * sizeof(CounterBlock), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  for (int i = 0; i < n_points; ++i) to_result(L, h[i], &out[i]);
  return OFDM_OK;
}

int ofdm_link_pack_sweep(ofdm_link* L, double* payload_dev, int32_t rank, int32_t world, void* stream) {
  if (!L || !payload_dev) return fail(OFDM_EINVAL, "null argument");
  if (world < 1 || world > 1000 || rank < 0 || rank >= world) return fail(OFDM_EINVAL, "rank %d of %d", rank, world);
  if (L->sweep_points < 1) return fail(OFDM_EINVAL, "no sweep launch to pack");
  DeviceGuard guard(L->device);
  pack_counters_kernel<<<L->sweep_points, (9 + world + 31) / 32 * 32, 0, (cudaStream_t)stream>>>(L->d_sweep, payload_dev, rank, world);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return OFDM_OK;
}

int ofdm_link_run_sweep(ofdm_link* L, int32_t n_points, const double* snr_db, const double* noise_sigma, uint64_t seed,
                        uint32_t first_point, uint64_t first_symbol, uint64_t n_symbols, ofdm_link_result* out) {
  if (!L || !out) return fail(OFDM_EINVAL, "null argument");
  int rc = ofdm_link_launch_sweep(L, n_points, snr_db, noise_sigma, seed, first_point, first_symbol, n_symbols, nullptr);
  if (rc) return rc;
  return ofdm_link_read_sweep(L, nullptr, n_points, out);
}

int ofdm_link_run_fused(ofdm_link* L, double snr_db, double noise_sigma, uint64_t seed, uint32_t point,
                        uint64_t first_symbol, uint64_t n_symbols, const ofdm_link_dump* dump, ofdm_link_result* out) {
  if (!L || !out) return fail(OFDM_EINVAL, "null argument");
  DeviceGuard guard(L->device);
  DumpStage stage;
  int rc = stage.alloc(dump, n_symbols, L->d.n_subcarriers, L->d.n_subcarriers + L->d.prefix_len);
  if (rc) return rc;
  if ((rc = ofdm_link_reset_counters(L, nullptr))) return rc;
  if ((rc = ofdm_link_launch_fused(L, snr_db, noise_sigma, seed, point, first_symbol, n_symbols, dump ? &stage.dev : nullptr, nullptr)))
    return rc;
  if ((rc = read_counters(L, nullptr, out))) return rc;
  return stage.fetch();
}

int ofdm_link_run_replay(ofdm_link* L, double snr_db, const uint8_t* bits, uint64_t n_bytes, const void* noise,
                         int32_t noise_dtype, uint64_t n_symbols, uint64_t compare_limit_bits,
                         const ofdm_link_dump* dump, ofdm_link_result* out) {
  if (!L || !out || !bits) return fail(OFDM_EINVAL, "null argument");
  DeviceGuard guard(L->device);
  const size_t np = L->d.n_subcarriers + L->d.prefix_len;
  const size_t noise_bytes = noise_dtype == OFDM_NOISE_C64 ? n_symbols * np * 8 : noise_dtype == OFDM_NOISE_C128 ? n_symbols * np * 16 : 0;
  uint8_t* d_bits = nullptr;
  void* d_noise = nullptr;
  DumpStage stage;
  int rc = stage.alloc(dump, n_symbols, L->d.n_subcarriers, (int)np);
  if (rc) return rc;
  CUDA_TRY(cudaMalloc(&d_bits, n_bytes ? n_bytes : 1));
  if (noise_bytes) {
    if (!noise) { cudaFree(d_bits); return fail(OFDM_EINVAL, "noise buffer missing"); }
    cudaError_t e = cudaMalloc(&d_noise, noise_bytes);
    if (e != cudaSuccess) { cudaFree(d_bits); return fail(OFDM_ENOMEM, "cudaMalloc(noise) failed: %s", cudaGetErrorString(e)); }
  }
  auto cleanup = [&]() { cudaFree(d_bits); cudaFree(d_noise); };
  cudaError_t e = cudaMemcpy(d_bits, bits, n_bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && noise_bytes) e = cudaMemcpy(d_noise, noise, noise_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cleanup(); return fail(OFDM_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e)); }
  // recorded post-equaliser noise: HOST memory here, staged on the device for the launch
  const void* post_host = L->post_recorded;
  void* d_post_noise = nullptr;
  if (L->d_post && post_host) {
    const size_t bytes = n_symbols * size_t(L->d.n_subcarriers) * 16;
    e = cudaMalloc(&d_post_noise, bytes ? bytes : 16);
    if (e == cudaSuccess) e = cudaMemcpy(d_post_noise, post_host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cleanup(); cudaFree(d_post_noise); return fail(OFDM_ECUDA, "staging the recorded post-equaliser noise failed: %s", cudaGetErrorString(e)); }
    L->post_recorded = d_post_noise;
  }
  rc = ofdm_link_reset_counters(L, nullptr);
  if (!rc) rc = ofdm_link_launch_replay(L, snr_db, d_bits, n_bytes, d_noise, noise_dtype, n_symbols, compare_limit_bits,
                                        dump ? &stage.dev : nullptr, nullptr);
  if (!rc) rc = read_counters(L, nullptr, out);
  if (!rc) rc = stage.fetch();
  L->post_recorded = post_host;
  cudaFree(d_post_noise);
  cleanup();
  return rc;
}

double ofdm_b200_measure_fp32_tflops(int32_t iters) {
  if (iters <= 0) iters = 4096;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    fail(OFDM_ECUDA, "no CUDA device");
    return -1.0;
  }
  float* d = nullptr;
  if (cudaMalloc(&d, 4) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    ffma_chain_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * 8.0 * 16.0 * double(iters) * double(blocks) * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  return best;
}

}  // extern "C"
