// Frame batches: many channel realisations ("frames"), a fixed number of OFDM symbols each, in ONE launch of the
// fast link kernel (BASELINE config #4: "fresh channel realisation per frame"; config #2 with a fresh Rayleigh
// draw per frame).  The reference fixes one channel per Simulation (simulation/models.py:155-212) and would run
// one Simulation.run() per realisation; here everything between the taps and the error counters stays on the GPU:
//
//   rayleigh_taps_kernel      examples/generate_channel_models.py:70-78  (CN(0,1) sqrt(exp(-l/2)), unit energy)
//   waterfill_bitload_kernel  power_allocation/models.py:140-225, constellation/models.py:297-321 (waterfill.cu)
//   frame_tables_kernel       what ofdm_link_create does on the host for one link (simulation/models.py:248-266,
//                             channel/models.py:14-16, equalization/models.py:33-49, noise/models.py:14-20)
//   ofdm_link_fast_kernel<..., ADAPT, FRAMES>   simulation/models.py:454-606 per (frame, OFDM symbol)
//
// The Philox counters of symbol s of frame f are those of global symbol (first_frame + f) * symbols_per_frame + s,
// so a frame reproduces ofdm_link_run_fused of a single link built from the same taps and orders.
#include <cmath>
#include <cstring>
#include <vector>

#include "link_fast.cuh"
#include "plan.h"

namespace ofdm {

int launch_fast_frames(int n_subcarriers, int sms, const FastParams& p, cudaStream_t stream);   // link_fast.cu

namespace {

constexpr int kFtThreads = 256;

// taps[f][l] = (n1 + j n2) / sqrt(2) * sqrt(exp(-l / 2)), normalised to unit energy (fp64 Box-Muller on Philox words)
__global__ void rayleigh_taps_kernel(double2* taps, int n_taps, long long n_frames, unsigned long long first_frame,
                                     unsigned long long seed) {
  const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const unsigned long long gf = first_frame + (unsigned long long)f;
  const PhiloxKey key{(uint32_t)seed, (uint32_t)(seed >> 32)};
  double2 h[kFastTaps];
  double energy = 0.0;
  for (int l = 0; l < n_taps; ++l) {
    const uint4 w = philox4x32<10>(make_uint4((uint32_t)gf, (uint32_t)(gf >> 32), (3u << 28) | (uint32_t)l, 0xFFFFFFFFu), key);
    const double u1 = ((double)w.x * 4294967296.0 + (double)w.y + 0.5) * (1.0 / 18446744073709551616.0);
    const double u2 = ((double)w.z * 4294967296.0 + (double)w.w + 0.5) * (1.0 / 18446744073709551616.0);
    const double rad = sqrt(-2.0 * log(u1)) * sqrt(0.5 * exp(-0.5 * l));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    h[l] = make_double2(rad * cs, rad * sn);
    energy += h[l].x * h[l].x + h[l].y * h[l].y;
  }
  const double inv = 1.0 / sqrt(energy);
  for (int l = 0; l < n_taps; ++l) taps[f * n_taps + l] = make_double2(h[l].x * inv, h[l].y * inv);
}

struct FrameTableParams {
  const double2* taps;    // [F][L] raw taps
  const double2* h_eq;    // [F][N] fft(raw taps, N)
  const int* orders;      // [F][N] gap-rule orders (ignored with fixed loading)
  float4* eq;             // [F][N]
  float2* level;          // [F][N]
  unsigned* masks;        // [F][N/4]
  FrameHeader* hdr;       // [F]
  int* orders_used;       // [F][N] the orders the link runs with (may be null)
  int n, n_taps, E, equalizer, fixed_order;
  double snr_lin;
};

// one block per frame; the arithmetic is ofdm_link_create's (ofdm_b200.cu), in fp64, per frame
__global__ void __launch_bounds__(kFtThreads) frame_tables_kernel(const FrameTableParams p) {
  __shared__ double red[2][kFtThreads / 32];
  const long long f = blockIdx.x;
  const int N = p.n, T = N / p.E;
  const double2* taps = p.taps + f * p.n_taps;
  double energy = 0.0;
  for (int l = 0; l < p.n_taps; ++l) energy += taps[l].x * taps[l].x + taps[l].y * taps[l].y;
  const double inv_norm = 1.0 / sqrt(energy);          // channel/models.py:14-16
  const double sqn = sqrt((double)N);
  double act_pow = 0.0, sum_h2 = 0.0;
  for (int k = threadIdx.x; k < N; k += kFtThreads) {
    int M = p.fixed_order > 0 ? p.fixed_order : p.orders[f * N + k];
    if (M > 256) M = 256;                                // 4-bit level fields: the caller bounds the orders
    if (M < 4) M = 1;
    int side = 1;
    while (side * side < M) side *= 2;
    const double2 H = p.h_eq[f * N + k];
    const double g = H.x * H.x + H.y * H.y;
    sum_h2 += g;
    float4 e = make_float4(0.f, 0.f, 1.f, 0.f);
    float2 lv = make_float2(0.f, -8388608.0f);
    if (side > 1) {
      const double knorm = sqrt(2.0 * (M - 1) / 3.0);
      const double dec = knorm / (2.0 * sqn * (side - 1));
      const float top = float(side - 1);
      if (p.equalizer == OFDM_EQ_NONE) e = make_float4((float)dec, 0.f, 1.f, top);
      else if (p.equalizer == OFDM_EQ_ZF && g == 0.0) e = make_float4((float)(dec * 1e10), 0.f, 1.f, top);
      else e = make_float4((float)(H.x * dec), (float)(H.y * dec), (float)g, top);
      lv = make_float2((float)(1.0 / knorm), -(8388608.0f + float(side - 1)));
      act_pow += g * inv_norm * inv_norm;                // |fft(normalised taps)_k|^2 on the active subcarriers
    }
    p.eq[f * N + k] = e;
    p.level[f * N + k] = lv;
    if (p.orders_used) p.orders_used[f * N + k] = side > 1 ? M : 0;
  }
  // packed field masks: word j of lane t covers k = t + T (4 j + i)
  for (int w = threadIdx.x; w < N / 4; w += kFtThreads) {
    const int j = w / T, t = w % T;
    unsigned word = 0;
    for (int i = 0; i < 4; ++i) {
      const int k = t + T * (4 * j + i);
      int M = p.fixed_order > 0 ? p.fixed_order : p.orders[f * N + k];
      if (M > 256) M = 256;
      int side = 1;
      while (side * side < M) side *= 2;
      if (M < 4) side = 1;
      word |= (unsigned)((side - 1) << 1) << (8 * i);
    }
    p.masks[f * (N / 4) + w] = word;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    act_pow += __shfl_xor_sync(0xffffffffu, act_pow, off);
    sum_h2 += __shfl_xor_sync(0xffffffffu, sum_h2, off);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = act_pow;
    red[1][threadIdx.x >> 5] = sum_h2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, h2 = 0.0;
    for (int i = 0; i < kFtThreads / 32; ++i) { a += red[0][i]; h2 += red[1][i]; }
    FrameHeader hd;
    for (int l = 0; l < kFastTaps; ++l)
      hd.taps[l] = l < p.n_taps ? make_float2((float)(taps[l].x * inv_norm / sqn), (float)(taps[l].y * inv_norm / sqn))
                                : make_float2(0.f, 0.f);
    for (int l = 0; l < kFastTaps; ++l)   // same float arithmetic as fill_fast() in ofdm_b200.cu
      hd.taps3[l] = make_float4(hd.taps[l].x, hd.taps[l].y - hd.taps[l].x, hd.taps[l].x + hd.taps[l].y, 0.f);
    hd.sigma = (float)sqrt((a / N) / p.snr_lin / 2.0);   // noise/models.py:14-20 with the analytic stream power
    const double mean_h2 = h2 / N;
    hd.mmse_c = p.equalizer != OFDM_EQ_MMSE ? 0.f : mean_h2 == 0.0 ? INFINITY : (float)(1.0 / (double(N) * double(N) * p.snr_lin * mean_h2));
    hd.pad[0] = hd.pad[1] = 0.f;
    p.hdr[f] = hd;
  }
}

// scratch arena kept across calls (cudaMalloc / cudaFree of a few hundred MB cost more than the kernels); one per
// host thread, regrown on demand, released when the thread ends
struct DeviceArena {
  unsigned char* base = nullptr;
  size_t bytes = 0;
  int device = -1;
  int reserve(size_t need, int dev) {
    if (base && device == dev && bytes >= need) return OFDM_OK;
    if (base) cudaFree(base);
    base = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&base, need);
    if (e != cudaSuccess) return fail(OFDM_ENOMEM, "cudaMalloc(%zu bytes of frame tables) failed: %s", need, cudaGetErrorString(e));
    bytes = need;
    device = dev;
    return OFDM_OK;
  }
  ~DeviceArena() { if (base) cudaFree(base); }
};
thread_local DeviceArena g_arena;

}  // namespace
}  // namespace ofdm

using namespace ofdm;

namespace {
struct FrameTableDump {   // test hook: the per-frame tables as the device built them (HOST buffers, any may be null)
  float* eq;              // [F][N][4]
  float* level;           // [F][N][2]
  uint32_t* masks;        // [F][N/4]
  float* hdr;             // [F][sizeof(FrameHeader) / 4]
};

int frames_run_impl(const ofdm_frames_desc* d, const double* taps, int64_t n_frames, uint64_t symbols_per_frame,
                    uint64_t seed, uint32_t point, uint64_t first_frame, ofdm_link_result* total,
                    ofdm_link_result* per_frame, int32_t* orders_out, double* taps_out, const FrameTableDump* dump) {
  if (!d || !total) return fail(OFDM_EINVAL, "null argument");
  const int N = d->n_subcarriers, L = d->n_taps, P = d->prefix_len;
  if (!fast_supports_n(N)) return fail(OFDM_EUNSUPPORTED, "frame batches need n_subcarriers = a power of two in 64..8192, got %d", N);
  if (L < 1 || L > kFastTaps) return fail(OFDM_EUNSUPPORTED, "frame batches need 1..%d taps, got %d", kFastTaps, L);
  if (P < L - 1 || P >= N) return fail(OFDM_EUNSUPPORTED, "frame batches need a cyclic prefix with n_taps - 1 <= prefix_len < N");
  if (d->equalizer < 0 || d->equalizer > 2) return fail(OFDM_EINVAL, "equalizer=%d", d->equalizer);
  if (d->loading == 0) {
    const int M = d->fixed_order;
    if (M != 4 && M != 16 && M != 64 && M != 256) return fail(OFDM_EUNSUPPORTED, "fixed_order=%d: need 4, 16, 64 or 256", M);
  } else if (d->loading == 1) {
    if (d->max_order < 4 || d->max_order > 256 || d->min_order < 0 || d->min_order > d->max_order)
      return fail(OFDM_EUNSUPPORTED, "adaptive frame batches need 4 <= max_order <= 256 (honor_order_bounds)");
    if (!(d->gap > 0.0)) return fail(OFDM_EINVAL, "gap must be positive");
  } else {
    return fail(OFDM_EINVAL, "loading=%d", d->loading);
  }
  std::memset(total, 0, sizeof(*total));
  if (n_frames <= 0 || symbols_per_frame == 0) return OFDM_OK;

  int dev = d->device;
  if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != dev) CUDA_TRY(cudaSetDevice(dev));
  struct Restore { int prev, dev; ~Restore() { if (prev != dev && prev >= 0) cudaSetDevice(prev); } } restore{prev, dev};
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  const size_t F = (size_t)n_frames;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_taps = 0, o_pow = up(o_taps + F * L * 16), o_ord = up(o_pow + F * N * 8), o_lvl = up(o_ord + F * N * 4),
               o_heq = up(o_lvl + F * 8), o_eq = up(o_heq + F * N * 16), o_level = up(o_eq + F * N * 16),
               o_mask = up(o_level + F * N * 8), o_hdr = up(o_mask + F * N), o_cnt = up(o_hdr + F * sizeof(FrameHeader)),
               o_used = up(o_cnt + F * 80), bytes = o_used + (orders_out ? F * N * 4 : 0);
  const int E = fast_samples_per_lane(N), T = N / E, Wd = T / E;
  const size_t tw_count = size_t(E) * (E + 2) + (Wd > 1 ? N / Wd : 0), o_tw = up(bytes);
  {
    const int rc_arena = g_arena.reserve(o_tw + tw_count * sizeof(float2), dev);
    if (rc_arena) return rc_arena;
  }
  unsigned char* a = g_arena.base;
  cudaStream_t stream = nullptr;
  double2* d_taps = reinterpret_cast<double2*>(a + o_taps);
  if (taps) {
    CUDA_TRY(cudaMemcpyAsync(d_taps, taps, F * L * 16, cudaMemcpyHostToDevice, stream));
  } else {
    rayleigh_taps_kernel<<<(unsigned)((F + 127) / 128), 128, 0, stream>>>(d_taps, L, (long long)F, first_frame, seed);
    count_launch();
    CUDA_TRY(cudaGetLastError());
  }
  CUDA_TRY(cudaMemsetAsync(a + o_cnt, 0, F * 80, stream));

  // gains, (water-filling,) gap-rule orders and fft(raw taps, N) per frame
  ofdm_waterfill_desc wd;
  std::memset(&wd, 0, sizeof(wd));
  wd.n_subcarriers = N;
  wd.n_taps = L;
  wd.scheme = OFDM_SCHEME_QAM;
  wd.waterfilling = d->loading == 1 ? d->waterfilling : 0;
  wd.min_order = d->min_order;
  wd.max_order = d->loading == 1 ? d->max_order : 0;
  wd.snr_db = d->snr_db;
  wd.total_power = (double)N;                            // simulation/models.py:297
  wd.gap = d->loading == 1 ? d->gap : 1.0;
  wd.tolerance = 1e-8;
  int rc = ofdm_waterfill_bitload_batched_dev(&wd, reinterpret_cast<const double*>(d_taps), n_frames,
                                              reinterpret_cast<double*>(a + o_pow), reinterpret_cast<int32_t*>(a + o_ord),
                                              reinterpret_cast<double*>(a + o_lvl), reinterpret_cast<double*>(a + o_heq),
                                              nullptr, nullptr, stream);
  if (rc) return rc;

  FrameTableParams tp;
  tp.taps = d_taps;
  tp.h_eq = reinterpret_cast<const double2*>(a + o_heq);
  tp.orders = reinterpret_cast<const int*>(a + o_ord);
  tp.eq = reinterpret_cast<float4*>(a + o_eq);
  tp.level = reinterpret_cast<float2*>(a + o_level);
  tp.masks = reinterpret_cast<unsigned*>(a + o_mask);
  tp.hdr = reinterpret_cast<FrameHeader*>(a + o_hdr);
  tp.orders_used = orders_out ? reinterpret_cast<int*>(a + o_used) : nullptr;
  tp.n = N;
  tp.n_taps = L;
  tp.E = E;
  tp.equalizer = d->equalizer;
  tp.fixed_order = d->loading == 0 ? d->fixed_order : 0;
  tp.snr_lin = std::pow(10.0, d->snr_db / 10.0);
  frame_tables_kernel<<<(unsigned)F, kFtThreads, 0, stream>>>(tp);
  count_launch();
  CUDA_TRY(cudaGetLastError());

  if (dump) {
    if (dump->eq) CUDA_TRY(cudaMemcpyAsync(dump->eq, tp.eq, F * N * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    if (dump->level) CUDA_TRY(cudaMemcpyAsync(dump->level, tp.level, F * N * sizeof(float2), cudaMemcpyDeviceToHost, stream));
    if (dump->masks) CUDA_TRY(cudaMemcpyAsync(dump->masks, tp.masks, F * (N / 4) * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    if (dump->hdr) CUDA_TRY(cudaMemcpyAsync(dump->hdr, tp.hdr, F * sizeof(FrameHeader), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
  }

  // pass-2 / pass-3 twiddles of the fast transform (same table as ofdm_link_create builds)
  const std::vector<float2> tw = build_fast_twiddles(N);
  float2* d_tw = reinterpret_cast<float2*>(a + o_tw);
  CUDA_TRY(cudaMemcpyAsync(d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, stream));

  FastParams fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.k4b = 0x4B000000u;
  fp.eq_tab = tp.eq;
  fp.tw = d_tw;
  fp.field_masks = tp.masks;
  fp.level_tab = tp.level;
  fp.frame_hdr = tp.hdr;
  fp.frame_counters = reinterpret_cast<unsigned long long*>(a + o_cnt);
  fp.frame_syms = symbols_per_frame;
  fp.n_frames = (unsigned)F;
  fp.tx_scale2 = (float)(1.0 / N);
  fp.y_scale = (float)(1.0 / std::sqrt((double)N));
  fp.prefix_len = P;
  fp.equalizer = d->equalizer;
  fp.seed = seed;
  fp.point = point;
  fp.sym_begin = first_frame * symbols_per_frame;
  fp.sym_count = (unsigned long long)F * symbols_per_frame;
  rc = launch_fast_frames(N, sms, fp, stream);
  if (rc) return rc;

  std::vector<unsigned long long> cnt(F * 10);
  CUDA_TRY(cudaMemcpyAsync(cnt.data(), a + o_cnt, F * 80, cudaMemcpyDeviceToHost, stream));
  if (orders_out) CUDA_TRY(cudaMemcpyAsync(orders_out, a + o_used, F * N * 4, cudaMemcpyDeviceToHost, stream));
  if (taps_out) CUDA_TRY(cudaMemcpyAsync(taps_out, d_taps, F * L * 16, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  for (size_t f = 0; f < F; ++f) {
    const unsigned long long* c = &cnt[f * 10];
    ofdm_link_result r;
    r.bit_errors = c[CNT_BIT_ERRORS];
    r.bits = c[CNT_BITS];
    r.symbol_errors = c[CNT_SYM_ERRORS];
    r.symbols = c[CNT_SYMBOLS];
    r.ofdm_symbols = c[CNT_OFDM_SYMBOLS];
    r.tx_samples = c[CNT_OFDM_SYMBOLS] * (uint64_t)(N + P);
    std::memcpy(&r.tx_power_sum, &c[8], 8);
    std::memcpy(&r.tx_power_max, &c[9], 8);
    if (per_frame) per_frame[f] = r;
    total->bit_errors += r.bit_errors;
    total->bits += r.bits;
    total->symbol_errors += r.symbol_errors;
    total->symbols += r.symbols;
    total->ofdm_symbols += r.ofdm_symbols;
    total->tx_samples += r.tx_samples;
    total->tx_power_sum += r.tx_power_sum;
    if (r.tx_power_max > total->tx_power_max) total->tx_power_max = r.tx_power_max;
  }
  return OFDM_OK;
}
}  // namespace

extern "C" int ofdm_frames_run(const ofdm_frames_desc* d, const double* taps, int64_t n_frames, uint64_t symbols_per_frame,
                               uint64_t seed, uint32_t point, uint64_t first_frame, ofdm_link_result* total,
                               ofdm_link_result* per_frame, int32_t* orders_out, double* taps_out) {
  return frames_run_impl(d, taps, n_frames, symbols_per_frame, seed, point, first_frame, total, per_frame, orders_out, taps_out, nullptr);
}

extern "C" int ofdm_frames_debug_tables(const ofdm_frames_desc* d, const double* taps, int64_t n_frames, uint64_t seed,
                                        uint64_t first_frame, float* eq, float* level, uint32_t* masks, float* hdr) {
  const FrameTableDump dump{eq, level, masks, hdr};
  ofdm_link_result total;
  return frames_run_impl(d, taps, n_frames, 1, seed, 0, first_frame, &total, nullptr, nullptr, nullptr, &dump);
}

extern "C" int ofdm_frames_header_floats(void) { return (int)(sizeof(FrameHeader) / sizeof(float)); }
