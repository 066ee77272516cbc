// ofdm_waterfill_bitload_batched: per channel realisation (one CTA each)
//   gains g_k = |fft(raw taps, N)_k|^2                      simulation/models.py:277-278
//   floor_k = N0 / (g_k N), bisection on the water level    power_allocation/models.py:161, 178-225
//   P_k = max(0, mu - floor_k) rescaled to the budget       power_allocation/models.py:165-176
//   reported water level = mean(P_k + N0/g_k | P_k > 1e-10) simulation/models.py:311-313
//   gap-rule constellation order per subcarrier             constellation/models.py:297-321 (QAM), 459-474 (PSK)
//   or the Shannon-capacity rule                            constellation/adaptive.py:271-329
//   capacity per subcarrier log2(1 + P g / N0 + 1e-12)      power_allocation/models.py:262-294
// All arithmetic in fp64 (the reference is fp64 and the order decision is a rounding of log2(1 + snr/gap)).
#include <cmath>
#include <cstring>
#include <vector>

#include "plan.h"

namespace ofdm {

constexpr int kWfThreads = 256;

struct WaterfillParams {
  const double2* taps;   // [F][L] raw taps
  double2* h_eq;         // [F][N] fft(raw taps, N)   (may be null)
  double* power;         // [F][N]
  int* orders;           // [F][N]
  double* water_level;   // [F]   (NaN when not water-filling)
  int* iterations;       // [F]   bisection steps used (may be null)
  double* capacity;      // [F][N] (may be null)
  int n, n_taps, scheme, waterfilling, min_order, max_order, order_rule;
  double noise_power, total_power, gap, tolerance, ser, capacity_scaling;
  int smem_floors;       // 1: the N floors of the bisection live in dynamic shared memory (N <= 4096), 0: in `power`
};

__device__ __forceinline__ double block_sum(double x, double* scratch) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[warp] = x;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kWfThreads / 32; ++i) s += scratch[i];
  return s;
}

__device__ __forceinline__ double block_max(double x, double* scratch) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, off));
  const int warp = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[warp] = x;
  __syncthreads();
  double s = scratch[0];
#pragma unroll
  for (int i = 1; i < kWfThreads / 32; ++i) s = fmax(s, scratch[i]);
  return s;
}

// H_k = sum_l h_l w^l with w = exp(-2 pi i k / N): one sincospi, Horner in l (error ~ L * 1e-16)
__device__ __forceinline__ double2 channel_response(const double2* taps, int n_taps, int k, int n) {
  double sn, cs;
  sincospi(-2.0 * double(k) / double(n), &sn, &cs);
  double re = taps[n_taps - 1].x, im = taps[n_taps - 1].y;
  for (int l = n_taps - 2; l >= 0; --l) {
    const double r2 = re * cs - im * sn + taps[l].x;
    im = re * sn + im * cs + taps[l].y;
    re = r2;
  }
  return make_double2(re, im);
}

__device__ __forceinline__ int gap_rule_order(double snr, const WaterfillParams& p) {
  if (p.order_rule == 1) {
    // calculate_constellation_orders: bits = clip(capacity * scaling, 0, log2 max); QAM: floor to even, PSK: floor;
    // below log2 min -> nothing transmitted
    const double cap = log2(1.0 + snr + 1e-12);
    double b = fmin(fmax(cap * p.capacity_scaling, 0.0), log2((double)p.max_order));
    b = p.scheme == 0 ? floor(b / 2.0) * 2.0 : floor(b);
    if (b < log2((double)p.min_order) || b <= 0.0) return 0;
    return 1 << (int)b;
  }
  int bits;
  if (p.scheme == 0) {  // QAM: round-half-even of log2(1 + snr/gap), made even
    bits = (int)rint(log2(1.0 + snr / p.gap));
    if (bits % 2 != 0) bits -= 1;
  } else {              // PSK: p.gap holds gamma* = Qinv(ser/2)^2 / (2 pi^2)
    const double g = sqrt(snr * p.gap) / (1.0 - sqrt(p.gap / (snr + 1e-10)));
    bits = (int)floor(log2(1.0 + snr / (g + 1e-10)) + 1e-10);
  }
  if (bits <= 0) return 0;
  if (bits > 30) bits = 30;
  int order = 1 << bits;
  if (p.max_order > 0) {  // honor_order_bounds (BASELINE config #4: QPSK .. 256-QAM); off = reference behaviour
    if (order > p.max_order) order = p.max_order;
    if (order < p.min_order) order = 0;
  }
  return order;
}

__global__ void __launch_bounds__(kWfThreads) waterfill_bitload_kernel(const WaterfillParams p) {
  __shared__ double scratch[kWfThreads / 32];
  __shared__ double2 s_taps[kMaxTaps];
  const int f = blockIdx.x, n = p.n;
  if (threadIdx.x < p.n_taps) s_taps[threadIdx.x] = p.taps[(size_t)f * p.n_taps + threadIdx.x];
  __syncthreads();
  extern __shared__ double s_floor[];
  double* power = p.power + (size_t)f * n;
  // the floors are read ~50 times by the bisection: shared memory when they fit, else the output row holds them
  double* floors = p.smem_floors ? s_floor : power;
  int* orders = p.orders + (size_t)f * n;

  // ---- gains and floors
  double max_floor = 0.0;
  for (int k = threadIdx.x; k < n; k += kWfThreads) {
    const double2 h = channel_response(s_taps, p.n_taps, k, n);
    if (p.h_eq) p.h_eq[(size_t)f * n + k] = h;
    const double g = h.x * h.x + h.y * h.y;
    const double fl = p.noise_power / (g * n);
    floors[k] = fl;
    max_floor = fmax(max_floor, fl);
  }
  __syncthreads();

  double mu = 0.0;
  int iters = 0;
  if (p.waterfilling) {
    max_floor = block_max(max_floor, scratch);
    double lo = 0.0, hi = p.total_power + max_floor;
    mu = (lo + hi) / 2;
    for (iters = 1; iters <= 100; ++iters) {
      mu = (lo + hi) / 2;
      double s = 0.0;
      for (int k = threadIdx.x; k < n; k += kWfThreads) s += fmax(0.0, mu - floors[k]);
      s = block_sum(s, scratch);
      if (fabs(s - p.total_power) < p.tolerance) break;
      if (s < p.total_power) lo = mu; else hi = mu;
    }
    if (iters > 100) iters = 100;
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += kWfThreads) s += fmax(0.0, mu - floors[k]);
    s = block_sum(s, scratch);
    const double scale = s > 0.0 ? p.total_power / s : 1.0;
    double lvl = 0.0, cnt = 0.0;
    for (int k = threadIdx.x; k < n; k += kWfThreads) {
      const double fl = floors[k];
      const double pk = fmax(0.0, mu - fl) * scale;
      power[k] = pk;
      const double2 h = channel_response(s_taps, p.n_taps, k, n);
      const double g = h.x * h.x + h.y * h.y;
      if (pk > 1e-10) { lvl += pk + p.noise_power / g; cnt += 1.0; }
      orders[k] = gap_rule_order(pk * g / p.noise_power, p);
      if (p.capacity) p.capacity[(size_t)f * n + k] = log2(1.0 + pk * g / p.noise_power + 1e-12);
    }
    lvl = block_sum(lvl, scratch);
    cnt = block_sum(cnt, scratch);
    if (threadIdx.x == 0) p.water_level[f] = cnt > 0.0 ? lvl / cnt : nan("");
  } else {
    const double pk = p.total_power / n;   // UniformPowerAllocation (power_allocation/models.py:61-69)
    for (int k = threadIdx.x; k < n; k += kWfThreads) {
      const double2 h = channel_response(s_taps, p.n_taps, k, n);
      const double g = h.x * h.x + h.y * h.y;
      power[k] = pk;
      orders[k] = gap_rule_order(pk * g / p.noise_power, p);
      if (p.capacity) p.capacity[(size_t)f * n + k] = log2(1.0 + pk * g / p.noise_power + 1e-12);
    }
    if (threadIdx.x == 0) p.water_level[f] = nan("");
  }
  if (threadIdx.x == 0 && p.iterations) p.iterations[f] = iters;
}

// Narrow links (N <= 128, the reference's shipped configurations are N = 64): ONE WARP per channel realisation, the
// floors in registers, warp shuffles instead of block barriers - the block-per-realisation kernel above spends its time in
// the two __syncthreads of each of the ~50 bisection steps with 192 of its 256 threads idle.  Same arithmetic per
// subcarrier and the same bisection; only the order of the additions inside the sums differs.
constexpr int kWfWarpJ = 4;   // subcarriers per lane
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}
__global__ void __launch_bounds__(kWfThreads) waterfill_bitload_warp_kernel(const WaterfillParams p, long long n_frames) {
  __shared__ double2 s_taps[kWfThreads / 32][kMaxTaps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n = p.n;
  const long long f = (long long)blockIdx.x * (kWfThreads / 32) + w;
  if (f >= n_frames) return;            // whole warps leave: no block-wide barrier below
  for (int l = lane; l < p.n_taps; l += 32) s_taps[w][l] = p.taps[(size_t)f * p.n_taps + l];
  __syncwarp();
  double* power = p.power + (size_t)f * n;
  int* orders = p.orders + (size_t)f * n;
  double fl[kWfWarpJ], g[kWfWarpJ];
  double max_floor = 0.0;
#pragma unroll
  for (int j = 0; j < kWfWarpJ; ++j) {
    const int k = lane + 32 * j;
    fl[j] = 0.0;
    g[j] = 1.0;
    if (k < n) {
      const double2 h = channel_response(s_taps[w], p.n_taps, k, n);
      if (p.h_eq) p.h_eq[(size_t)f * n + k] = h;
      g[j] = h.x * h.x + h.y * h.y;
      fl[j] = p.noise_power / (g[j] * n);
      max_floor = fmax(max_floor, fl[j]);
    }
  }
  double mu = 0.0;
  int iters = 0;
  if (p.waterfilling) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) max_floor = fmax(max_floor, __shfl_xor_sync(0xffffffffu, max_floor, off));
    auto filled = [&](double level) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < kWfWarpJ; ++j)
        if (lane + 32 * j < n) s += fmax(0.0, level - fl[j]);
      return warp_sum(s);
    };
    double lo = 0.0, hi = p.total_power + max_floor;
    mu = (lo + hi) / 2;
    for (iters = 1; iters <= 100; ++iters) {
      mu = (lo + hi) / 2;
      const double s = filled(mu);
      if (fabs(s - p.total_power) < p.tolerance) break;
      if (s < p.total_power) lo = mu; else hi = mu;
    }
    if (iters > 100) iters = 100;
    const double s = filled(mu);
    const double scale = s > 0.0 ? p.total_power / s : 1.0;
    double lvl = 0.0, cnt = 0.0;
#pragma unroll
    for (int j = 0; j < kWfWarpJ; ++j) {
      const int k = lane + 32 * j;
      if (k < n) {
        const double pk = fmax(0.0, mu - fl[j]) * scale;
        power[k] = pk;
        if (pk > 1e-10) { lvl += pk + p.noise_power / g[j]; cnt += 1.0; }
        orders[k] = gap_rule_order(pk * g[j] / p.noise_power, p);
        if (p.capacity) p.capacity[(size_t)f * n + k] = log2(1.0 + pk * g[j] / p.noise_power + 1e-12);
      }
    }
    lvl = warp_sum(lvl);
    cnt = warp_sum(cnt);
    if (lane == 0) p.water_level[f] = cnt > 0.0 ? lvl / cnt : nan("");
  } else {
    const double pk = p.total_power / n;   // UniformPowerAllocation (power_allocation/models.py:61-69)
#pragma unroll
    for (int j = 0; j < kWfWarpJ; ++j) {
      const int k = lane + 32 * j;
      if (k < n) {
        power[k] = pk;
        orders[k] = gap_rule_order(pk * g[j] / p.noise_power, p);
        if (p.capacity) p.capacity[(size_t)f * n + k] = log2(1.0 + pk * g[j] / p.noise_power + 1e-12);
      }
    }
    if (lane == 0) p.water_level[f] = nan("");
  }
  if (lane == 0 && p.iterations) p.iterations[f] = iters;
}

}  // namespace ofdm

using namespace ofdm;

extern "C" {

int ofdm_waterfill_bitload_batched_dev(const ofdm_waterfill_desc* d, const double* taps_dev, int64_t n_realisations,
                                       double* power_dev, int32_t* orders_dev, double* water_level_dev,
                                       double* h_eq_dev, int32_t* iterations_dev, double* capacity_dev, void* stream) {
  if (!d || !taps_dev || !power_dev || !orders_dev || !water_level_dev) return fail(OFDM_EINVAL, "null argument");
  if (d->n_subcarriers < 1 || d->n_taps < 1 || d->n_taps > kMaxTaps) return fail(OFDM_EINVAL, "bad n_subcarriers / n_taps");
  if (d->total_power < 0) return fail(OFDM_EINVAL, "Total power must be non-negative, got %g", d->total_power);
  if (d->order_rule != 0 && d->order_rule != 1) return fail(OFDM_EINVAL, "order_rule=%d", d->order_rule);
  if (d->order_rule == 1 && (d->min_order < 2 || d->max_order < d->min_order))
    return fail(OFDM_EINVAL, "the capacity rule needs 2 <= min_order <= max_order");
  if (n_realisations <= 0) return OFDM_OK;
  WaterfillParams p;
  p.taps = reinterpret_cast<const double2*>(taps_dev);
  p.h_eq = reinterpret_cast<double2*>(h_eq_dev);
  p.power = power_dev;
  p.orders = orders_dev;
  p.water_level = water_level_dev;
  p.iterations = iterations_dev;
  p.capacity = capacity_dev;
  p.order_rule = d->order_rule;
  p.capacity_scaling = d->capacity_scaling;
  p.n = d->n_subcarriers;
  p.n_taps = d->n_taps;
  p.scheme = d->scheme;
  p.waterfilling = d->waterfilling;
  p.min_order = d->min_order;
  p.max_order = d->max_order;
  p.noise_power = std::pow(10.0, -d->snr_db / 10.0);   // simulation/models.py:279
  p.total_power = d->total_power;
  p.gap = d->gap;
  p.tolerance = d->tolerance > 0 ? d->tolerance : 1e-8;
  p.ser = 0.0;
  p.smem_floors = d->n_subcarriers <= 4096 ? 1 : 0;    // 32 KB of doubles: inside the 48 KB every kernel may use without opting in
  if (d->n_subcarriers <= 32 * kWfWarpJ) {
    const long long per_block = kWfThreads / 32;
    waterfill_bitload_warp_kernel<<<(unsigned)((n_realisations + per_block - 1) / per_block), kWfThreads, 0, (cudaStream_t)stream>>>(p, n_realisations);
  } else {
    waterfill_bitload_kernel<<<(unsigned)n_realisations, kWfThreads, p.smem_floors ? size_t(d->n_subcarriers) * sizeof(double) : 0,
                               (cudaStream_t)stream>>>(p);
  }
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return OFDM_OK;
}

int ofdm_waterfill_bitload_batched(const ofdm_waterfill_desc* d, const double* taps, int64_t n_realisations,
                                   double* power, int32_t* orders, double* water_level, double* h_eq,
                                   int32_t* iterations, double* capacity) {
  if (!d || !taps || !power || !orders || !water_level) return fail(OFDM_EINVAL, "null argument");
  const size_t F = (size_t)n_realisations, N = (size_t)d->n_subcarriers, L = (size_t)d->n_taps;
  if (F == 0) return OFDM_OK;
  unsigned char* arena = nullptr;
  auto up16 = [](size_t x) { return (x + 15) & ~size_t(15); };   // double2 accesses need 16-byte alignment
  const size_t o_taps = 0, o_pow = up16(o_taps + F * L * 16), o_ord = up16(o_pow + F * N * 8),
               o_lvl = up16(o_ord + F * N * 4), o_heq = up16(o_lvl + F * 8), o_it = up16(o_heq + (h_eq ? F * N * 16 : 0)),
               o_cap = up16(o_it + F * 4), total = o_cap + (capacity ? F * N * 8 : 0);
  CUDA_TRY(cudaMalloc(&arena, total));
  int rc = OFDM_OK;
  cudaError_t e = cudaMemcpy(arena + o_taps, taps, F * L * 16, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) rc = fail(OFDM_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  if (!rc)
    rc = ofdm_waterfill_bitload_batched_dev(d, reinterpret_cast<const double*>(arena + o_taps), n_realisations,
                                            reinterpret_cast<double*>(arena + o_pow), reinterpret_cast<int32_t*>(arena + o_ord),
                                            reinterpret_cast<double*>(arena + o_lvl),
                                            h_eq ? reinterpret_cast<double*>(arena + o_heq) : nullptr,
                                            reinterpret_cast<int32_t*>(arena + o_it),
                                            capacity ? reinterpret_cast<double*>(arena + o_cap) : nullptr, nullptr);
  auto back = [&](void* dst, size_t off, size_t bytes) {
    if (rc || !dst) return;
    cudaError_t ee = cudaMemcpy(dst, arena + off, bytes, cudaMemcpyDeviceToHost);
    if (ee != cudaSuccess) rc = fail(OFDM_ECUDA, "D2H copy failed: %s", cudaGetErrorString(ee));
  };
  back(power, o_pow, F * N * 8);
  back(orders, o_ord, F * N * 4);
  back(water_level, o_lvl, F * 8);
  back(h_eq, o_heq, F * N * 16);
  back(iterations, o_it, F * 4);
  back(capacity, o_cap, F * N * 8);
  cudaFree(arena);
  return rc;
}

}  // extern "C"
