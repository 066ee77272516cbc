/* ofdm_b200.h - C ABI of the B200-native OFDM link simulator (libofdm_b200.so).
 *
 * The reference (JomarJunior/ofdm-based-systems) is pure Python and has no FFI: its seam for this path
 * is the class registry of `Simulation` (src/ofdm_based_systems/simulation/models.py:73-103) and the
 * body of `Simulation.run()` between :454 and :606.  Every entry point below names the reference
 * lines it replaces.  Plain pointers and sizes only; no torch types.  All functions return 0 on
 * success or a negative OFDM_E* code; ofdm_b200_last_error() gives the message of the last failure on
 * the calling thread.  INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 */
#ifndef OFDM_B200_H
#define OFDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDM_B200_ABI_VERSION 1

enum { OFDM_OK = 0, OFDM_EINVAL = -1, OFDM_ECUDA = -2, OFDM_EUNSUPPORTED = -3, OFDM_ENOMEM = -4 };

/* values of the reference's str-Enums (configuration/enums.py:4-67) as integers */
enum { OFDM_PREFIX_NONE = 0, OFDM_PREFIX_CYCLIC = 1, OFDM_PREFIX_ZERO = 2 };
enum { OFDM_MOD_OFDM = 0, OFDM_MOD_SC_OFDM = 1 };
enum { OFDM_EQ_NONE = 0, OFDM_EQ_ZF = 1, OFDM_EQ_MMSE = 2 };
enum { OFDM_SCHEME_QAM = 0, OFDM_SCHEME_PSK = 1 };
enum { OFDM_NOISE_NONE = 0, OFDM_NOISE_C64 = 2, OFDM_NOISE_C128 = 3 };

typedef struct ofdm_link ofdm_link; /* opaque: one configured link (tables resident in HBM) */

/* What Simulation.run() fixes before its hot loop (simulation/models.py:226-410). */
typedef struct ofdm_link_desc {
  int32_t n_subcarriers; /* power of two, 8 .. 8192                       (num_subcarriers, :110)      */
  int32_t prefix_type;   /* OFDM_PREFIX_*                                  (PREFIX_SCHEME_MAPPERS, :83) */
  int32_t prefix_len;    /* int(prefix_length_ratio * channel.order)       (:251-253)                   */
  int32_t modulator;     /* OFDM_MOD_*                                     (MODULATOR_SCHEME_MAPPERS)   */
  int32_t equalizer;     /* OFDM_EQ_*                                      (EQUALIZATOR_SCHEME_MAPPERS) */
  int32_t scheme;        /* OFDM_SCHEME_*                                  (CONSTELLATION_SCHEME_..)    */
  int32_t n_taps;        /* 1 .. 32                                                                      */
  int32_t device;        /* CUDA device ordinal, -1 = the calling thread's current device               */
} ofdm_link_desc;

typedef struct ofdm_link_result {
  uint64_t bit_errors;    /* simulation/models.py:597                                   */
  uint64_t bits;          /* bits actually compared                                     */
  uint64_t symbol_errors; /* :604-606                                                   */
  uint64_t symbols;       /* constellation symbols compared (inactive subcarriers too)  */
  uint64_t ofdm_symbols;
  uint64_t tx_samples;    /* ofdm_symbols * (N + P)                                     */
  double tx_power_sum;    /* sum |tx|^2 over every tx sample, prefix included (:519-522) */
  double tx_power_max;    /* max |tx|^2                                                  */
} ofdm_link_result;

/* Optional per-symbol dumps (device or host memory, see the call); any pointer may be NULL. */
typedef struct ofdm_link_dump {
  float* y;            /* [n_symbols][N] complex64: ortho FFT output before the equaliser            */
  float* z;            /* [n_symbols][N] complex64: what the demapper sees (results["received_symbols"]) */
  uint16_t* rx_labels; /* [n_symbols][N]                                                             */
  uint16_t* tx_labels; /* [n_symbols][N]                                                             */
  float* noise;        /* [n_symbols][N+P] complex64: the noise the fused mode generated             */
} ofdm_link_dump;

const char* ofdm_b200_last_error(void);
int ofdm_b200_abi_version(void);
/* number of CUDA devices visible, <0 on error (used by the host side to fail loudly without a GPU) */
int ofdm_b200_device_count(void);

/* Build one link: uploads the channel taps, the equaliser table and the per-subcarrier loading.
 *   taps_chan : n_taps complex128 (re,im interleaved), ALREADY unit energy  (channel/models.py:14-16,37-44)
 *   h_eq      : N complex128, the equaliser's frequency response = fft(RAW taps, N) (simulation/models.py:263-266)
 *   orders    : N constellation orders, 0 = subcarrier carries nothing (constellation/adaptive.py:52-80);
 *               a constant array reproduces the fixed mapper (constellation/models.py:150-321)
 *   amp       : N tx amplitude scale factors applied on top of the unit-power constellation, or NULL
 */
int ofdm_link_create(const ofdm_link_desc* desc, const double* taps_chan, const double* h_eq,
                     const int32_t* orders, const double* amp, ofdm_link** out);
/* Same with APPLIED power loading, which Simulation.run() never does (simulation/models.py:508) but the reference's
 * experiments do (examples/overview.py:142, examples/waterfilling_noise_bump_experiment.py:148, 165-169):
 *   amp     : N tx amplitudes sqrt(P_k) multiplying the unit-power constellation points
 *   rx_gain : N receiver gains multiplying the equalised subcarriers before the demapper (1 / sqrt(P_k), 1 where
 *             P_k ~ 0), or NULL = no compensation */
int ofdm_link_create_loaded(const ofdm_link_desc* desc, const double* taps_chan, const double* h_eq,
                            const int32_t* orders, const double* amp, const double* rx_gain, ofdm_link** out);
void ofdm_link_destroy(ofdm_link* link);
int ofdm_link_bits_per_ofdm_symbol(const ofdm_link* link);
/* 1 when this link shape runs on the register-resident fast kernel (csrc/link_fast.cuh), 0 when it runs on
 * the general kernel (csrc/link_kernel.cuh); both compute the same chain */
int ofdm_link_uses_fast_kernel(const ofdm_link* link);
/* bytes ofdm_link_create copied host -> device for this link (tables; bench.py's h2d accounting) */
uint64_t ofdm_link_table_bytes(const ofdm_link* link);

/* Fused Monte-Carlo mode: replaces simulation/models.py:454-606 for OFDM symbols
 * [first_symbol, first_symbol + n_symbols) of SNR point `point`; bits and AWGN come from
 * Philox4x32 (7 rounds on the fast kernel, 10 on the general one) keyed by `seed` with counter
 * (symbol, stream, point), so any sharding of the symbol
 * range over GPUs produces the same union.  noise_sigma is the per-component standard deviation
 * sqrt(mean|stream|^2 / snr_lin / 2) (noise/models.py:14-20), computed by the caller.
 * Synchronous: returns when `out` is filled.  dump pointers, if given, are HOST memory. */
int ofdm_link_run_fused(ofdm_link* link, double snr_db, double noise_sigma, uint64_t seed, uint32_t point,
                        uint64_t first_symbol, uint64_t n_symbols, const ofdm_link_dump* dump,
                        ofdm_link_result* out);

/* A whole BER-vs-SNR sweep of one link in ONE kernel launch: replaces the sequential loop over Simulation.run() in
 * SimulationRunner.run_all (main.py:234-240, one Simulation per SNR from create_from_simulation_settings,
 * simulation/models.py:155-212).  Point i runs OFDM symbols [first_symbol, first_symbol + n_symbols) at snr_db[i] /
 * noise_sigma[i] with the Philox streams of point index first_point + i, i.e. exactly what n_points calls of
 * ofdm_link_run_fused(point = first_point + i) would produce; the SNR point is a slice of the grid (32 points per launch,
 * more points are queued back to back).  Synchronous, HOST arrays; out[n_points]. */
int ofdm_link_run_sweep(ofdm_link* link, int32_t n_points, const double* snr_db, const double* noise_sigma,
                        uint64_t seed, uint32_t first_point, uint64_t first_symbol, uint64_t n_symbols,
                        ofdm_link_result* out);

/* Replay mode with HOST buffers: identical bits / noise as fed to the reference.
 *   bits  : the BytesIO content of generate_bits (bits_generation/models.py:27-55), MSB first
 *   noise : complex noise over the serial stream, (N+P) samples per OFDM symbol, as added by
 *           AWGNoiseModel.add_noise (noise/models.py:19-22); dtype OFDM_NOISE_C64 / _C128 / _NONE
 *   compare_limit_bits : bit positions >= this are not compared (zip() truncation, simulation/models.py:597);
 *                        0 = compare everything
 * Copies inputs to the device, runs, copies results (and dumps) back. */
int ofdm_link_run_replay(ofdm_link* link, double snr_db, const uint8_t* bits, uint64_t n_bytes,
                         const void* noise, int32_t noise_dtype, uint64_t n_symbols,
                         uint64_t compare_limit_bits, const ofdm_link_dump* dump, ofdm_link_result* out);

/* Asynchronous variants on DEVICE memory: enqueue on `stream` (a cudaStream_t, NULL = legacy default),
 * accumulate into the link's device-resident counters; ofdm_link_read_result synchronises the stream,
 * returns the totals since the last ofdm_link_reset_counters and leaves them untouched. */
int ofdm_link_launch_fused(ofdm_link* link, double snr_db, double noise_sigma, uint64_t seed, uint32_t point,
                           uint64_t first_symbol, uint64_t n_symbols, const ofdm_link_dump* dump_dev,
                           void* stream);
int ofdm_link_launch_replay(ofdm_link* link, double snr_db, const uint8_t* bits_dev, uint64_t n_bytes,
                            const void* noise_dev, int32_t noise_dtype, uint64_t n_symbols,
                            uint64_t compare_limit_bits, const ofdm_link_dump* dump_dev, void* stream);
int ofdm_link_reset_counters(ofdm_link* link, void* stream);
int ofdm_link_read_result(ofdm_link* link, void* stream, ofdm_link_result* out);
/* device address of the raw counter block (8 x uint64, then double sum, then uint64 max bits): lets the
 * multi-GPU host side all-reduce the counters in place with NCCL */
void* ofdm_link_counters_device_ptr(ofdm_link* link);
/* Packs the link's counters into one row of the all-reduce payload (DEVICE memory, 9 + world doubles): 8 counters and
 * the power sum as doubles (exact below 2^53), the power maximum in slot 9 + rank and zeros in the other ranks' slots,
 * so that ONE SUM all-reduce combines counters, sums and maxima of all ranks.  Asynchronous on `stream`. */
int ofdm_link_pack_counters(ofdm_link* link, double* payload_row_dev, int32_t rank, int32_t world, void* stream);
/* Asynchronous sweep: clears the link's per-point counter blocks, queues the sweep on `stream`.  snr_db / noise_sigma
 * are HOST arrays (read before the call returns).  ofdm_link_read_sweep synchronises the stream and returns the points
 * of the last sweep launch; ofdm_link_pack_sweep packs them into the all-reduce payload [n_points][9 + world] (DEVICE
 * memory, layout of ofdm_link_pack_counters per row) so that ONE all-reduce per sweep combines all ranks. */
int ofdm_link_launch_sweep(ofdm_link* link, int32_t n_points, const double* snr_db, const double* noise_sigma,
                           uint64_t seed, uint32_t first_point, uint64_t first_symbol, uint64_t n_symbols, void* stream);
int ofdm_link_read_sweep(ofdm_link* link, void* stream, int32_t n_points, ofdm_link_result* out);
int ofdm_link_pack_sweep(ofdm_link* link, double* payload_dev, int32_t rank, int32_t world, void* stream);
/* Post-equaliser stage of the reference's water-filling robustness experiment
 * (examples/waterfilling_noise_bump_experiment.py:163-183; the per-subcarrier variant of examples/overview.py:210-214
 * is the rx_gain of ofdm_link_create_loaded).  After the equaliser and before the demapper the link then
 *   1. adds coloured noise of variance 10^(-snr_db/10) * noise_profile[k] on subcarrier k (:163-171) - from Philox stream 3
 *      in fused mode, or the recorded matrix `recorded_noise` ([n_symbols][N] complex128, as the reference drew it) in
 *      replay mode (HOST memory for ofdm_link_run_replay, DEVICE memory for ofdm_link_launch_replay);
 *   2. applies rx_gain (:173-176);
 *   3. accumulates sum |z|^2 and the number of values when `measure_power` is set (:178-179, read them with
 *      ofdm_link_read_z_power);
 *   4. multiplies by z_scale (:180-181): the caller's second pass over the same seed / recorded streams with
 *      z_scale = 1 / sqrt(sum / count) of the first pass.
 * The reference normalises over the whole run, so the two passes are two launches over the same symbol range.
 * A link with a post stage runs on the general kernel.  post = NULL removes the stage. */
typedef struct ofdm_link_post {
  const double* noise_profile; /* [N] >= 0, or NULL = no injected noise                                      */
  const void* recorded_noise;  /* replay mode only, or NULL                                                   */
  double z_scale;              /* 1.0 = none                                                                  */
  int32_t measure_power;       /* 1: accumulate sum |z|^2 (before z_scale)                                    */
  int32_t reserved;
} ofdm_link_post;
int ofdm_link_set_post(ofdm_link* link, const ofdm_link_post* post);
/* sum |z|^2 and the number of equalised values since the last ofdm_link_reset_counters (the synchronous run_* calls
 * reset first); synchronises `stream` */
int ofdm_link_read_z_power(ofdm_link* link, void* stream, double* sum_abs2, uint64_t* n_values);
/* kernel launches issued by this library since load (for bench.py's gpu_launches) */
uint64_t ofdm_b200_launch_count(void);

/* Batched water-filling + gap-rule bit loading, one channel realisation per row (fp64 on the device).
 * Replaces, per realisation: gains = |fft(raw taps, N)|^2 (simulation/models.py:277-278),
 * WaterfillingPowerAllocation.allocate / _find_water_level (power_allocation/models.py:140-225) or
 * UniformPowerAllocation.allocate (:61-69), the reported water level (simulation/models.py:311-313) and
 * calculate_bit_loading_order per subcarrier (constellation/models.py:297-321 QAM, :459-474 PSK). */
typedef struct ofdm_waterfill_desc {
  int32_t n_subcarriers;
  int32_t n_taps;
  int32_t scheme;        /* OFDM_SCHEME_* selects the gap rule                                            */
  int32_t waterfilling;  /* 1 = water-filling, 0 = uniform allocation                                     */
  int32_t min_order;     /* honoured only when max_order > 0 (the reference never clamps)     */
  int32_t max_order;
  double snr_db;         /* N0 = 10^(-snr_db/10)                                                          */
  double total_power;    /* N in adaptive mode, 1.0 in fixed mode (simulation/models.py:297, 486)         */
  double gap;            /* QAM: Qinv(ser/4)^2 / 3;  PSK: Qinv(ser/2)^2 / (2 pi^2)   (caller: scipy norm.isf) */
  double tolerance;      /* bisection stop, default 1e-8                                                  */
  int32_t order_rule;    /* 0: gap rule above; 1: Shannon capacity rule, calculate_constellation_orders
                            (constellation/adaptive.py:271-329) with min_order / max_order always applied  */
  int32_t reserved;
  double capacity_scaling; /* order_rule 1: bits = capacity * capacity_scaling (JSON capacity_scaling_factor) */
} ofdm_waterfill_desc;
/* taps: [n][n_taps] complex128 RAW taps; outputs power [n][N] f64, orders [n][N] i32, water_level [n] f64 (NaN
 * when uniform), optional h_eq [n][N] complex128, iterations [n] i32 and capacity [n][N] f64 = log2(1 + P g / N0
 * + 1e-12) per subcarrier (calculate_capacity, power_allocation/models.py:262-294; summed over a row it is what
 * compare_allocations :296-334 compares).  HOST buffers, synchronous. */
int ofdm_waterfill_bitload_batched(const ofdm_waterfill_desc* desc, const double* taps, int64_t n_realisations,
                                   double* power, int32_t* orders, double* water_level, double* h_eq,
                                   int32_t* iterations, double* capacity);
/* same on DEVICE buffers, asynchronous on `stream` */
int ofdm_waterfill_bitload_batched_dev(const ofdm_waterfill_desc* desc, const double* taps_dev, int64_t n_realisations,
                                       double* power_dev, int32_t* orders_dev, double* water_level_dev,
                                       double* h_eq_dev, int32_t* iterations_dev, double* capacity_dev, void* stream);

/* Frame batches: `n_frames` channel realisations, `symbols_per_frame` OFDM symbols each, in one launch of the link
 * kernel (what the reference would run as one Simulation per realisation, simulation/models.py:155-212, :454-606).
 * OFDM modulator, QAM, cyclic prefix >= channel memory, <= 8 taps, N a power of two in 64 .. 8192. */
typedef struct ofdm_frames_desc {
  int32_t n_subcarriers;
  int32_t prefix_len;
  int32_t equalizer;     /* OFDM_EQ_*                                                                          */
  int32_t n_taps;
  int32_t loading;       /* 0: fixed_order on every subcarrier; 1: per-frame gap-rule orders (adaptive mode,
                            simulation/models.py:277-352), bounded by min_order / max_order (max_order <= 256)  */
  int32_t fixed_order;   /* 4, 16, 64 or 256                                                                    */
  int32_t waterfilling;  /* loading 1: power fed to the gap rule, 1 = water-filling, 0 = uniform                */
  int32_t min_order;
  int32_t max_order;
  int32_t device;        /* CUDA device ordinal, -1 = current                                                   */
  double snr_db;
  double gap;            /* QAM gap Qinv(ser/4)^2 / 3 (loading 1)                                               */
} ofdm_frames_desc;
/* taps: [n_frames][n_taps] complex128 RAW taps (HOST), or NULL = a fresh Rayleigh realisation per frame drawn on the
 * device (examples/generate_channel_models.py:70-78: CN(0,1) sqrt(exp(-l/2)), unit energy) from Philox keyed by `seed`
 * with the global frame index first_frame + f.  OFDM symbol s of frame f uses the Philox counters of global symbol
 * (first_frame + f) * symbols_per_frame + s, so frames shard over GPUs like symbols do.
 * Outputs (HOST): total (sums; tx_power_max is the maximum), optional per_frame [n_frames], orders [n_frames][N] the
 * links ran with, taps_out [n_frames][n_taps] complex128. */
int ofdm_frames_run(const ofdm_frames_desc* desc, const double* taps, int64_t n_frames, uint64_t symbols_per_frame,
                    uint64_t seed, uint32_t point, uint64_t first_frame, ofdm_link_result* total,
                    ofdm_link_result* per_frame, int32_t* orders, double* taps_out);

/* Test hooks (tests/test_frames_gpu.py): the folded fp32 tables the register-resident kernel reads, as the device
 * holds them - for one link built on the host by ofdm_link_create (eq [N][4], level [N][2], masks [N/4], taps [8][2],
 * taps3 [8][4]; level / masks only for links with per-subcarrier tables), and for a frame batch built on the device
 * (eq [F][N][4], level [F][N][2], masks [F][N/4], hdr [F][ofdm_frames_header_floats()] = taps [8][2], taps3 [8][4],
 * sigma, mmse_c, 2 pads).  Any output pointer may be NULL.  HOST buffers, synchronous. */
int ofdm_link_debug_tables(const ofdm_link* link, float* eq, float* level, uint32_t* masks, float* taps, float* taps3);
int ofdm_frames_debug_tables(const ofdm_frames_desc* desc, const double* taps, int64_t n_frames, uint64_t seed,
                             uint64_t first_frame, float* eq, float* level, uint32_t* masks, float* hdr);
int ofdm_frames_header_floats(void);

/* FP32 FFMA-chain microbenchmark: returns measured TFLOP/s (2 flop per FFMA) on the current device,
 * the roofline denominator SURVEY 8(d) asks for; <0 on error. */
double ofdm_b200_measure_fp32_tflops(int32_t iters);

#ifdef __cplusplus
}
#endif
#endif /* OFDM_B200_H */
