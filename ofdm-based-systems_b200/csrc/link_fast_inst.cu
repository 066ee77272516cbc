// One translation unit per team shape: nvcc -DOFDM_FAST_E=<samples per lane> -DOFDM_FAST_T=<lanes per OFDM symbol>
#include <cstdlib>

#include "link_fast.cuh"
#include "plan.h"

#if !defined(OFDM_FAST_E) || !defined(OFDM_FAST_T)
#error "compile with -DOFDM_FAST_E=<samples per lane> -DOFDM_FAST_T=<lanes per OFDM symbol>"
#endif

namespace ofdm {

template <int E, int T>
int launch_frames_shape(int sms, const FastParams& p, cudaStream_t stream);

// frame batches (frames.cu): one block per (frame, chunk of symbols) unit, grid-stride over the units
template <>
int launch_frames_shape<OFDM_FAST_E, OFDM_FAST_T>(int sms, const FastParams& p0, cudaStream_t stream) {
  constexpr int E = OFDM_FAST_E, T = OFDM_FAST_T, BLOCK = 512, SYNC = 0;
  using G = FastGeometry<E, T, BLOCK>;
  auto kern = ofdm_link_fast_kernel<E, T, false, true, false, BLOCK, SYNC, true, true>;
  int occ = 1;
  const int rc = blocks_per_sm(kern, G::BLOCK, G::SMEM_BYTES, &occ);
  if (rc != OFDM_OK) return rc;
  FastParams p = p0;
  // split a frame into chunks (multiples of the symbols a block holds in flight) until every SM has ~2 units
  const unsigned long long slots = (unsigned long long)sms * occ, S = p.frame_syms;
  const unsigned long long max_chunks = (S + G::TEAMS - 1) / G::TEAMS;
  unsigned long long cpf = (2 * slots + p.n_frames - 1) / p.n_frames;
  if (cpf < 1) cpf = 1;
  if (cpf > max_chunks) cpf = max_chunks;
  unsigned long long chunk = (S + cpf - 1) / cpf;
  chunk = (chunk + G::TEAMS - 1) / G::TEAMS * G::TEAMS;
  cpf = (S + chunk - 1) / chunk;
  p.chunk_syms = (unsigned)chunk;
  p.chunks_per_frame = (unsigned)cpf;
  const unsigned long long units = cpf * p.n_frames;
  const unsigned long long grid = units < slots ? units : slots;
  kern<<<(unsigned)grid, G::BLOCK, G::SMEM_BYTES, stream>>>(p);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return OFDM_OK;
}

template <int E, int T, bool DUMP, bool REPLAY, bool ADAPT = false, bool SC = false, bool ISI = false, bool PSK = false,
          int TAPS = kFastTaps, int SYNC = 0, int OPT = kOptDefault>
static int launch_fast_kernel(const ofdm_link* L, const FastParams& p0, cudaStream_t stream) {
  constexpr int BLOCK = 512;
  using G = FastGeometry<E, T, BLOCK>;
  static_assert((OPT & kOptCommon) == kOptCommon, "the link's tables are laid out for the common formulation");
  auto kern = ofdm_link_fast_kernel<E, T, DUMP, true, REPLAY, BLOCK, SYNC, ADAPT, false, SC, ISI, PSK, kFastRounds, 2, TAPS, OPT>;
  int occ = 1;
  const int rc = blocks_per_sm(kern, G::BLOCK, G::SMEM_BYTES, &occ);
  if (rc != OFDM_OK) return rc;
  const unsigned long long need = (p0.sym_count + G::TEAMS - 1) / G::TEAMS;
  unsigned long long per_point = (unsigned long long)L->sms * occ;
  if (need < per_point) per_point = need;
  if (per_point == 0) return OFDM_OK;
  const unsigned points = p0.n_points > 1 ? p0.n_points : 1;
  // the kernel counts a team's symbols in 32 bits: ranges beyond 2^40 symbols are queued in pieces (same counters;
  // contiguous pieces keep an inter-symbol-interference chain intact because every piece recomputes its halo symbol)
  constexpr unsigned long long kPiece = 1ull << 40;
  for (unsigned long long done = 0; done < p0.sym_count; done += kPiece) {
    FastParams p = p0;
    p.sym_begin = p0.sym_begin + done;
    p.sym_count = p0.sym_count - done < kPiece ? p0.sym_count - done : kPiece;
    if (REPLAY && done) return fail(OFDM_EUNSUPPORTED, "recorded streams longer than 2^40 OFDM symbols");
    kern<<<dim3((unsigned)per_point, points), G::BLOCK, G::SMEM_BYTES, stream>>>(p);
    count_launch();
    CUDA_TRY(cudaGetLastError());
  }
  return OFDM_OK;
}

template <int E, int T>
int launch_fast_shape(const ofdm_link* L, const FastParams& p, bool dump, bool replay, bool adapt, bool sc, bool isi,
                      bool psk, cudaStream_t stream);

// Instantiations per team shape.  The headline path (one QAM order, OFDM, guard interval >= channel memory, Philox bits and
// noise, counters only) exists for 1, 4 and 8 evaluated taps; every other link shape has a counters-only fused kernel
// and ONE dump-capable kernel per input mode (recorded streams always run through the dump-capable one, its dump
// pointers may be null), except the headline shape whose recorded-stream path is the HBM-bandwidth benchmark.
template <>
int launch_fast_shape<OFDM_FAST_E, OFDM_FAST_T>(const ofdm_link* L, const FastParams& p, bool dump, bool replay,
                                                bool adapt, bool sc, bool isi, bool psk, cudaStream_t stream) {
  constexpr int E = OFDM_FAST_E, T = OFDM_FAST_T;
  // Teams wider than a warp (N >= 2048) carry fewer instantiations - they are 4 MB of SASS per shape: every channel is
  // evaluated with 8 taps, and M-PSK combined with SC-OFDM or inter-symbol interference stays on the general kernel
  // (fast_supports_combo() below tells ofdm_link_create, which decides the kernel of a link).
  constexpr bool kWide = T >= 64;
  if ((psk && adapt) || (adapt && sc) || !fast_supports_combo(E * T, adapt, sc, isi, psk))
    return fail(OFDM_EUNSUPPORTED, "this link shape runs on the general kernel");
#define OFDM_FAST_VARIANT(ADAPT_, SC_, ISI_, PSK_)                                                                     \
  do {                                                                                                                 \
    if (replay) return launch_fast_kernel<E, T, true, true, ADAPT_, SC_, ISI_, PSK_>(L, p, stream);                    \
    if (dump) return launch_fast_kernel<E, T, true, false, ADAPT_, SC_, ISI_, PSK_>(L, p, stream);                     \
    /* channels of at most 4 taps (every shipped model but two): the FIR skips the four exact zeros */                 \
    if constexpr (!kWide) {                                                                                            \
      if (L->d.n_taps <= 4) return launch_fast_kernel<E, T, false, false, ADAPT_, SC_, ISI_, PSK_, 4>(L, p, stream);   \
    }                                                                                                                  \
    return launch_fast_kernel<E, T, false, false, ADAPT_, SC_, ISI_, PSK_>(L, p, stream);                              \
  } while (0)
  // rare combinations: one counters-only kernel (8 evaluated taps) beside the dump-capable ones
#define OFDM_FAST_RARE(SC_, ISI_, PSK_)                                                                                \
  do {                                                                                                                 \
    if (replay) return launch_fast_kernel<E, T, true, true, false, SC_, ISI_, PSK_>(L, p, stream);                     \
    return dump ? launch_fast_kernel<E, T, true, false, false, SC_, ISI_, PSK_>(L, p, stream)                          \
                : launch_fast_kernel<E, T, false, false, false, SC_, ISI_, PSK_>(L, p, stream);                        \
  } while (0)
  if (adapt && isi) {                                          // per-subcarrier orders / power loading, prefix shorter than the channel memory
    if (replay) return launch_fast_kernel<E, T, true, true, true, false, true, false>(L, p, stream);
    return dump ? launch_fast_kernel<E, T, true, false, true, false, true, false>(L, p, stream)
                : launch_fast_kernel<E, T, false, false, true, false, true, false>(L, p, stream);
  }
  if constexpr (!kWide) {
    if (psk && isi && sc) OFDM_FAST_RARE(true, true, true);    // M-ary PSK on single-carrier symbols with inter-symbol interference
    if (psk && isi) OFDM_FAST_RARE(false, true, true);         // M-ary PSK, prefix shorter than the channel memory
    if (psk && sc) OFDM_FAST_RARE(true, false, true);          // M-ary PSK, single-carrier OFDM
  }
#undef OFDM_FAST_RARE
  if (psk) OFDM_FAST_VARIANT(false, false, false, true);       // M-ary PSK, one order
  if (isi && sc) OFDM_FAST_VARIANT(false, true, true, false);  // SC-OFDM with a prefix shorter than the channel memory
  if (isi) OFDM_FAST_VARIANT(false, false, true, false);       // prefix shorter than the channel memory: chained symbols
  if (sc) OFDM_FAST_VARIANT(false, true, false, false);        // single-carrier OFDM: one order on every sample
  if (adapt) OFDM_FAST_VARIANT(true, false, false, false);     // per-subcarrier orders / applied power loading
#undef OFDM_FAST_VARIANT
  if (replay) {   // the counters-only recorded-stream kernel is the HBM-bandwidth benchmark of the narrow shapes
    if constexpr (!kWide) {
      if (!dump) return launch_fast_kernel<E, T, false, true>(L, p, stream);
    }
    return launch_fast_kernel<E, T, true, true>(L, p, stream);
  }
  // the dump-capable kernel evaluates the FIR in the same form as the counters-only kernel of the link (plain complex
  // product for one tap, Gauss form otherwise), so that the two return identical counters
  if (dump) {
    if constexpr (!kWide) {
      if (L->d.n_taps <= 1) return launch_fast_kernel<E, T, true, false, false, false, false, false, 1, 0, kOptOneTap>(L, p, stream);
    }
    return launch_fast_kernel<E, T, true, false>(L, p, stream);
  }
  // The warps of a block run free (SYNC = 0).  Round 1 aligned the warps that share a scheduler at the section
  // boundaries (named barrier per scheduler, SYNC = 2): they then share instruction fetches (stall_no_instruction 0.16
  // instead of 0.49 per issue) but meet the shared-memory exchanges and the MUFU section together.  With the round-2
  // instruction stream (two Philox calls per 8 noise samples, Gauss-form FIR) free-running warps are faster on every
  // one-warp team shape: -3 % time at N = 1024, -10 % at N = 256, -14 % at N = 64 (profiles/r2_fast_kernel_history.md).
  constexpr int S = 0;
  // ZF / no equaliser: sigma2 = 0, so the per-symbol noise estimate of the MMSE form is dropped; one tap: the plain complex
  // product (4 FFMA) beats the Gauss form (3 FFMA + 3 FADD)
  constexpr int kZf = kOptDefault | kOptNoEstimate;
  const bool mmse = L->d.equalizer == OFDM_EQ_MMSE;
  if constexpr (!kWide) {
    if (L->d.n_taps <= 1)
      return mmse ? launch_fast_kernel<E, T, false, false, false, false, false, false, 1, S, kOptOneTap>(L, p, stream)
                  : launch_fast_kernel<E, T, false, false, false, false, false, false, 1, S, kOptOneTap | kOptNoEstimate>(L, p, stream);
    if (L->d.n_taps <= 4)
      return mmse ? launch_fast_kernel<E, T, false, false, false, false, false, false, 4>(L, p, stream)
                  : launch_fast_kernel<E, T, false, false, false, false, false, false, 4, S, kZf>(L, p, stream);
  }
  return mmse ? launch_fast_kernel<E, T, false, false>(L, p, stream)
              : launch_fast_kernel<E, T, false, false, false, false, false, false, 8, S, kZf>(L, p, stream);
}

}  // namespace ofdm
