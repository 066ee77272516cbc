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
  constexpr int E = OFDM_FAST_E, T = OFDM_FAST_T, BLOCK = 512, SYNC = T > 32 ? 0 : 2;
  using G = FastGeometry<E, T, BLOCK>;
  auto kern = ofdm_link_fast_kernel<E, T, false, true, false, BLOCK, SYNC, true, true>;
  static int occ = 0;
  if (occ == 0) {
    if (G::SMEM_BYTES > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::BLOCK, G::SMEM_BYTES));
    if (occ <= 0) occ = 1;
  }
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

template <int E, int T, bool DUMP, bool REPLAY, int BLOCK = 512, int SYNC = 2, bool ADAPT = false, bool SC = false,
          bool ISI = false, bool PSK = false>
static int launch_fast_kernel(const ofdm_link* L, const FastParams& p, cudaStream_t stream) {
  using G = FastGeometry<E, T, BLOCK>;
  auto kern = ofdm_link_fast_kernel<E, T, DUMP, true, REPLAY, BLOCK, SYNC, ADAPT, false, SC, ISI, PSK>;
  static int occ = 0;   // per process: attribute + occupancy query cost ~0.1 ms each
  if (occ == 0) {
    if (G::SMEM_BYTES > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::BLOCK, G::SMEM_BYTES));
    if (occ <= 0) occ = 1;
  }
  const unsigned long long need = (p.sym_count + G::TEAMS - 1) / G::TEAMS;
  unsigned long long grid = (unsigned long long)L->sms * occ;
  if (need < grid) grid = need;
  if (grid == 0) return OFDM_OK;
  kern<<<(unsigned)grid, G::BLOCK, G::SMEM_BYTES, stream>>>(p);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return OFDM_OK;
}

template <int E, int T>
int launch_fast_shape(const ofdm_link* L, const FastParams& p, bool dump, bool replay, bool adapt, bool sc, bool isi,
                      bool psk, cudaStream_t stream);

template <>
int launch_fast_shape<OFDM_FAST_E, OFDM_FAST_T>(const ofdm_link* L, const FastParams& p, bool dump, bool replay,
                                                bool adapt, bool sc, bool isi, bool psk, cudaStream_t stream) {
  constexpr int E = OFDM_FAST_E, T = OFDM_FAST_T;
  constexpr int SYNC_DEFAULT = T > 32 ? 0 : 2;
  if (psk) {  // M-ary PSK, one order
    if (adapt || sc || isi) return fail(OFDM_EUNSUPPORTED, "this PSK link shape runs on the general kernel");
    if (replay) return dump ? launch_fast_kernel<E, T, true, true, 512, SYNC_DEFAULT, false, false, false, true>(L, p, stream)
                            : launch_fast_kernel<E, T, false, true, 512, SYNC_DEFAULT, false, false, false, true>(L, p, stream);
    return dump ? launch_fast_kernel<E, T, true, false, 512, SYNC_DEFAULT, false, false, false, true>(L, p, stream)
                : launch_fast_kernel<E, T, false, false, 512, SYNC_DEFAULT, false, false, false, true>(L, p, stream);
  }
  if (isi && sc) {  // single-carrier OFDM with a prefix shorter than the channel memory
    if (adapt) return fail(OFDM_EUNSUPPORTED, "inter-symbol interference with loading tables runs on the general kernel");
    if (replay) return dump ? launch_fast_kernel<E, T, true, true, 512, SYNC_DEFAULT, false, true, true>(L, p, stream)
                            : launch_fast_kernel<E, T, false, true, 512, SYNC_DEFAULT, false, true, true>(L, p, stream);
    return dump ? launch_fast_kernel<E, T, true, false, 512, SYNC_DEFAULT, false, true, true>(L, p, stream)
                : launch_fast_kernel<E, T, false, false, 512, SYNC_DEFAULT, false, true, true>(L, p, stream);
  }
  if (isi) {  // prefix shorter than the channel memory: chained symbols
    if (adapt) return fail(OFDM_EUNSUPPORTED, "inter-symbol interference with loading tables runs on the general kernel");
    if (replay) return dump ? launch_fast_kernel<E, T, true, true, 512, SYNC_DEFAULT, false, false, true>(L, p, stream)
                            : launch_fast_kernel<E, T, false, true, 512, SYNC_DEFAULT, false, false, true>(L, p, stream);
    return dump ? launch_fast_kernel<E, T, true, false, 512, SYNC_DEFAULT, false, false, true>(L, p, stream)
                : launch_fast_kernel<E, T, false, false, 512, SYNC_DEFAULT, false, false, true>(L, p, stream);
  }
  if (sc) {  // single-carrier OFDM: one order on every sample
    if (adapt) return fail(OFDM_EUNSUPPORTED, "SC-OFDM with per-sample orders runs on the general kernel");
    if (replay) return dump ? launch_fast_kernel<E, T, true, true, 512, SYNC_DEFAULT, false, true>(L, p, stream)
                            : launch_fast_kernel<E, T, false, true, 512, SYNC_DEFAULT, false, true>(L, p, stream);
    return dump ? launch_fast_kernel<E, T, true, false, 512, SYNC_DEFAULT, false, true>(L, p, stream)
                : launch_fast_kernel<E, T, false, false, 512, SYNC_DEFAULT, false, true>(L, p, stream);
  }
  if (adapt) {  // per-subcarrier orders / applied power loading
    if (replay) return dump ? launch_fast_kernel<E, T, true, true, 512, SYNC_DEFAULT, true>(L, p, stream)
                            : launch_fast_kernel<E, T, false, true, 512, SYNC_DEFAULT, true>(L, p, stream);
    return dump ? launch_fast_kernel<E, T, true, false, 512, SYNC_DEFAULT, true>(L, p, stream)
                : launch_fast_kernel<E, T, false, false, 512, SYNC_DEFAULT, true>(L, p, stream);
  }
  if (replay) return dump ? launch_fast_kernel<E, T, true, true, 512, SYNC_DEFAULT>(L, p, stream)
                          : launch_fast_kernel<E, T, false, true, 512, SYNC_DEFAULT>(L, p, stream);
  if (dump) return launch_fast_kernel<E, T, true, false, 512, SYNC_DEFAULT>(L, p, stream);
  // One-warp teams: the warps that share a scheduler walk the code in step (named barrier per scheduler, SYNC = 2;
  // free-running warps lose ~3 % to instruction-cache misses at N = 1024, profiles/).  Multi-warp teams already meet
  // at their team barriers and lose ~5 % to an extra one (N = 4096), so they run with SYNC = 0.
  // OFDM_B200_FAST_VARIANT=4 / =2 force SYNC = 0 / 2 for experiments.
  static const int variant = [] { const char* v = std::getenv("OFDM_B200_FAST_VARIANT"); return v ? std::atoi(v) : 0; }();
  const int sync = variant == 4 ? 0 : variant == 2 ? 2 : SYNC_DEFAULT;
  return sync == 0 ? launch_fast_kernel<E, T, false, false, 512, 0>(L, p, stream)
                   : launch_fast_kernel<E, T, false, false, 512, 2>(L, p, stream);
}

}  // namespace ofdm
