// One translation unit per team width: nvcc -DOFDM_FAST_E=<8|16|32> -c link_fast_inst.cu
#include <cstdlib>

#include "link_fast.cuh"
#include "plan.h"

#ifndef OFDM_FAST_E
#error "compile with -DOFDM_FAST_E=<lanes per OFDM symbol>"
#endif

namespace ofdm {

template <int E, bool DUMP, bool REPLAY, int BLOCK = 512, int SYNC = 2>
static int launch_fast_kernel(const ofdm_link* L, const FastParams& p, cudaStream_t stream) {
  using G = FastGeometry<E, BLOCK>;
  auto kern = ofdm_link_fast_kernel<E, DUMP, true, REPLAY, BLOCK, SYNC>;
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

template <int E>
int launch_fast_width(const ofdm_link* L, const FastParams& p, bool dump, bool replay, cudaStream_t stream);

template <>
int launch_fast_width<OFDM_FAST_E>(const ofdm_link* L, const FastParams& p, bool dump, bool replay, cudaStream_t stream) {
  constexpr int E = OFDM_FAST_E;
  if (replay) {
    if (dump) return launch_fast_kernel<E, true, true>(L, p, stream);
#if OFDM_FAST_E == 32
    static const int rvariant = [] { const char* v = std::getenv("OFDM_B200_REPLAY_VARIANT"); return v ? std::atoi(v) : 0; }();
    if (rvariant == 1) return launch_fast_kernel<E, false, true, 512, 0>(L, p, stream);
#endif
    return launch_fast_kernel<E, false, true>(L, p, stream);
  }
  if (dump) return launch_fast_kernel<E, true, false>(L, p, stream);
#if OFDM_FAST_E == 32
  // OFDM_B200_FAST_VARIANT=4: no cross-warp barriers (the experiment behind profiles/: free-running warps lose
  // ~3 % to instruction-cache misses)
  static const int variant = [] { const char* v = std::getenv("OFDM_B200_FAST_VARIANT"); return v ? std::atoi(v) : 0; }();
  if (variant == 4) return launch_fast_kernel<E, false, false, 512, 0>(L, p, stream);
#endif
  return launch_fast_kernel<E, false, false>(L, p, stream);
}

}  // namespace ofdm
