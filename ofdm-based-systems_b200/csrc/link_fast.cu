// Instantiations + launch glue of the fast-path kernel (link_fast.cuh).
#include "link_fast.cuh"
#include "plan.h"

#include <cstdlib>

namespace ofdm {

template <int E, bool DUMP, int BLOCK = 512, int SYNC = 2, int NROUNDS = 10, int FIR_UNROLL = 2>
static int launch_fast_e(const ofdm_link* L, const FastParams& p, cudaStream_t stream) {
  using G = FastGeometry<E, BLOCK>;
  auto kern = ofdm_link_fast_kernel<E, DUMP, true, BLOCK, SYNC, NROUNDS, FIR_UNROLL>;
  static int occ = 0;
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

bool fast_supports_n(int n) { return n == 64 || n == 256 || n == 1024; }

int launch_fast(const ofdm_link* L, const FastParams& p, bool dump, cudaStream_t stream) {
  switch (L->d.n_subcarriers) {
    case 64: return dump ? launch_fast_e<8, true>(L, p, stream) : launch_fast_e<8, false>(L, p, stream);
    case 256: return dump ? launch_fast_e<16, true>(L, p, stream) : launch_fast_e<16, false>(L, p, stream);
    case 1024: {
      if (dump) return launch_fast_e<32, true>(L, p, stream);
      static const int variant = [] { const char* v = std::getenv("OFDM_B200_FAST_VARIANT"); return v ? std::atoi(v) : 0; }();
      switch (variant) {
        case 1: return launch_fast_e<32, false, 512, 1>(L, p, stream);
        case 3: return launch_fast_e<32, false, 512, 3>(L, p, stream);
        case 4: return launch_fast_e<32, false, 512, 0>(L, p, stream);
        case 5: return launch_fast_e<32, false, 256, 2>(L, p, stream);
        case 6: return launch_fast_e<32, false, 512, 2, 7, 2>(L, p, stream);
        case 7: return launch_fast_e<32, false, 512, 2, 10, 1>(L, p, stream);
        case 8: return launch_fast_e<32, false, 640, 0>(L, p, stream);
        case 9: return launch_fast_e<32, false, 512, 0, 10, 4>(L, p, stream);
        default: return launch_fast_e<32, false>(L, p, stream);
      }
    }
    default: return fail(OFDM_EUNSUPPORTED, "no fast plan for n_subcarriers=%d", L->d.n_subcarriers);
  }
}

}  // namespace ofdm
