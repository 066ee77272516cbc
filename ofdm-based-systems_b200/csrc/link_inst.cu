// One translation unit per supported FFT size: nvcc -DOFDM_INST_N=<N> -c link_inst.cu
#include "link_kernel.cuh"
#include "plan.h"

#ifndef OFDM_INST_N
#error "compile with -DOFDM_INST_N=<number of subcarriers>"
#endif

namespace ofdm {

template <int N>
int configure_kernel(ofdm_link* L) {
  constexpr int E = elements_per_thread(N);
  using G = Geometry<N, E>;
  L->E = E;
  L->T = G::T;
  L->block = G::BLOCK;
  L->teams = G::TEAMS;
  L->smem = G::SMEM_BYTES;
  auto kern = ofdm_link_kernel<N, E>;
  static int occ = 0;   // per process: the attribute and the occupancy query cost ~0.1 ms each
  if (occ == 0) {
    if (G::SMEM_BYTES > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, G::BLOCK, G::SMEM_BYTES));
    if (occ <= 0) occ = 1;
  }
  L->occ = occ;
  return OFDM_OK;
}

template <int N>
int launch_kernel(const ofdm_link* L, const LinkParams& p, cudaStream_t stream) {
  constexpr int E = elements_per_thread(N);
  using G = Geometry<N, E>;
  const unsigned long long need = (p.sym_count + G::TEAMS - 1) / G::TEAMS;
  unsigned long long grid = (unsigned long long)L->sms * L->occ;
  if (need < grid) grid = need;
  if (grid == 0) return OFDM_OK;
  ofdm_link_kernel<N, E><<<(unsigned)grid, G::BLOCK, G::SMEM_BYTES, stream>>>(p);
  count_launch();
  CUDA_TRY(cudaGetLastError());
  return OFDM_OK;
}

template int configure_kernel<OFDM_INST_N>(ofdm_link*);
template int launch_kernel<OFDM_INST_N>(const ofdm_link*, const LinkParams&, cudaStream_t);

}  // namespace ofdm
