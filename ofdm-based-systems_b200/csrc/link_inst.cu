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
  // resident blocks per SM on the link's device (the caller has made it current); cached per (kernel, device)
  return blocks_per_sm(ofdm_link_kernel<N, E>, G::BLOCK, G::SMEM_BYTES, &L->occ);
}

template <int N>
int launch_kernel(const ofdm_link* L, const LinkParams& p, cudaStream_t stream) {
  constexpr int E = elements_per_thread(N);
  using G = Geometry<N, E>;
  const unsigned long long need = (p.sym_count - p.sym_lo + G::TEAMS - 1) / G::TEAMS;
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
