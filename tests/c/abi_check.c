/* Plain-C consumer of include/ofdm_b200.h: the header must compile as C (no C++-isms) and the library must link from
 * C.  Without a GPU the compute entry points fail with a message instead of falling back to anything. */
#include <stdio.h>
#include <string.h>

#include "ofdm_b200.h"

int main(void) {
  ofdm_link_desc desc;
  ofdm_frames_desc frames;
  ofdm_waterfill_desc wf;
  ofdm_link_result res;
  ofdm_link* link = NULL;
  double taps[2] = {1.0, 0.0}, h_eq[128];
  int32_t orders[64];
  int i, rc;
  memset(&desc, 0, sizeof desc);
  memset(&frames, 0, sizeof frames);
  memset(&wf, 0, sizeof wf);
  memset(&res, 0, sizeof res);
  if (ofdm_b200_abi_version() != OFDM_B200_ABI_VERSION) return 2;
  for (i = 0; i < 64; ++i) { h_eq[2 * i] = 1.0; h_eq[2 * i + 1] = 0.0; orders[i] = 4; }
  desc.n_subcarriers = 64; desc.prefix_type = OFDM_PREFIX_CYCLIC; desc.modulator = OFDM_MOD_OFDM;
  desc.equalizer = OFDM_EQ_ZF; desc.scheme = OFDM_SCHEME_QAM; desc.n_taps = 1; desc.device = -1;
  printf("devices=%d\n", ofdm_b200_device_count());
  rc = ofdm_link_create(&desc, taps, h_eq, orders, NULL, &link);
  if (rc == OFDM_OK) {
    rc = ofdm_link_run_fused(link, 10.0, 0.2236, 1u, 0u, 0u, 1000u, NULL, &res);
    printf("run rc=%d bits=%llu errors=%llu fast=%d\n", rc, (unsigned long long)res.bits, (unsigned long long)res.bit_errors,
           ofdm_link_uses_fast_kernel(link));
    ofdm_link_destroy(link);
    return rc == OFDM_OK && res.bits == 128000u ? 0 : 3;
  }
  printf("create rc=%d: %s\n", rc, ofdm_b200_last_error());
  desc.n_subcarriers = 100;                       /* argument errors are reported before any device work */
  rc = ofdm_link_create(&desc, taps, h_eq, orders, NULL, &link);
  return rc == OFDM_EUNSUPPORTED && strlen(ofdm_b200_last_error()) > 0 ? 0 : 4;
}
