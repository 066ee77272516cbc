// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the Box-Muller transform used
// for the in-register AWGN.  Replaces numpy's Generator(PCG64).bytes (reference
// bits_generation/models.py:37) and np.random.normal (reference noise/models.py:19-21) in fused mode.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ofdm {

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                      uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
  const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
  const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  c0 = hi1 ^ c1 ^ k0;
  c1 = lo1;
  c2 = hi0 ^ c3 ^ k1;
  c3 = lo0;
}

template <int ROUNDS = 10>
__host__ __device__ __forceinline__ uint4 philox4x32(uint4 ctr, PhiloxKey key) {
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w, k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += W0;
    k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Two 32-bit words -> one circularly-symmetric complex standard normal pair (re, im each N(0,1)).
// The radius uses the full 32-bit uniform mapped to (x + 0.5) * 2^-32 so the tail reaches 6.6 sigma
// (SURVEY 7.4-2); the angle uses the second word.
__device__ __forceinline__ float2 box_muller(uint32_t wr, uint32_t wa) {
  const float u1 = fmaf((float)wr, 2.3283064365386963e-10f, 1.1641532182693481e-10f);  // (x+0.5)/2^32
  const float rad = sqrtf(-2.0f * __logf(u1));
  const float ang = (float)wa * 1.4629180792671596e-09f;  // 2*pi / 2^32
  float s, c;
  __sincosf(ang, &s, &c);
  return make_float2(rad * c, rad * s);
}

}  // namespace ofdm
