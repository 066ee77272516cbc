// Dispatch of the fast-path kernel (link_fast.cuh); the kernels are instantiated in link_fast_inst.cu,
// one translation unit per team shape (nvcc -DOFDM_FAST_E=<E> -DOFDM_FAST_T=<T>) so that the build parallelises.
#include "link_fast.cuh"
#include "plan.h"

#include <cmath>
#include <vector>

namespace ofdm {

template <int E, int T>
int launch_fast_shape(const ofdm_link* L, const FastParams& p, bool dump, bool replay, bool adapt, bool sc, bool isi, bool psk, cudaStream_t stream);

template <int E, int T>
int launch_frames_shape(int sms, const FastParams& p, cudaStream_t stream);

int launch_fast_frames(int n_subcarriers, int sms, const FastParams& p, cudaStream_t stream) {
  switch (n_subcarriers) {
    case 64: return launch_frames_shape<8, 8>(sms, p, stream);
    case 128: return launch_frames_shape<8, 16>(sms, p, stream);
    case 256: return launch_frames_shape<16, 16>(sms, p, stream);
    case 512: return launch_frames_shape<16, 32>(sms, p, stream);
    case 1024: return launch_frames_shape<32, 32>(sms, p, stream);
    case 2048: return launch_frames_shape<32, 64>(sms, p, stream);
    case 4096: return launch_frames_shape<32, 128>(sms, p, stream);
    case 8192: return launch_frames_shape<32, 256>(sms, p, stream);
    default: return fail(OFDM_EUNSUPPORTED, "no fast plan for n_subcarriers=%d", n_subcarriers);
  }
}

// Twiddles of the fast transform: pass 2, exp(-2 pi i k r / E^2) for leg r = 1 .. E-1 of lane column k at
// [k * (E + 2) + r - 1] (padded rows, read as 128-bit pairs); then for teams wider than E the pass-3 base twiddles
// exp(-2 pi i j / N), j < N / (T/E).
std::vector<float2> build_fast_twiddles(int N) {
  const int E = fast_samples_per_lane(N), T = N / E, W = T / E, RS = E + 2;
  std::vector<float2> tw(size_t(E) * RS, make_float2(0.f, 0.f));
  // row k (lane column): the pairs {W^(k n), W^(k (n + E/2))}, n < E/2, W = exp(-2 pi i / E^2) - one 128-bit load per
  // first-stage butterfly of the second codelet (kOptFusedTwiddle in link_fast.cuh)
  for (int k = 0; k < E; ++k)
    for (int r = 0; r < E; ++r) {
      const double ang = -2.0 * M_PI * double(k * r) / double(E * E);
      tw[size_t(k) * RS + 2 * (r % (E / 2)) + r / (E / 2)] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  if (W > 1)
    for (int j = 0; j < N / W; ++j) {
      const double ang = -2.0 * M_PI * double(j) / double(N);
      tw.push_back(make_float2((float)std::cos(ang), (float)std::sin(ang)));
    }
  return tw;
}

bool fast_supports_n(int n) { return n >= 64 && n <= 8192 && (n & (n - 1)) == 0; }
// which combinations of the kernel's flags are instantiated for a transform size (link_fast_inst.cu)
bool fast_supports_combo(int n, bool adapt, bool sc, bool isi, bool psk) {
  if ((psk && adapt) || (adapt && sc)) return false;
  if (n >= 2048 && psk && (sc || isi)) return false;
  return true;
}
int fast_samples_per_lane(int n) { return n <= 128 ? 8 : n <= 512 ? 16 : 32; }

int launch_fast(const ofdm_link* L, const FastParams& p, bool dump, bool replay, bool adapt, bool sc, bool isi, bool psk, cudaStream_t stream) {
  switch (L->d.n_subcarriers) {
    case 64: return launch_fast_shape<8, 8>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 128: return launch_fast_shape<8, 16>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 256: return launch_fast_shape<16, 16>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 512: return launch_fast_shape<16, 32>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 1024: return launch_fast_shape<32, 32>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 2048: return launch_fast_shape<32, 64>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 4096: return launch_fast_shape<32, 128>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    case 8192: return launch_fast_shape<32, 256>(L, p, dump, replay, adapt, sc, isi, psk, stream);
    default: return fail(OFDM_EUNSUPPORTED, "no fast plan for n_subcarriers=%d", L->d.n_subcarriers);
  }
}

}  // namespace ofdm
