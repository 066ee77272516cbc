// Dispatch of the fast-path kernel (link_fast.cuh); the kernels are instantiated in link_fast_inst.cu,
// one translation unit per team shape (nvcc -DOFDM_FAST_E=<E> -DOFDM_FAST_T=<T>) so that the build parallelises.
#include "link_fast.cuh"
#include "plan.h"

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
    default: return fail(OFDM_EUNSUPPORTED, "no fast plan for n_subcarriers=%d", n_subcarriers);
  }
}

bool fast_supports_n(int n) { return n >= 64 && n <= 4096 && (n & (n - 1)) == 0; }
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
    default: return fail(OFDM_EUNSUPPORTED, "no fast plan for n_subcarriers=%d", L->d.n_subcarriers);
  }
}

}  // namespace ofdm
