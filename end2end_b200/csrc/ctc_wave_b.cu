// Wave-kernel instantiations with four and eight lattice warps per sweep.
#include "ctc_wave_impl.cuh"

namespace e2e {

int launch_wave_b(int K, int NW, const void* wpv, cudaStream_t s) {
  const WaveParams& wp = *reinterpret_cast<const WaveParams*>(wpv);
  if (K == 4 && NW == 4) return launch_wave_k<4, 4>(wp, s);
  if (K == 4 && NW == 8) return launch_wave_k<4, 8>(wp, s);
  if (K == 8 && NW == 8) return launch_wave_k<8, 8>(wp, s);
  set_error("wave: no variant with %d cells per lane x %d lattice warps", K, NW);
  return E2E_ERR_UNSUPPORTED;
}

}  // namespace e2e
