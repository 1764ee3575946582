// Lattice kernel instantiations with FOUR lattice warps per sweep (latency mode).
#include "ctc_lattice_impl.cuh"

namespace e2e {
int launch_lattice_nw4(int K, const void* lpv, const LossPlan& p, cudaStream_t s) {
  const LatticeParams& lp = *reinterpret_cast<const LatticeParams*>(lpv);
  switch (K) {
    case 2: return launch_k<2, 4>(lp, p, s);
    case 4: return launch_k<4, 4>(lp, p, s);
    case 8: return launch_k<8, 4>(lp, p, s);
    case 16: return launch_k<16, 4>(lp, p, s);
  }
  set_error("lattice: no 4-warp variant with %d cells per lane", K);
  return E2E_ERR_UNSUPPORTED;
}
}  // namespace e2e
