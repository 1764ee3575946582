// Lattice kernel instantiations with TWO lattice warps per sweep.
#include "ctc_lattice_impl.cuh"

namespace e2e {
int launch_lattice_nw2(int K, const void* lpv, const LossPlan& p, cudaStream_t s) {
  const LatticeParams& lp = *reinterpret_cast<const LatticeParams*>(lpv);
  switch (K) {
    case 2: return launch_k<2, 2>(lp, p, s);
    case 40: return launch_k<40, 2>(lp, p, s);
  }
  set_error("lattice: no 2-warp variant with %d cells per lane", K);
  return E2E_ERR_UNSUPPORTED;
}
}  // namespace e2e
