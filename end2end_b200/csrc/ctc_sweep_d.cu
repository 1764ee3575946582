// Sweep-kernel instantiations: 32..40 cells per lane.
#include "ctc_sweep_impl.cuh"

namespace e2e {
int launch_sweep_d(int K, bool f64, const void* spv, size_t smem, cudaStream_t s) {
  const SweepParams& sp = *reinterpret_cast<const SweepParams*>(spv);
  if (f64) {
    if (K == 40) return launch_sweep_k<40, true>(sp, smem, s);
  } else {
    switch (K) {
      case 32: return launch_sweep_k<32, false>(sp, smem, s);
      case 36: return launch_sweep_k<36, false>(sp, smem, s);
      case 40: return launch_sweep_k<40, false>(sp, smem, s);
    }
  }
  set_error("sweep: no variant with %d cells per lane (f64=%d)", K, (int)f64);
  return E2E_ERR_UNSUPPORTED;
}
}  // namespace e2e
