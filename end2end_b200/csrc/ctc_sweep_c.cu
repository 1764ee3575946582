// Sweep-kernel instantiations: 20..28 cells per lane.
#include "ctc_sweep_impl.cuh"

namespace e2e {
int launch_sweep_c(int K, bool f64, const void* spv, size_t smem, cudaStream_t s) {
  const SweepParams& sp = *reinterpret_cast<const SweepParams*>(spv);
  if (f64) {
    if (K == 24) return launch_sweep_k<24, true>(sp, smem, s);
  } else {
    switch (K) {
      case 20: return launch_sweep_k<20, false>(sp, smem, s);
      case 24: return launch_sweep_k<24, false>(sp, smem, s);
      case 28: return launch_sweep_k<28, false>(sp, smem, s);
    }
  }
  set_error("sweep: no variant with %d cells per lane (f64=%d)", K, (int)f64);
  return E2E_ERR_UNSUPPORTED;
}
}  // namespace e2e
