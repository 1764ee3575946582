// Sweep-kernel instantiations: 10..16 cells per lane.
#include "ctc_sweep_impl.cuh"

namespace e2e {
int launch_sweep_b(int K, bool f64, const void* spv, size_t smem, cudaStream_t s) {
  const SweepParams& sp = *reinterpret_cast<const SweepParams*>(spv);
  if (f64) {
    if (K == 16) return launch_sweep_k<16, true>(sp, smem, s);
  } else {
    switch (K) {
      case 10: return launch_sweep_k<10, false>(sp, smem, s);
      case 12: return launch_sweep_k<12, false>(sp, smem, s);
      case 14: return launch_sweep_k<14, false>(sp, smem, s);
      case 16: return launch_sweep_k<16, false>(sp, smem, s);
    }
  }
  set_error("sweep: no variant with %d cells per lane (f64=%d)", K, (int)f64);
  return E2E_ERR_UNSUPPORTED;
}
}  // namespace e2e
