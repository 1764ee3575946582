// Lattice kernel, block rows per lane 6 / 8 / 10 (targets up to 639 labels).
#include "ctc_fused_impl.cuh"

namespace e2e {
int launch_fused_c(int gather, const void* fp, cudaStream_t s) {
  const FzParams& p = *reinterpret_cast<const FzParams*>(fp);
  return gather ? launch_fused_k<10, true>(p, s) : launch_fused_k<10, false>(p, s);
}
}  // namespace e2e
