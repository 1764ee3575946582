// Lattice kernel, block rows per lane <= 4 (targets up to 255 labels).
#include "ctc_fused_impl.cuh"

namespace e2e {
int launch_fused_b(int gather, const void* fp, cudaStream_t s) {
  const FzParams& p = *reinterpret_cast<const FzParams*>(fp);
  return gather ? launch_fused_k<4, true>(p, s) : launch_fused_k<4, false>(p, s);
}
}  // namespace e2e
