// K2 dispatch: picks the (cells per lane, lattice warps) instantiation of the lattice kernel
// (ctc_lattice_impl.cuh; instantiated in ctc_lattice_nw{1,2,4}.cu so the variants compile in parallel).
#include <cstdlib>

#include "common.cuh"

namespace e2e {

int launch_lattice_nw1(int K, const void* lp, const LossPlan& p, cudaStream_t s);
int launch_lattice_nw2(int K, const void* lp, const LossPlan& p, cudaStream_t s);
int launch_lattice_nw4(int K, const void* lp, const LossPlan& p, cudaStream_t s);
int launch_lattice_variant(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                           const void* in_len, const void* tgt_len, void* losses, void* grads, double scale,
                           char* ws, cudaStream_t s);

int launch_lattice(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                   const void* in_len, const void* tgt_len, void* losses, void* grads, double scale, char* ws,
                   cudaStream_t s) {
  return launch_lattice_variant(d, p, logits, targets, in_len, tgt_len, losses, grads, scale, ws, s);
}

}  // namespace e2e
