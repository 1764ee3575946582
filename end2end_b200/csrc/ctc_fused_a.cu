// Lattice-kernel dispatch + the short-target instantiations (block rows per lane: 1, 2).
#include "ctc_fused_impl.cuh"

namespace e2e {

int launch_fused_b(int gather, const void* fp, cudaStream_t s);   // NB = 4
int launch_fused_c(int gather, const void* fp, cudaStream_t s);   // NB = 10

size_t fused_ctl_bytes() { return (sizeof(FzCtl) + 15) & ~(size_t)15; }

int launch_fused(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                 const void* in_len, const void* tgt_len, void* losses, void* grads, double scale, char* ws,
                 cudaStream_t s) {
  if (p.dense && grads == nullptr) { set_error("lattice: fused mode needs a gradient buffer"); return E2E_ERR_INVALID_ARGUMENT; }
  FzParams fp;
  fp.logits = logits; fp.dtype = d.dtype; fp.sb = d.logits_stride_b; fp.st = d.logits_stride_t;
  fp.grads = grads; fp.gsb = d.grads_stride_b; fp.gst = d.grads_stride_t; fp.scale = scale;
  fp.stats = ws + p.off_stats;
  fp.emis = p.emis_stride > 0 ? reinterpret_cast<const double*>(ws + p.off_emis) : nullptr; fp.emis_stride = p.emis_stride;
  fp.post = reinterpret_cast<float*>(ws + p.off_post); fp.post_stride = p.post_stride; fp.cells = p.cells;
  fp.targets = targets; fp.tgt_is64 = d.targets_itype == E2E_I64; fp.ts_b = d.targets_stride_b;
  fp.in_len = in_len; fp.tgt_len = tgt_len; fp.len_is64 = d.lengths_itype == E2E_I64;
  fp.B = d.batch; fp.T = d.max_frames; fp.V = d.alphabet; fp.Lmax = d.max_targets;
  fp.blank = d.blank_idx; fp.from_logits = d.from_logits;
  fp.losses = losses;
  fp.status = reinterpret_cast<int*>(ws + p.off_status);
  fp.flags = reinterpret_cast<int*>(ws + p.off_flags);
  fp.meet = reinterpret_cast<int*>(ws + p.off_meet);
  fp.stash = reinterpret_cast<uint32_t*>(ws + p.off_stash);
  fp.roww = p.roww;
  fp.L = p.fz;
  const int gather = p.fz.gather;
  switch (p.fz.NB) {
    case 1: return gather ? launch_fused_k<1, true>(fp, s) : launch_fused_k<1, false>(fp, s);
    case 2: return gather ? launch_fused_k<2, true>(fp, s) : launch_fused_k<2, false>(fp, s);
    case 4: return launch_fused_b(gather, &fp, s);
    case 10: return launch_fused_c(gather, &fp, s);
  }
  set_error("lattice: no kernel class for %d block rows", p.fz.NB);
  return E2E_ERR_UNSUPPORTED;
}

}  // namespace e2e
