// Lattice kernel instantiations with ONE lattice warp per sweep, the parameter block and the trace hooks.
#include "ctc_lattice_impl.cuh"

namespace e2e {

int launch_lattice_nw2(int K, const void* lp, const LossPlan& p, cudaStream_t s);
int launch_lattice_nw4(int K, const void* lp, const LossPlan& p, cudaStream_t s);

// E2E_CTC_TRACE=1: a device buffer the kernel stamps per-chunk clocks into (debugging aid)
static long long* g_trace = nullptr;
static long long* lattice_trace_buffer() {
  static int on = -1;
  if (on < 0) { const char* v = getenv("E2E_CTC_TRACE"); on = (v && *v == '1') ? 1 : 0; }
  if (on && !g_trace) {
    const size_t n = (size_t)2 * 2 * 3 * kTraceChunks * kTraceEvents * sizeof(long long);
    if (cudaMalloc(&g_trace, n) != cudaSuccess) g_trace = nullptr; else cudaMemset(g_trace, 0, n);
  }
  return on ? g_trace : nullptr;
}
int lattice_trace_read(long long* host, size_t n) {
  if (!g_trace) return 0;
  cudaDeviceSynchronize();
  const size_t have = (size_t)2 * 2 * 3 * kTraceChunks * kTraceEvents;
  if (n > have) n = have;
  cudaMemcpy(host, g_trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
  return (int)n;
}

int launch_lattice_nw1(int K, const void* lpv, const LossPlan& p, cudaStream_t s) {
  const LatticeParams& lp = *reinterpret_cast<const LatticeParams*>(lpv);
  switch (K) {
    case 2: return launch_k<2, 1>(lp, p, s);
    case 4: return launch_k<4, 1>(lp, p, s);
    case 8: return launch_k<8, 1>(lp, p, s);
    case 16: return launch_k<16, 1>(lp, p, s);
    case 24: return launch_k<24, 1>(lp, p, s);
    case 40: return launch_k<40, 1>(lp, p, s);
  }
  set_error("lattice: no 1-warp variant with %d cells per lane", K);
  return E2E_ERR_UNSUPPORTED;
}

int launch_lattice_variant(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                           const void* in_len, const void* tgt_len, void* losses, void* grads, double scale,
                           char* ws, cudaStream_t s) {
  LatticeParams lp;
  lp.logits = logits; lp.dtype = d.dtype; lp.sb = d.logits_stride_b; lp.st = d.logits_stride_t;
  lp.grads = grads; lp.gsb = d.grads_stride_b; lp.gst = d.grads_stride_t; lp.scale = scale;
  lp.stats = ws + p.off_stats;
  lp.targets = targets; lp.tgt_is64 = d.targets_itype == E2E_I64; lp.ts_b = d.targets_stride_b;
  lp.in_len = in_len; lp.tgt_len = tgt_len; lp.len_is64 = d.lengths_itype == E2E_I64;
  lp.B = d.batch; lp.T = d.max_frames; lp.V = d.alphabet; lp.Lmax = d.max_targets;
  lp.blank = d.blank_idx; lp.from_logits = d.from_logits;
  lp.losses = losses;
  lp.status = reinterpret_cast<int*>(ws + p.off_status);
  lp.flags = reinterpret_cast<int*>(ws + p.off_flags);
  lp.stash = reinterpret_cast<uint32_t*>(ws + p.off_stash);
  lp.post = reinterpret_cast<float*>(ws + p.off_post);
  lp.np = p.np; lp.nc = p.nc; lp.pfd = p.pfd; lp.sm = p.sm; lp.trace = lattice_trace_buffer();
  lp.chunk_log2 = p.chunk_log2; lp.lstride = p.lstride; lp.dense = p.dense; lp.rowlen_max = p.rowlen;
  lp.post_stride = p.post_stride; lp.vpad = p.vpad;
  if (p.dense && grads == nullptr) { set_error("lattice: fused mode needs a gradient buffer"); return E2E_ERR_INVALID_ARGUMENT; }
  switch (p.NW) {
    case 1: return launch_lattice_nw1(p.K, &lp, p, s);
    case 2: return launch_lattice_nw2(p.K, &lp, p, s);
    case 4: return launch_lattice_nw4(p.K, &lp, p, s);
  }
  set_error("lattice: unsupported shape (cells per lane %d, warps %d)", p.K, p.NW);
  return E2E_ERR_UNSUPPORTED;
}

}  // namespace e2e
