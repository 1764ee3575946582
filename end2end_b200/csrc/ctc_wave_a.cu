// Wave-kernel dispatch + the one/two-lattice-warp instantiations.
#include <cstdio>

#include "ctc_wave_impl.cuh"

namespace e2e {

int launch_wave_b(int K, int NW, const void* wp, cudaStream_t s);

size_t wave_ctl_bytes() { return (sizeof(WaveCtl) + 15) & ~(size_t)15; }

int launch_wave(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                const void* in_len, const void* tgt_len, void* losses, void* grads, double scale, char* ws,
                cudaStream_t s) {
  if (grads == nullptr) { set_error("wave: fused mode needs a gradient buffer"); return E2E_ERR_INVALID_ARGUMENT; }
  WaveParams wp;
  wp.logits = logits; wp.dtype = d.dtype; wp.sb = d.logits_stride_b; wp.st = d.logits_stride_t;
  wp.grads = grads; wp.gsb = d.grads_stride_b; wp.gst = d.grads_stride_t; wp.scale = scale;
  wp.targets = targets; wp.tgt_is64 = d.targets_itype == E2E_I64; wp.ts_b = d.targets_stride_b;
  wp.in_len = in_len; wp.tgt_len = tgt_len; wp.len_is64 = d.lengths_itype == E2E_I64;
  wp.B = d.batch; wp.T = d.max_frames; wp.V = d.alphabet; wp.Lmax = d.max_targets;
  wp.blank = d.blank_idx; wp.from_logits = d.from_logits;
  wp.losses = losses;
  wp.status = reinterpret_cast<int*>(ws + p.off_status);
  wp.flags = reinterpret_cast<int*>(ws + p.off_flags);
  wp.meet = reinterpret_cast<int*>(ws + p.off_meet);
  wp.stash = reinterpret_cast<uint32_t*>(ws + p.off_stash);
  wp.L = p.wv;
  // E2E_CTC_WAVE_DBG=1: per-warp cycle counters of utterance 0, printed after a blocking sync (debugging aid)
  static int dbg_on = -1;
  static long long* dbg_buf = nullptr;
  if (dbg_on < 0) { const char* v = getenv("E2E_CTC_WAVE_DBG"); dbg_on = (v && *v == '1') ? 1 : 0; }
  if (dbg_on && !dbg_buf && cudaMalloc(&dbg_buf, 2 * 16 * 8 * sizeof(long long)) != cudaSuccess) dbg_buf = nullptr;
  wp.dbg = dbg_on ? dbg_buf : nullptr;
  if (wp.dbg) cudaMemsetAsync(wp.dbg, 0, 2 * 16 * 8 * sizeof(long long), s);
  const int K = p.wv.K, NW = p.wv.NW;
  int rc;
  if (K == 4 && NW == 1) rc = launch_wave_k<4, 1>(wp, s);
  else if (K == 4 && NW == 2) rc = launch_wave_k<4, 2>(wp, s);
  else rc = launch_wave_b(K, NW, &wp, s);
  if (wp.dbg && rc == E2E_OK) {
    long long h[2 * 16 * 8];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, wp.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const int nwarps = NW + p.wv.NC + p.wv.NP;
    for (int c = 0; c < 2; c++)
      for (int w = 0; w < nwarps && w < 16; w++) {
        const long long* r = h + (c * 16 + w) * 8;
        const char* role = w < p.wv.NP ? "producer" : (w < p.wv.NP + p.wv.NC ? "combiner" : "lattice ");
        fprintf(stderr, "[wave dbg] cta %d warp %2d %s total %8lld | %8lld %8lld %8lld %8lld %8lld\n", c, w, role, r[0], r[1], r[2], r[3], r[4], r[5]);
      }
  }
  return rc;
}

}  // namespace e2e
