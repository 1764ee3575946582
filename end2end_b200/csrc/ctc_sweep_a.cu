// Sweep-kernel dispatch + the narrow instantiations (2..8 cells per lane).
#include "ctc_sweep_impl.cuh"

namespace e2e {

int launch_sweep_b(int K, bool f64, const void* sp, size_t smem, cudaStream_t s);
int launch_sweep_c(int K, bool f64, const void* sp, size_t smem, cudaStream_t s);
int launch_sweep_d(int K, bool f64, const void* sp, size_t smem, cudaStream_t s);

static int launch_sweep_a(int K, bool f64, const SweepParams& sp, size_t smem, cudaStream_t s) {
  if (f64) {
    switch (K) {
      case 2: return launch_sweep_k<2, true>(sp, smem, s);
      case 4: return launch_sweep_k<4, true>(sp, smem, s);
      case 8: return launch_sweep_k<8, true>(sp, smem, s);
    }
  } else {
    switch (K) {
      case 2: return launch_sweep_k<2, false>(sp, smem, s);
      case 4: return launch_sweep_k<4, false>(sp, smem, s);
      case 6: return launch_sweep_k<6, false>(sp, smem, s);
      case 8: return launch_sweep_k<8, false>(sp, smem, s);
    }
  }
  set_error("sweep: no variant with %d cells per lane (f64=%d)", K, (int)f64);
  return E2E_ERR_UNSUPPORTED;
}

// Utterance indices by falling frame count (ties by index): order[rank] = b.  One thread per utterance counts the
// utterances that go before it; lengths outside [1, T] (rejected later by the lattice kernel) sort as 0.
__global__ void __launch_bounds__(256) ctc_order_kernel(const void* in_len, int is64, int B, int T, int* __restrict__ order) {
  __shared__ int tile[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  auto key = [&](int j) { const long long v = load_index(in_len, is64, j); return (v < 1 || v > T) ? 0 : (int)v; };
  const int mine = i < B ? key(i) : 0;
  int rank = 0;
  for (int j0 = 0; j0 < B; j0 += 256) {
    __syncthreads();
    tile[threadIdx.x] = j0 + threadIdx.x < B ? key(j0 + threadIdx.x) : -1;
    __syncthreads();
    const int n = min(256, B - j0);
    for (int q = 0; q < n; q++) { const int k = tile[q]; rank += (k > mine) || (k == mine && j0 + q < i); }
  }
  if (i < B) order[rank] = i;
}

int launch_sweep(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                 const void* in_len, const void* tgt_len, void* losses, void* grads, double scale, char* ws,
                 cudaStream_t s) {
  SweepParams sp;
  sp.logits = logits; sp.dtype = d.dtype; sp.sb = d.logits_stride_b; sp.st = d.logits_stride_t;
  sp.grads = grads; sp.gsb = d.grads_stride_b; sp.gst = d.grads_stride_t; sp.scale = scale;
  sp.stats = ws + p.off_stats;
  sp.targets = targets; sp.tgt_is64 = d.targets_itype == E2E_I64; sp.ts_b = d.targets_stride_b;
  sp.in_len = in_len; sp.tgt_len = tgt_len; sp.len_is64 = d.lengths_itype == E2E_I64;
  sp.B = d.batch; sp.T = d.max_frames; sp.V = d.alphabet; sp.Lmax = d.max_targets;
  sp.blank = d.blank_idx; sp.from_logits = d.from_logits;
  sp.losses = losses;
  sp.status = reinterpret_cast<int*>(ws + p.off_status);
  sp.flags = reinterpret_cast<int*>(ws + p.off_flags);
  sp.stash = reinterpret_cast<uint32_t*>(ws + p.off_stash);
  sp.order = nullptr;
  if (p.off_order) {
    int* order = reinterpret_cast<int*>(ws + p.off_order);
    KernelTimer timer(kKernelOrder, s);
    ctc_order_kernel<<<(unsigned)((d.batch + 255) / 256), 256, 0, s>>>(in_len, d.lengths_itype == E2E_I64, d.batch, d.max_frames, order);
    E2E_CUDA_TRY(cudaGetLastError());
    sp.order = order;
  }
  sp.post = reinterpret_cast<float*>(ws + p.off_post);
  sp.dense = p.dense; sp.post_stride = p.post_stride; sp.cells = p.cells;
  sp.cf = p.sw.cf; sp.es = p.sw.es; sp.rawrow = p.sw.rawrow; sp.vpad = p.sw.vpad;
  sp.off_lab = p.sw.off_lab; sp.off_warp = p.sw.off_warp; sp.warp_bytes = p.sw.warp_bytes;
  sp.w_E = p.sw.w_E; sp.w_raw = p.sw.w_raw; sp.w_stat = p.sw.w_stat; sp.w_rs = p.sw.w_rs;
  sp.w_acc = p.sw.w_acc; sp.w_stage = p.sw.w_stage;
  if (p.dense && grads == nullptr) { set_error("sweep: fused mode needs a gradient buffer"); return E2E_ERR_INVALID_ARGUMENT; }
  const bool f64 = d.dtype == E2E_F64;
  if (p.K <= 8) return launch_sweep_a(p.K, f64, sp, p.smem, s);
  if (p.K <= 16) return launch_sweep_b(p.K, f64, &sp, p.smem, s);
  if (p.K <= 28) return launch_sweep_c(p.K, f64, &sp, p.smem, s);
  return launch_sweep_d(p.K, f64, &sp, p.smem, s);
}

}  // namespace e2e
