// K5 -- greedy CTC decode: row argmax, then collapse repeats / drop blanks.
//
// Replaces CTCDecoder::decode_greedy (src/decoders/ctc_decoder.cpp:443-490):
//   K5a  argmax over the alphabet per frame (ctc_decoder.cpp:451, torch.argmax semantics: the FIRST
//        maximum wins, a NaN is larger than every number and the first NaN wins); one warp per
//        row, warp-shuffle (value,index) reduction, frames beyond the utterance length are not read;
//   K5b  per utterance, emit frame t's symbol iff it is not blank and differs from frame t-1's
//        (ctc_decoder.cpp:471-482) -- a warp-ballot / popc compaction scan over the frames -- into
//        the zero-padded [B,T] int64 output, plus the decoded length.
// Results are integers: parity with the reference is bit-exact.
//
// HBM traffic: reads B*T*V*sizeof(logit) once; writes B*T*8 + B*8 bytes.
#include "common.cuh"

namespace e2e {
namespace {

constexpr int kRowsPerBlock = 8;

template <typename A>
__device__ __forceinline__ bool better(A va, int ia, A vb, int ib) {
  // true when (va, ia) beats (vb, ib)
  const bool na = va != va, nb = vb != vb;
  if (na || nb) return na && (!nb || ia < ib);
  return va > vb || (va == vb && ia < ib);
}

template <typename T>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
ctc_argmax_kernel(const T* __restrict__ logits, long long sb, long long st, int B, int T_, int V,
                  const void* in_len, int len_is64, int* __restrict__ sym) {
  using acc_t = typename Elem<T>::acc_t;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= (long long)B * T_) return;
  const int b = (int)(row / T_), t = (int)(row % T_);
  if (in_len != nullptr && t >= load_index(in_len, len_is64, b)) return;
  const T* x = logits + b * sb + t * st;
  acc_t bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const acc_t xv = Elem<T>::load(x + v);
    if (bi == 0x7fffffff || better(xv, v, bv, bi)) { bv = xv; bi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const acc_t ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
  }
  if (lane == 0) sym[row] = bi;
}

__global__ void __launch_bounds__(256)
ctc_collapse_kernel(const int* __restrict__ sym, int T_, int blank, const void* in_len, int len_is64,
                    long long* __restrict__ decoded, long long* __restrict__ decoded_len) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  long long Ti_ll = in_len ? load_index(in_len, len_is64, b) : (long long)T_;
  const int Ti = (int)max(0LL, min(Ti_ll, (long long)T_));
  const int* s = sym + (long long)b * T_;
  long long* out = decoded + (long long)b * T_;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < Ti; t0 += 256) {
    const int t = t0 + tid;
    int cur = blank;
    bool keep = false;
    if (t < Ti) {
      cur = s[t];
      const int prev = t > 0 ? s[t - 1] : blank;
      keep = cur != blank && cur != prev;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[w] = __popc(mask);
    __syncthreads();
    int off = s_base;
    for (int q = 0; q < w; q++) off += s_warp[q];
    if (keep) out[off + __popc(mask & ((1u << lane) - 1u))] = cur;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < 8; q++) tot += s_warp[q];
      s_base += tot;
    }
    __syncthreads();
  }
  const int n = s_base;
  for (int t = n + tid; t < T_; t += 256) out[t] = 0;  // zero padding (at::zeros_like, ctc_decoder.cpp:452)
  if (tid == 0) decoded_len[b] = n;
}

template <typename T>
int launch_typed(const e2e_ctc_desc& d, const void* logits, const void* in_len, int* sym, cudaStream_t s) {
  const long long rows = (long long)d.batch * d.max_frames;
  const unsigned grid = (unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock);
  KernelTimer timer(kKernelArgmax, s);
  ctc_argmax_kernel<T><<<grid, kRowsPerBlock * 32, 0, s>>>(
      reinterpret_cast<const T*>(logits), d.logits_stride_b, d.logits_stride_t, d.batch, d.max_frames,
      d.alphabet, in_len, d.lengths_itype == E2E_I64, sym);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace

int launch_greedy(const e2e_ctc_desc& d, const void* logits, const void* in_len, int64_t* decoded,
                  int64_t* decoded_len, char* ws, cudaStream_t s) {
  int* sym = reinterpret_cast<int*>(ws);
  int rc;
  switch (d.dtype) {
    case E2E_F32: rc = launch_typed<float>(d, logits, in_len, sym, s); break;
    case E2E_BF16: rc = launch_typed<__nv_bfloat16>(d, logits, in_len, sym, s); break;
    case E2E_F16: rc = launch_typed<__half>(d, logits, in_len, sym, s); break;
    case E2E_F64: rc = launch_typed<double>(d, logits, in_len, sym, s); break;
    default: set_error("greedy: unsupported dtype %d", d.dtype); return E2E_ERR_INVALID_ARGUMENT;
  }
  if (rc != E2E_OK) return rc;
  KernelTimer timer(kKernelCollapse, s);
  ctc_collapse_kernel<<<(unsigned)d.batch, 256, 0, s>>>(
      sym, d.max_frames, d.blank_idx, in_len, d.lengths_itype == E2E_I64,
      reinterpret_cast<long long*>(decoded), reinterpret_cast<long long*>(decoded_len));
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace e2e
