// K5 -- greedy CTC decode: row argmax, then collapse repeats / drop blanks.
//
// Replaces CTCDecoder::decode_greedy (src/decoders/ctc_decoder.cpp:443-490):
//   K5a  argmax over the alphabet per frame (ctc_decoder.cpp:451, torch.argmax semantics: the FIRST
//        maximum wins, a NaN is larger than every number and the first NaN wins); one warp per
//        row, the row fetched with 128-bit loads that are all in flight at once, warp-shuffle (key,index)
//        reduction, frames beyond the utterance length are not read;
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

// torch.argmax order as ONE unsigned key per value: larger key = better; NaN is the largest key; -0.0 and +0.0
// compare equal (x + 0 canonicalises the zero) so that ties between them fall to the lower index, as in torch.
__device__ __forceinline__ uint32_t order_key(float x) {
  if (x != x) return 0xffffffffu;
  const uint32_t u = __float_as_uint(x + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long order_key(double x) {
  if (x != x) return ~0ull;
  const unsigned long long u = (unsigned long long)__double_as_longlong(x + 0.0);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
template <typename K> struct KeyShfl;
template <> struct KeyShfl<uint32_t> {
  static __device__ __forceinline__ uint32_t xor_(uint32_t k, int o) { return __shfl_xor_sync(0xffffffffu, k, o); }
};
template <> struct KeyShfl<unsigned long long> {
  static __device__ __forceinline__ unsigned long long xor_(unsigned long long k, int o) { return __shfl_xor_sync(0xffffffffu, k, o); }
};

// 16-byte vector loads of the element type, widened to the compare type
template <typename T> struct GVec;
template <> struct GVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p)); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <> struct GVec<double> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void load(const double* p, double (&o)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p)); o[0] = v.x; o[1] = v.y;
  }
};
template <> struct GVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};
template <> struct GVec<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
};

// One warp per frame row.  NV > 0: the row (<= 32*N*NV symbols, 16-byte aligned) is fetched with NV 128-bit loads
// per lane, all in flight at once; NV == 0: scalar streaming loop (unaligned rows, huge alphabets).  A lane walks its
// symbols in increasing index order, so a strict > keeps the first maximum; across lanes the lower index breaks ties.
template <typename T, int NV>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
ctc_argmax_kernel(const T* __restrict__ logits, long long sb, long long st, int B, int T_, int V,
                  const void* in_len, int len_is64, int* __restrict__ sym) {
  using acc_t = typename Elem<T>::acc_t;
  using key_t = decltype(order_key(acc_t(0)));
  constexpr int N = GVec<T>::N;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= (long long)B * T_) return;
  const int b = (int)(row / T_), t = (int)(row % T_);
  if (in_len != nullptr && t >= load_index(in_len, len_is64, b)) return;
  const T* x = logits + b * sb + t * st;
  key_t bk = 0;              // below every real key (the smallest is ~(-inf bits) > 0 ... see below): index decides
  int bi = 0x7fffffff;
  if (NV > 0) {
    acc_t v[NV > 0 ? NV : 1][N];
#pragma unroll
    for (int j = 0; j < NV; j++) {
      const int i = (j * 32 + lane) * N;
      if (i < V) GVec<T>::load(x + i, v[j]);
    }
#pragma unroll
    for (int j = 0; j < NV; j++) {
      const int i = (j * 32 + lane) * N;
      if (i < V) {
#pragma unroll
        for (int k = 0; k < N; k++) {
          const key_t kk = order_key(v[j][k]);
          if (bi == 0x7fffffff || kk > bk) { bk = kk; bi = i + k; }
        }
      }
    }
  } else {
    for (int i = lane; i < V; i += 32) {
      const key_t kk = order_key(Elem<T>::load(x + i));
      if (bi == 0x7fffffff || kk > bk) { bk = kk; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const key_t ok = KeyShfl<key_t>::xor_(bk, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ok > bk || (ok == bk && oi < bi))) { bk = ok; bi = oi; }
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
  constexpr int N = GVec<T>::N;
  const long long rows = (long long)d.batch * d.max_frames;
  const unsigned grid = (unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock);
  const int V = d.alphabet;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(logits) % 16 == 0) && ((d.logits_stride_b * sizeof(T)) % 16 == 0) &&
                      ((d.logits_stride_t * sizeof(T)) % 16 == 0) && V % N == 0 && V <= 32 * N * 8;
  const int nv = vec_ok ? (V + 32 * N - 1) / (32 * N) : 0;
  KernelTimer timer(kKernelArgmax, s);
#define E2E_K5(NV) ctc_argmax_kernel<T, NV><<<grid, kRowsPerBlock * 32, 0, s>>>(reinterpret_cast<const T*>(logits), d.logits_stride_b, \
      d.logits_stride_t, d.batch, d.max_frames, V, in_len, d.lengths_itype == E2E_I64, sym)
  if (nv == 0) E2E_K5(0); else if (nv == 1) E2E_K5(1); else if (nv == 2) E2E_K5(2); else if (nv <= 4) E2E_K5(4); else E2E_K5(8);
#undef E2E_K5
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
