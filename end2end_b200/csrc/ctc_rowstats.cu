// K1 -- fused row log-softmax statistics.
//
// Replaces F.log_softmax(logits, dim=2) of the reference's module (modules/ctc_loss.py:37-40) and
// the exp(logits_2d) term of the engine (src/losses/ctc_loss.cpp:117) WITHOUT materialising the
// B*T*V log-probabilities: per row (b,t) it emits only {max_v x, log sum_v exp(x - max)} and the
// consumers (lattice emission staging, gradient kernel) rebuild lp = (x - max) - logsum in
// registers with the same fp32 operation order torch's CPU kernel uses.  One warp per row,
// warp-shuffle max / sum reductions, 128-bit loads when the row is 16-byte aligned.
//
// HBM traffic: reads B*T*V*sizeof(logit), writes 8 (16 for f64) bytes per row.
#include "common.cuh"

namespace e2e {
namespace {

constexpr int kRowsPerBlock = 8;

template <typename T, int VEC>
struct VecLoad;
template <>
struct VecLoad<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
};
template <>
struct VecLoad<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      o[2 * i] = __uint_as_float(w[i] << 16);
      o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct VecLoad<__half, 8> {
  static __device__ __forceinline__ void load(const __half* p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 f = __half22float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
};
template <>
struct VecLoad<double, 2> {
  static __device__ __forceinline__ void load(const double* p, double (&o)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    o[0] = v.x; o[1] = v.y;
  }
};

__device__ __forceinline__ float exp_acc(float x) { return expf(x); }
__device__ __forceinline__ double exp_acc(double x) { return exp(x); }
__device__ __forceinline__ float log_acc(float x) { return logf(x); }
__device__ __forceinline__ double log_acc(double x) { return log(x); }

template <typename T, int VEC>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
ctc_row_stats_kernel(const T* __restrict__ logits, long long stride_b, long long stride_t, int B,
                     int T_, int V, bool vec_ok, typename Elem<T>::acc_t* __restrict__ stats) {
  using acc_t = typename Elem<T>::acc_t;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= (long long)B * T_) return;
  const int b = (int)(row / T_), t = (int)(row % T_);
  const T* x = logits + b * stride_b + t * stride_t;

  // pass 1: max (a NaN in the row poisons the statistics, as it does in torch)
  acc_t m = -INFINITY;
  bool has_nan = false;
  const int nvec = vec_ok ? V / VEC : 0;
  for (int i = lane; i < nvec; i += 32) {
    acc_t v[VEC];
    VecLoad<T, VEC>::load(x + (long long)i * VEC, v);
#pragma unroll
    for (int k = 0; k < VEC; k++) { m = v[k] > m ? v[k] : m; has_nan |= (v[k] != v[k]); }
  }
  for (int i = nvec * VEC + lane; i < V; i += 32) {
    const acc_t v = Elem<T>::load(x + i);
    m = v > m ? v : m; has_nan |= (v != v);
  }
  m = warp_max(m);
  // pass 2: sum exp(x - max); the row is L1/L2 resident from pass 1
  acc_t s = 0;
  for (int i = lane; i < nvec; i += 32) {
    acc_t v[VEC];
    VecLoad<T, VEC>::load(x + (long long)i * VEC, v);
#pragma unroll
    for (int k = 0; k < VEC; k++) s += exp_acc(v[k] - m);
  }
  for (int i = nvec * VEC + lane; i < V; i += 32) s += exp_acc(Elem<T>::load(x + i) - m);
  s = warp_sum(s);
  has_nan = __any_sync(0xffffffffu, has_nan);
  if (lane == 0) {
    acc_t ls = log_acc(s);
    if (has_nan) { m = NAN; ls = NAN; }
    stats[2 * row] = m;
    stats[2 * row + 1] = ls;
  }
}

template <typename T, int VEC>
int launch_typed(const e2e_ctc_desc& d, const void* logits, void* stats, cudaStream_t s) {
  const long long rows = (long long)d.batch * d.max_frames;
  const unsigned grid = (unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock);
  const size_t vb = sizeof(T) * VEC;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(logits) % vb == 0) &&
                      ((d.logits_stride_b * sizeof(T)) % vb == 0) &&
                      ((d.logits_stride_t * sizeof(T)) % vb == 0) && d.alphabet >= VEC * 8;
  KernelTimer timer(kKernelRowStats, s);
  ctc_row_stats_kernel<T, VEC><<<grid, kRowsPerBlock * 32, 0, s>>>(
      reinterpret_cast<const T*>(logits), d.logits_stride_b, d.logits_stride_t, d.batch,
      d.max_frames, d.alphabet, vec_ok, reinterpret_cast<typename Elem<T>::acc_t*>(stats));
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace

int launch_row_stats(const e2e_ctc_desc& d, const void* logits, void* stats, cudaStream_t s) {
  switch (d.dtype) {
    case E2E_F32: return launch_typed<float, 4>(d, logits, stats, s);
    case E2E_BF16: return launch_typed<__nv_bfloat16, 8>(d, logits, stats, s);
    case E2E_F16: return launch_typed<__half, 8>(d, logits, stats, s);
    case E2E_F64: return launch_typed<double, 2>(d, logits, stats, s);
  }
  set_error("row_stats: unsupported dtype %d", d.dtype);
  return E2E_ERR_INVALID_ARGUMENT;
}

}  // namespace e2e
