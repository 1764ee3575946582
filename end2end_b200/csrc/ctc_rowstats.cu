// K1 -- fused row log-softmax statistics.
//
// Replaces F.log_softmax(logits, dim=2) of the reference's module (modules/ctc_loss.py:37-40) and
// the exp(logits_2d) term of the engine (src/losses/ctc_loss.cpp:117) WITHOUT materialising the
// B*T*V log-probabilities: per row (b,t) it emits only {max_v x, log sum_v exp(x - max)} and the
// consumers (lattice emission staging, gradient kernel) rebuild lp = (x - max) - logsum in
// registers with the same fp32 operation order torch's CPU kernel uses.  One warp per row,
// warp-shuffle max / sum reductions, 128-bit loads when the row is 16-byte aligned.
//
// HBM traffic: reads B*T*V*sizeof(logit), writes 8 (16 for f64) bytes per row (+ 8*(Lmax+1) bytes per row of
// compact emissions for the general lattice kernel's large-alphabet mode).
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

// Optional second output: the emissions the lattice needs, COMPACT per utterance (large alphabets).  The warp
// has just streamed its row, so the <= L+1 values p(t, label_k), p(t, blank) are L1/L2 hits here, while the
// lattice kernel's producer warps would each pay a dependent HBM round trip per gathered block (measured on
// BASELINE config 4: the lattice warp spent 65 % of its cycles waiting for gathered emissions).  Row layout:
// [label 0 .. label Lmax-1 | blank], doubles, exactly the values the lattice kernel's own gather would produce.
struct EmisArgs {
  double* emis; int stride;                 // nullptr: statistics only
  const void* targets; int tgt_is64; long long ts_b;
  const void* in_len; const void* tgt_len; int len_is64;
  int Lmax, blank, from_logits;
};

// p(t, v) relative to the row's log-sum-exp -- the same arithmetic as the lattice kernels' emission staging
// (ctc_fused_impl.cuh fz_emission): raw logits: exp of torch's fp32 log_softmax argument; log-prob input: the
// argument x - (m + ls) split into an fp32 head and tail, -inf (a masked symbol) an exact zero.
__device__ __forceinline__ double k1_emission(float x, float m, float ls, int from_logits) {
  if (from_logits) return (double)expf((x - m) - ls);
  const double d = (double)x - ((double)m + (double)ls);
  const float hi = (float)d;
  if (!(hi > -INFINITY)) return hi != hi ? (double)hi : 0.0;
  const float lo = (float)(d - (double)hi);
  return (double)expf(hi) * (1.0 + (double)lo);
}
__device__ __forceinline__ double k1_emission(double x, double m, double ls, int) { return exp((x - m) - ls); }

template <typename T>
__device__ __forceinline__ void k1_gather(const EmisArgs& em, const T* x, int b, int t, long long row, int V,
                                          typename Elem<T>::acc_t m, typename Elem<T>::acc_t ls, int lane) {
  const long long Ti = load_index(em.in_len, em.len_is64, b), Li = load_index(em.tgt_len, em.len_is64, b);
  if (t >= Ti || Li < 0 || Li > em.Lmax) return;          // padding frame / rejected utterance: never read
  double* er = em.emis + row * em.stride;
  for (int k = lane; k <= (int)Li; k += 32) {
    const bool isb = k == (int)Li;
    const long long sym = isb ? em.blank : load_index(em.targets, em.tgt_is64, (long long)b * em.ts_b + k);
    double e = 0.0;
    if (sym >= 0 && sym < V) e = k1_emission(Elem<T>::load(x + sym), m, ls, em.from_logits);
    er[isb ? em.Lmax : k] = e;
  }
}

// Register-resident variant: the whole row is loaded ONCE (NV 16-byte loads per lane, all in flight), reduced with
// warp shuffles and exponentiated from registers.
template <typename T, int VEC, int NV>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
ctc_row_stats_reg_kernel(const T* __restrict__ logits, long long stride_b, long long stride_t, int B,
                         int T_, int V, typename Elem<T>::acc_t* __restrict__ stats, const EmisArgs em) {
  using acc_t = typename Elem<T>::acc_t;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (row >= (long long)B * T_) return;
  const int b = (int)(row / T_), t = (int)(row % T_);
  const T* x = logits + b * stride_b + t * stride_t;
  acc_t v[NV][VEC];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    const int i = (j * 32 + lane) * VEC;
    if (i < V) VecLoad<T, VEC>::load(x + i, v[j]);
    else {
#pragma unroll
      for (int k = 0; k < VEC; k++) v[j][k] = -INFINITY;
    }
  }
  acc_t m = -INFINITY;
  bool has_nan = false;
#pragma unroll
  for (int j = 0; j < NV; j++)
#pragma unroll
    for (int k = 0; k < VEC; k++) { m = v[j][k] > m ? v[j][k] : m; has_nan |= (v[j][k] != v[j][k]); }
  m = warp_max(m);
  // the same summation order as the streaming kernel below (per lane: vectors in order; then the butterfly)
  acc_t s = 0;
#pragma unroll
  for (int j = 0; j < NV; j++)
#pragma unroll
    for (int k = 0; k < VEC; k++) s += exp_acc(v[j][k] - m);   // exp(-inf) = 0 for the columns past V
  s = warp_sum(s);
  has_nan = __any_sync(0xffffffffu, has_nan);
  acc_t ls = log_acc(s);
  if (has_nan) { m = NAN; ls = NAN; }
  if (lane == 0) { stats[2 * row] = m; stats[2 * row + 1] = ls; }
  if (em.emis != nullptr) k1_gather<T>(em, x, b, t, row, V, m, ls, lane);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kRowsPerBlock * 32)
ctc_row_stats_kernel(const T* __restrict__ logits, long long stride_b, long long stride_t, int B,
                     int T_, int V, bool vec_ok, typename Elem<T>::acc_t* __restrict__ stats, const EmisArgs em) {
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
  acc_t ls = log_acc(s);
  if (has_nan) { m = NAN; ls = NAN; }
  if (lane == 0) { stats[2 * row] = m; stats[2 * row + 1] = ls; }
  if (em.emis != nullptr) k1_gather<T>(em, x, b, t, row, V, m, ls, lane);
}

template <typename T, int VEC>
int launch_typed(const e2e_ctc_desc& d, const void* logits, void* stats, const EmisArgs& em, cudaStream_t s) {
  using acc_t = typename Elem<T>::acc_t;
  const long long rows = (long long)d.batch * d.max_frames;
  const unsigned grid = (unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock);
  const size_t vb = sizeof(T) * VEC;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(logits) % vb == 0) &&
                      ((d.logits_stride_b * sizeof(T)) % vb == 0) &&
                      ((d.logits_stride_t * sizeof(T)) % vb == 0) && d.alphabet >= VEC * 8;
  const T* lg = reinterpret_cast<const T*>(logits);
  acc_t* st = reinterpret_cast<acc_t*>(stats);
  KernelTimer timer(kKernelRowStats, s);
  const int V = d.alphabet;
  if (vec_ok && V % VEC == 0 && V <= 32 * VEC * 8) {
    const int nv = (V + 32 * VEC - 1) / (32 * VEC);
#define E2E_K1_REG(NV) ctc_row_stats_reg_kernel<T, VEC, NV><<<grid, kRowsPerBlock * 32, 0, s>>>(lg, d.logits_stride_b, d.logits_stride_t, d.batch, d.max_frames, V, st, em)
    if (nv <= 1) E2E_K1_REG(1); else if (nv <= 2) E2E_K1_REG(2); else if (nv <= 4) E2E_K1_REG(4); else E2E_K1_REG(8);
#undef E2E_K1_REG
  } else {
    ctc_row_stats_kernel<T, VEC><<<grid, kRowsPerBlock * 32, 0, s>>>(lg, d.logits_stride_b, d.logits_stride_t, d.batch,
                                                                    d.max_frames, V, vec_ok, st, em);
  }
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace

int launch_row_stats(const e2e_ctc_desc& d, const void* logits, void* stats, double* emis, int emis_stride,
                     const void* targets, const void* in_len, const void* tgt_len, cudaStream_t s) {
  EmisArgs em;
  em.emis = emis; em.stride = emis_stride;
  em.targets = targets; em.tgt_is64 = d.targets_itype == E2E_I64; em.ts_b = d.targets_stride_b;
  em.in_len = in_len; em.tgt_len = tgt_len; em.len_is64 = d.lengths_itype == E2E_I64;
  em.Lmax = d.max_targets; em.blank = d.blank_idx; em.from_logits = d.from_logits;
  switch (d.dtype) {
    case E2E_F32: return launch_typed<float, 4>(d, logits, stats, em, s);
    case E2E_BF16: return launch_typed<__nv_bfloat16, 8>(d, logits, stats, em, s);
    case E2E_F16: return launch_typed<__half, 8>(d, logits, stats, em, s);
    case E2E_F64: return launch_typed<double, 2>(d, logits, stats, em, s);
  }
  set_error("row_stats: unsupported dtype %d", d.dtype);
  return E2E_ERR_INVALID_ARGUMENT;
}

}  // namespace e2e
