// K3 -- logits gradient, and K4 -- loss reduction.
//
// K3 replaces the tail of CTCLossEngine::compute_2d (src/losses/ctc_loss.cpp:102-117:
// grads = exp(lp) - exp(label-summed(alpha+beta) - logZ)), the autograd multiply
// `grads * grad_output.view(-1,1,1)` (functions/forward_backward.py:34) and -- when the input was
// raw logits -- the log_softmax backward of modules/ctc_loss.py:40, in ONE pass over the logits:
//
//     grad[b,t,v] = scale_b * ( softmax(x[b,t,:])[v] - sum_{s: ext[s]==v} post[b,t,s] )
//
// One warp per frame row.  The per-cell posteriors written by the lattice kernel are summed per
// label with shared-memory atomics into a V-float row (blank cells, half of the lattice, are
// pre-reduced with a warp shuffle), then the row is written back coalesced together with the
// softmax term.  Padding rows (t >= T_i) carry exp(lp) for log-prob input (the reference's engine
// contract) or 0 for fused-logits input (what the reference's log_softmax backward leaves there);
// an infeasible utterance gets an all-NaN block, padding rows included.
//
// HBM traffic: reads logits once, writes grads once (+ the lattice posteriors, which are L2-hot).
#include "common.cuh"

namespace e2e {
namespace {

struct GradParams {
  const void* logits; void* grads; long long sb, st, gsb, gst;
  const void* stats; const float* post; const int* flags;
  const void* targets; int tgt_is64; long long ts_b;
  const void* in_len; const void* tgt_len; int len_is64;
  const void* grad_out; int grad_out_count; double host_scale;
  int B, T, V, Lmax, blank, from_logits, cells, post_stride, rows_per_block, vec_ok, vstride;
};

__device__ __forceinline__ float exp_acc(float x) { return expf(x); }
__device__ __forceinline__ double exp_acc(double x) { return exp(x); }

// 16-byte vectors of the element type (the row loops move 128 bits per lane per access when the rows are aligned)
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p)); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec16<double> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void load(const double* p, double (&o)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p)); o[0] = v.x; o[1] = v.y;
  }
  static __device__ __forceinline__ void store(double* p, const double (&v)[2]) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&o)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const GradParams p) {
  using acc_t = typename Elem<T>::acc_t;
  constexpr int N = Vec16<T>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* s_lab = reinterpret_cast<int*>(smem_raw);
  float* s_acc = reinterpret_cast<float*>(s_lab + ((p.Lmax + 3) & ~3));

  const int tiles = (p.T + p.rows_per_block - 1) / p.rows_per_block;
  const int b = blockIdx.x / tiles;
  const int t0 = (blockIdx.x % tiles) * p.rows_per_block;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int flag = p.flags[b];
  int Ti = 0, Li = 0;
  if (!flag) {
    Ti = (int)load_index(p.in_len, p.len_is64, b);
    Li = (int)load_index(p.tgt_len, p.len_is64, b);
  }
  const bool any_valid = !flag && t0 < Ti;
  if (any_valid)
    for (int i = threadIdx.x; i < Li; i += blockDim.x)
      s_lab[i] = (int)load_index(p.targets, p.tgt_is64, (long long)b * p.ts_b + i);
  __syncthreads();

  const int t = t0 + w;
  if (w >= p.rows_per_block || t >= p.T) return;

  acc_t scale = (acc_t)p.host_scale;
  if (p.grad_out != nullptr)
    scale *= Elem<T>::load(reinterpret_cast<const T*>(p.grad_out) + (p.grad_out_count == 1 ? 0 : b));

  const T* x = reinterpret_cast<const T*>(p.logits) + b * p.sb + t * p.st;
  T* g = reinterpret_cast<T*>(p.grads) + b * p.gsb + t * p.gst;
  const long long row = (long long)b * p.T + t;
  const int nvec = p.vec_ok ? p.V / N : 0;     // 128-bit accesses over [0, nvec*N), scalar tail after

  if (flag) {  // infeasible (or rejected) utterance: NaN block, as -inf - (-inf) gives in the reference
    acc_t nanv[N];
#pragma unroll
    for (int k = 0; k < N; k++) nanv[k] = (acc_t)NAN;
    for (int i = lane; i < nvec; i += 32) Vec16<T>::store(g + i * N, nanv);
    for (int v = nvec * N + lane; v < p.V; v += 32) Elem<T>::store(g + v, (acc_t)NAN);
    return;
  }
  if (t >= Ti) {  // padding frame
    if (p.from_logits) {
      acc_t z[N];
#pragma unroll
      for (int k = 0; k < N; k++) z[k] = scale * (acc_t)0;
      for (int i = lane; i < nvec; i += 32) Vec16<T>::store(g + i * N, z);
      for (int v = nvec * N + lane; v < p.V; v += 32) Elem<T>::store(g + v, scale * (acc_t)0);
    } else {
      for (int i = lane; i < nvec; i += 32) {
        acc_t xv[N];
        Vec16<T>::load(x + i * N, xv);
#pragma unroll
        for (int k = 0; k < N; k++) xv[k] = scale * exp_acc(xv[k]);
        Vec16<T>::store(g + i * N, xv);
      }
      for (int v = nvec * N + lane; v < p.V; v += 32) Elem<T>::store(g + v, scale * exp_acc(Elem<T>::load(x + v)));
    }
    return;
  }

  // the row's logits are requested BEFORE the posterior scatter so that the HBM latency overlaps it
  constexpr int PRE = 8;                      // vectors per lane held in registers: rows up to 32*N*PRE elements
  acc_t xr[PRE][N];
  const bool pre = nvec > 0 && nvec <= 32 * PRE;
  if (pre) {
#pragma unroll
    for (int j = 0; j < PRE; j++) { const int i = j * 32 + lane; if (i < nvec) Vec16<T>::load(x + i * N, xr[j]); }
  }
  // compact posterior row written by the lattice kernel's combiners: [label 0 .. label L-1 | ... | blank total].  Its
  // values (<= 4 per lane up to 128 labels) are requested together with the row, ahead of the accumulator clear: the
  // scatter below waited on these loads for 28 % of the kernel's samples
  const float* post = p.post + row * p.post_stride;
  constexpr int PK = 4;
  float pv[PK];
  const bool post_pre = Li <= 32 * PK;
  if (post_pre) {
#pragma unroll
    for (int q = 0; q < PK; q++) { const int li = lane + 32 * q; pv[q] = li < Li ? __ldcg(post + li) : 0.f; }
  }
  const float pblank = lane == 0 ? __ldcg(post + p.cells / 2) : 0.f;
  float* acc = s_acc + (size_t)w * p.vstride;
  for (int v = lane * 4; v < p.vstride; v += 128) *reinterpret_cast<float4*>(acc + v) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  if (post_pre) {
#pragma unroll
    for (int q = 0; q < PK; q++) { const int li = lane + 32 * q; if (li < Li) atomicAdd(&acc[s_lab[li]], pv[q]); }
  } else {
    for (int li = lane; li < Li; li += 32) atomicAdd(&acc[s_lab[li]], __ldcg(post + li));
  }
  if (lane == 0) atomicAdd(&acc[p.blank], pblank);
  __syncwarp();

  acc_t m = 0, ls = 0;
  if (p.from_logits) {
    m = reinterpret_cast<const acc_t*>(p.stats)[2 * row];
    ls = reinterpret_cast<const acc_t*>(p.stats)[2 * row + 1];
  }
  auto one = [&](acc_t xv, float a) {
    const acc_t first = p.from_logits ? exp_acc((xv - m) - ls) : exp_acc(xv);
    return scale * (first - (acc_t)a);
  };
  if (pre) {
#pragma unroll
    for (int j = 0; j < PRE; j++) {
      const int i = j * 32 + lane;
      if (i < nvec) {
        acc_t r[N];
#pragma unroll
        for (int k = 0; k < N; k++) r[k] = one(xr[j][k], acc[i * N + k]);
        Vec16<T>::store(g + i * N, r);
      }
    }
  } else {
    for (int i = lane; i < nvec; i += 32) {
      acc_t xv[N], r[N];
      Vec16<T>::load(x + i * N, xv);
#pragma unroll
      for (int k = 0; k < N; k++) r[k] = one(xv[k], acc[i * N + k]);
      Vec16<T>::store(g + i * N, r);
    }
  }
  for (int v = nvec * N + lane; v < p.V; v += 32) Elem<T>::store(g + v, one(Elem<T>::load(x + v), acc[v]));
}

template <typename T>
int launch_typed(const GradParams& gp, cudaStream_t s) {
  GradParams p = gp;
  constexpr size_t vb = 16;
  p.vec_ok = (reinterpret_cast<uintptr_t>(p.logits) % vb == 0) && (reinterpret_cast<uintptr_t>(p.grads) % vb == 0) &&
             ((p.sb * sizeof(T)) % vb == 0) && ((p.st * sizeof(T)) % vb == 0) && ((p.gsb * sizeof(T)) % vb == 0) &&
             ((p.gst * sizeof(T)) % vb == 0) && p.V >= 32;
  p.vstride = (p.V + 3) & ~3;
  const size_t lab_bytes = (size_t)((p.Lmax + 3) & ~3) * sizeof(int);
  int rpb = 8;
  while (rpb > 1 && lab_bytes + (size_t)rpb * p.vstride * sizeof(float) > 96 * 1024) rpb >>= 1;
  p.rows_per_block = rpb;
  const size_t smem = lab_bytes + (size_t)rpb * p.vstride * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("grad: alphabet %d too large for the shared-memory row accumulator", p.V);
    return E2E_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    E2E_CUDA_TRY(cudaFuncSetAttribute(ctc_grad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (p.T + rpb - 1) / rpb;
  KernelTimer timer(kKernelGrad, s);
  ctc_grad_kernel<T><<<(unsigned)(p.B * tiles), 256, smem, s>>>(p);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

// ---- K4: losses[B] -> scale * sum, fp64 accumulation in a fixed order (deterministic) ----------
__global__ void __launch_bounds__(256)
ctc_loss_reduce_kernel(const void* losses, int dtype, int B, double scale, void* out, double* out64) {
  __shared__ double s[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < B; i += 256) a += load_as_double(losses, dtype, i);
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double r = s[0] * scale;
    if (out) store_from_double(out, dtype, 0, r);
    if (out64) { out64[0] = r; out64[1] = (double)B; }  // {partial sum, utterance count} for the all-reduce
  }
}

// ---- in-place gradient scaling by the upstream gradient (functions/forward_backward.py:34) -------
// The fused forward already wrote scale * d loss / d logits; autograd's grad_output is almost always
// exactly 1, which only the device knows: every block checks its utterance's factor first and leaves
// without touching memory when it is 1.  NaN blocks stay NaN (NaN * 0 = NaN, as in the reference).
template <typename T>
__global__ void __launch_bounds__(256)
ctc_scale_rows_kernel(T* grads, long long gsb, long long gst, int T_, int V, const T* grad_out, int count) {
  using acc_t = typename Elem<T>::acc_t;
  const int b = blockIdx.y;
  const acc_t g = Elem<T>::load(grad_out + (count == 1 ? 0 : b));
  if (g == (acc_t)1) return;
  const long long n = (long long)T_ * V;
  for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < n; k += (long long)gridDim.x * 256) {
    const long long t = k / V, v = k - t * V;
    T* q = grads + b * gsb + t * gst + v;
    Elem<T>::store(q, Elem<T>::get(*q) * g);
  }
}

template <typename T>
int launch_scale_typed(const e2e_ctc_desc& d, void* grads, const void* grad_out, int count, cudaStream_t s) {
  const long long n = (long long)d.max_frames * d.alphabet;
  unsigned gx = (unsigned)((n + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  KernelTimer timer(kKernelScale, s);
  ctc_scale_rows_kernel<T><<<dim3(gx, (unsigned)d.batch), 256, 0, s>>>(
      reinterpret_cast<T*>(grads), d.grads_stride_b, d.grads_stride_t, d.max_frames, d.alphabet,
      reinterpret_cast<const T*>(grad_out), count);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace

int launch_scale_rows(const e2e_ctc_desc& d, void* grads, const void* grad_out, int grad_out_count, cudaStream_t s) {
  switch (d.dtype) {
    case E2E_F32: return launch_scale_typed<float>(d, grads, grad_out, grad_out_count, s);
    case E2E_BF16: return launch_scale_typed<__nv_bfloat16>(d, grads, grad_out, grad_out_count, s);
    case E2E_F16: return launch_scale_typed<__half>(d, grads, grad_out, grad_out_count, s);
    case E2E_F64: return launch_scale_typed<double>(d, grads, grad_out, grad_out_count, s);
  }
  set_error("scale_rows: unsupported dtype %d", d.dtype);
  return E2E_ERR_INVALID_ARGUMENT;
}

int launch_grad(const e2e_ctc_desc& d, const LossPlan& pl, const void* logits, const void* targets,
                const void* in_len, const void* tgt_len, const void* grad_out, int grad_out_count,
                double host_scale, void* grads, const char* ws, cudaStream_t s) {
  GradParams p;
  p.logits = logits; p.grads = grads;
  p.sb = d.logits_stride_b; p.st = d.logits_stride_t; p.gsb = d.grads_stride_b; p.gst = d.grads_stride_t;
  p.stats = ws + pl.off_stats;
  p.post = reinterpret_cast<const float*>(ws + pl.off_post);
  p.flags = reinterpret_cast<const int*>(ws + pl.off_flags);
  p.targets = targets; p.tgt_is64 = d.targets_itype == E2E_I64; p.ts_b = d.targets_stride_b;
  p.in_len = in_len; p.tgt_len = tgt_len; p.len_is64 = d.lengths_itype == E2E_I64;
  p.grad_out = grad_out; p.grad_out_count = grad_out_count; p.host_scale = host_scale;
  p.B = d.batch; p.T = d.max_frames; p.V = d.alphabet; p.Lmax = d.max_targets; p.blank = d.blank_idx;
  p.from_logits = d.from_logits; p.cells = pl.cells; p.post_stride = pl.post_stride; p.rows_per_block = 8;
  switch (d.dtype) {
    case E2E_F32: return launch_typed<float>(p, s);
    case E2E_BF16: return launch_typed<__nv_bfloat16>(p, s);
    case E2E_F16: return launch_typed<__half>(p, s);
    case E2E_F64: return launch_typed<double>(p, s);
  }
  set_error("grad: unsupported dtype %d", d.dtype);
  return E2E_ERR_INVALID_ARGUMENT;
}

int launch_reduce(const void* losses, int dtype, int B, double scale, void* out, double* out64,
                  cudaStream_t s) {
  KernelTimer timer(kKernelReduce, s);
  ctc_loss_reduce_kernel<<<1, 256, 0, s>>>(losses, dtype, B, scale, out, out64);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace e2e
