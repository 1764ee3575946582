// K2 -- alpha/beta recursion over the blank-extended label lattice (S = 2L+1 cells).
//
// Replaces CTCLossEngine::compute_2d's alpha (src/losses/ctc_loss.cpp:33-61), loss (:63-70) and
// beta (:72-100) loops and the alpha+beta part of the gradient (:102-115).
//
// Design (B200-first):
//  * One 2-CTA thread-block cluster per utterance.  CTA rank 0 runs the forward (alpha) sweep
//    t = 0..T-1, rank 1 the backward (beta) sweep t = T-1..0, CONCURRENTLY on two SMs.  They meet in the
//    middle: each stores its first half to the workspace (L2), a single cluster barrier
//    (barrier.cluster release/acquire) publishes the halves, and in its second half each sweep
//    multiplies its live state with the other sweep's stored state, which yields the posterior
//    of every lattice cell exactly once.  The dependent chain is T steps instead of 2T.
//  * Arithmetic is LINEAR-domain fp64 with a per-lane block exponent (value = x * 2^e), not
//    log-space: a cell update is 2 DADD + 1-2 DMUL on the 64/clk/SM fp64 pipe and no MUFU, and
//    the error is ~1e-16 per step, far inside the 1e-5 parity budget where an fp32 log-space
//    recursion is not.  Each lane owns K consecutive cells (K even, so cells
//    alternate blank,label); renormalisation is lane-local integer work, lagged by one frame so it
//    stays off the dependent chain.  The s-1 / s-2 transitions cross lanes by warp shuffle and
//    cross warps through a double-buffered shared-memory slot + one named barrier per frame; the
//    repeat-label skip is a per-lane bit mask.
//  * Warp specialisation: kProducerWarps producer warps gather the per-frame emissions
//    p(t, label) = exp(logit - rowmax - logsumexp) at the label indices into a shared-memory ring
//    (mbarrier full/empty hand-off, `chunk` frames at a time); the lattice warps only do LDS.
//  * The stored half is compressed to the top 32 bits of the fp64 value (11-bit exponent, 21-bit
//    mantissa, round-to-nearest => 2.4e-7 relative) + one int32 block exponent per lane.
//
// Per-utterance outputs: loss[b]; post[b][t][s] = alpha*beta/Z as fp32 for the gradient kernel.
#include "common.cuh"

namespace e2e {
namespace {

struct LatticeParams {
  const void* logits; int dtype; long long sb, st;
  const void* stats;
  const void* targets; int tgt_is64; long long ts_b;
  const void* in_len; const void* tgt_len; int len_is64;
  int B, T, V, Lmax, blank, from_logits;
  void* losses;
  int* status; int* flags;
  uint32_t* hv; int* he; float* post;
  int cells, lanes, chunk_log2, lstride, dense, rowlen_max;
};

constexpr int kNumChunks = 4;   // E-ring / cp.async pipeline depth, in chunks of 2^chunk_log2 frames

struct __align__(16) Boundary {
  double x0, x1;
  int e, pad0, pad1, pad2;
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive (release, cluster scope) on the mbarrier at the same shared-memory offset in the peer CTA.
__device__ __forceinline__ void mbar_arrive_peer(uint64_t* bar, uint32_t peer_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(peer_rank)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void chain_barrier(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// top 32 bits of a non-negative finite double, rounded to nearest on the dropped low word
__device__ __forceinline__ uint32_t pack_hi32(double v) {
  return (uint32_t)__double2hiint(v) + ((uint32_t)__double2loint(v) >> 31);
}
__device__ __forceinline__ double unpack_hi32(uint32_t h) { return __hiloint2double((int)h, 0); }

template <int K> struct VecIO;
template <> struct VecIO<2> {
  static __device__ __forceinline__ void st_u32(uint32_t* p, const uint32_t (&v)[2]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1]);
  }
  static __device__ __forceinline__ void ld_u32(const uint32_t* p, uint32_t (&v)[2]) {
    const uint2 r = __ldcg(reinterpret_cast<const uint2*>(p));
    v[0] = r.x; v[1] = r.y;
  }
  static __device__ __forceinline__ void st_f32(float* p, const float (&v)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
};
template <> struct VecIO<4> {
  static __device__ __forceinline__ void st_u32(uint32_t* p, const uint32_t (&v)[4]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ void ld_u32(const uint32_t* p, uint32_t (&v)[4]) {
    const uint4 r = __ldcg(reinterpret_cast<const uint4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
  static __device__ __forceinline__ void st_f32(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct VecIO<8> {
  static __device__ __forceinline__ void st_u32(uint32_t* p, const uint32_t (&v)[8]) {
    reinterpret_cast<uint4*>(p)[0] = make_uint4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<uint4*>(p)[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void ld_u32(const uint32_t* p, uint32_t (&v)[8]) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void st_f32(float* p, const float (&v)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};

// ---- emission staging (producer warps) --------------------------------------------------------
// p(t, v) relative to the row's log-sum-exp, as a double.  Float inputs: the exponent argument is
// formed exactly as torch's fp32 log_softmax does ((x - max) - logsum, fp32) when the input is raw
// logits, so the emission equals exp(double(lp32)) of the reference up to one fp32 exp rounding.
__device__ __forceinline__ double emission_f32(float x, float m, float ls, int from_logits) {
  double d;
  if (from_logits) d = (double)((x - m) - ls);
  else d = (double)x - ((double)m + (double)ls);
  const float hi = (float)d;
  const float lo = (float)(d - (double)hi);
  return (double)expf(hi) * (1.0 + (double)lo);
}
__device__ __forceinline__ double emission_f64(double x, double m, double ls) { return exp((x - m) - ls); }

// cp.async (LDGSTS) of 4 / 8 / 16 bytes, global -> this CTA's shared memory, no register staging
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Producer warps: stage the emissions of `chunk` frames at a time into the E ring.
//   dense  mode (V <= Lmax+1): E[f][v]   = p(t_f, v)        for every symbol v      (rowlen = V)
//   gather mode              : E[f][0]   = p(t_f, blank), E[f][1+i] = p(t_f, label_i) (rowlen = L_i+1)
// Every item (raw logit + its row's {max, logsumexp}) is fetched with cp.async into a slot that
// only the issuing thread reads back, kNumChunks chunks ahead of its conversion, so the L2 latency
// of the gathers never reaches the lattice warps; conversion = one expf per item.
template <bool BWD, bool F64>
__device__ void run_producer(const LatticeParams& p, int b, int Ti, int Li, const int* s_lab, double* s_E,
                             unsigned char* s_raw, uint64_t* s_full, uint64_t* s_empty, double* s_lsesum,
                             int ptid, int npt) {
  constexpr int SLOT = F64 ? 32 : 16;          // {raw, max, logsumexp} per item
  const int cs = p.chunk_log2, CF = 1 << cs;
  const int rowlen = p.dense ? p.V : Li + 1;
  const int nchunks = (Ti + CF - 1) >> cs;
  const int slots_per_chunk = CF * p.rowlen_max;
  const size_t esz = p.dtype == E2E_F32 ? 4 : (p.dtype == E2E_F64 ? 8 : 2);
  const char* lbase = reinterpret_cast<const char*>(p.logits);

  auto symbol = [&](int k) { return p.dense ? k : (k == 0 ? p.blank : s_lab[k - 1]); };
  auto issue = [&](int c) {
    if (c < nchunks) {
      const int nf = min(CF, Ti - (c << cs));
      unsigned char* chunk_raw = s_raw + (size_t)(c & (kNumChunks - 1)) * slots_per_chunk * SLOT;
      for (int it = ptid; it < nf * rowlen; it += npt) {
        const int f = it / rowlen, k = it - f * rowlen;
        const int i = (c << cs) + f;
        const int t = BWD ? (Ti - 1 - i) : i;
        const long long row = (long long)b * p.T + t;
        const long long e = (long long)b * p.sb + (long long)t * p.st + symbol(k);
        unsigned char* slot = chunk_raw + (size_t)it * SLOT;
        if (F64) {
          cp_async<8>(slot, lbase + e * 8);
          cp_async<16>(slot + 16, reinterpret_cast<const char*>(p.stats) + row * 16);
        } else {
          // 16-bit elements: fetch the aligned 32-bit word that holds the element
          const char* src = lbase + e * esz;
          cp_async<4>(slot, reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3));
          cp_async<8>(slot + 8, reinterpret_cast<const char*>(p.stats) + row * 8);
        }
      }
    }
    cp_async_commit();   // one group per chunk, also when empty, so the wait depth stays uniform
  };

  for (int c = 0; c < kNumChunks; ++c) issue(c);
  double lsesum = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const int slot_c = c & (kNumChunks - 1);
    cp_async_wait<kNumChunks - 1>();                       // this thread's copies of chunk c have landed
    mbar_wait(&s_empty[slot_c], ((c / kNumChunks) & 1) ^ 1);  // the lattice warps released the E slot
    const int nf = min(CF, Ti - (c << cs));
    const unsigned char* chunk_raw = s_raw + (size_t)slot_c * slots_per_chunk * SLOT;
    for (int it = ptid; it < nf * rowlen; it += npt) {
      const int f = it / rowlen, k = it - f * rowlen;
      const unsigned char* slot = chunk_raw + (size_t)it * SLOT;
      double em;
      if (F64) {
        const double x = *reinterpret_cast<const double*>(slot);
        const double2 st = *reinterpret_cast<const double2*>(slot + 16);
        em = emission_f64(x, st.x, st.y);
      } else {
        const uint32_t raw = *reinterpret_cast<const uint32_t*>(slot);
        const float2 st = *reinterpret_cast<const float2*>(slot + 8);
        float x;
        if (p.dtype == E2E_F32) {
          x = __uint_as_float(raw);
        } else {
          const int i = (c << cs) + f;
          const int t = BWD ? (Ti - 1 - i) : i;
          const long long e = (long long)b * p.sb + (long long)t * p.st + symbol(k);
          const uint32_t half = ((reinterpret_cast<uintptr_t>(lbase) + (uintptr_t)(e * 2)) & 2) ? (raw >> 16) : (raw & 0xffffu);
          x = p.dtype == E2E_BF16 ? __uint_as_float(half << 16) : __half2float(__ushort_as_half((unsigned short)half));
        }
        em = emission_f32(x, st.x, st.y, p.from_logits);
      }
      s_E[(size_t)((slot_c << cs) + f) * p.lstride + k] = em;
    }
    if (!BWD && !p.from_logits && ptid == 0) {
      // sum of the row normalisers (log-prob input only), fixed order => deterministic loss
      for (int f = 0; f < nf; ++f) {
        const long long row = (long long)b * p.T + ((c << cs) + f);
        if (F64) lsesum += __ldg(reinterpret_cast<const double*>(p.stats) + 2 * row) + __ldg(reinterpret_cast<const double*>(p.stats) + 2 * row + 1);
        else lsesum += (double)__ldg(reinterpret_cast<const float*>(p.stats) + 2 * row) + (double)__ldg(reinterpret_cast<const float*>(p.stats) + 2 * row + 1);
      }
      if (c == nchunks - 1) *s_lsesum = lsesum;     // rides on the last hand-off (mbarrier release/acquire)
    }
    mbar_arrive(&s_full[slot_c]);
    issue(c + kNumChunks);
  }
  cp_async_wait<0>();
}

// ---- block-wide sum of per-lane values v * 2^ex over the Wi lattice warps ---------------------
// Two-phase: maximum exponent of the non-zero lanes, then the exponent-aligned sum.  Every lane of
// every active lattice warp must call it (named barrier 1).  Result: total = z * 2^ez.
__device__ __forceinline__ void block_sum_scaled(double v, int ex, double* s_redd, int* s_redi, int w,
                                                 int lane, int Wi, double* z, int* ez) {
  const int nthr = Wi * 32;
  int emax = warp_max_int(v > 0.0 ? ex : 4 * kNegExp);
  if (lane == 0) s_redi[w] = emax;
  chain_barrier(nthr);
  emax = s_redi[0];
  for (int q = 1; q < Wi; q++) emax = max(emax, s_redi[q]);
  const double part = warp_sum(v > 0.0 ? v * pow2i(ex - emax) : 0.0);
  if (lane == 0) s_redd[w] = part;
  chain_barrier(nthr);
  double t = 0.0;
  for (int q = 0; q < Wi; q++) t += s_redd[q];
  chain_barrier(nthr);   // the scratch may be reused right away
  *z = t;
  *ez = emax;
}

// ---- one lattice frame for one lane -----------------------------------------------------------
template <int K, bool BWD>
__device__ __forceinline__ void lattice_step(double (&x)[K], int& e, int& sh, const double* Erow,
                                             const int (&eidx)[K / 2], int bidx, unsigned bvalid, unsigned skipm,
                                             const Boundary* bnd_rd, Boundary* bnd_wr, int w, int Wi,
                                             int lane, double (&val)[K], int& en_out) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int H = K / 2;
  // emissions for this frame
  const double pb = Erow[bidx];
  double pl[H];
#pragma unroll
  for (int h = 0; h < H; h++) pl[h] = Erow[eidx[h]];

  // boundary cells of the neighbouring lane (previous frame, raw value + block exponent)
  double bx0, bx1 = 0.0;
  int be;
  if (!BWD) {
    bx0 = __shfl_up_sync(FULL, x[K - 1], 1);
    be = __shfl_up_sync(FULL, e, 1);
    if (lane == 0) {
      if (w > 0) { bx0 = bnd_rd[w - 1].x0; be = bnd_rd[w - 1].e; }
      else { bx0 = 0.0; be = kNegExp; }
    }
  } else {
    bx0 = __shfl_down_sync(FULL, x[0], 1);
    bx1 = __shfl_down_sync(FULL, x[1], 1);
    be = __shfl_down_sync(FULL, e, 1);
    if (lane == 31) {
      if (w + 1 < Wi) { bx0 = bnd_rd[w + 1].x0; bx1 = bnd_rd[w + 1].x1; be = bnd_rd[w + 1].e; }
      else { bx0 = 0.0; bx1 = 0.0; be = kNegExp; }
    }
  }
  // align own block (pending normalisation shift sh) and the boundary to a common exponent
  const int eo = (e == kNegExp) ? kNegExp : e - sh;
  const int en = max(eo, be);
  const double fo = pow2i(e - en);
  const double fb = pow2i(be - en);

  double y[K];
  if (!BWD) {
    // cell j gathers j, j-1 and (label cells, when allowed) j-2 of the previous frame
    const double bxs = bx0 * fb;
    y[0] = fma(fo, x[0], bxs);
    y[1] = fma(fo, x[1] + x[0], (skipm & 1u) ? bxs : 0.0);
#pragma unroll
    for (int j = 2; j < K; j++) {
      double so = x[j] + x[j - 1];
      if ((j & 1) && ((skipm >> (j >> 1)) & 1u)) so += x[j - 2];
      y[j] = fo * so;
    }
  } else {
    // cell j gathers j, j+1 and (label cells, when allowed) j+2 of the next frame
    const double bs = (bx0 + (((skipm >> (H - 1)) & 1u) ? bx1 : 0.0)) * fb;
    y[K - 1] = fma(fo, x[K - 1], bs);
    y[K - 2] = fo * (x[K - 2] + x[K - 1]);
#pragma unroll
    for (int j = K - 3; j >= 0; j--) {
      double so = x[j] + x[j + 1];
      if ((j & 1) && ((skipm >> (j >> 1)) & 1u)) so += x[j + 2];
      y[j] = fo * so;
    }
  }
  int mhi = 0;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const double pj = (j & 1) ? pl[j >> 1] : (((bvalid >> (j >> 1)) & 1u) ? pb : 0.0);
    x[j] = pj * y[j];
    mhi = max(mhi, __double2hiint(x[j]));
    val[j] = BWD ? y[j] : x[j];
  }
  // lagged lane-local renormalisation: next frame scales by 2^sh so the block maximum is in [1,2)
  if (mhi == 0) { e = kNegExp; sh = 0; }
  else { e = en; sh = 1023 - (mhi >> 20); }
  en_out = en;
  if (Wi > 1) {
    if (!BWD) { if (lane == 31) { bnd_wr[w].x0 = x[K - 1]; bnd_wr[w].e = e; } }
    else { if (lane == 0) { bnd_wr[w].x0 = x[0]; bnd_wr[w].x1 = x[1]; bnd_wr[w].e = e; } }
  }
}

template <int K, bool BWD>
__device__ void run_chain(const LatticeParams& p, int b, int Ti, int Li, const int* s_lab,
                          const double* s_E, uint64_t* s_full, uint64_t* s_empty, uint64_t* s_meet,
                          Boundary* s_bnd, double* s_redd, int* s_redi, const double* s_lsesum, int w,
                          int lane, int Wi) {
  constexpr int H = K / 2;
  constexpr int PF = 4;  // prefetch distance (frames) for the other sweep's stored half
  const int S = 2 * Li + 1;
  const int lane_g = w * 32 + lane;
  const int s0 = lane_g * K;
  const bool lane_active = s0 < S;
  const int cs = p.chunk_log2;
  const int ring_mask = (kNumChunks << cs) - 1, chunk_mask = (1 << cs) - 1;
  const int nthr = Wi * 32;

  // emission ring columns: dense mode indexes by symbol, gather mode by label position
  const int zero_slot = p.dense ? p.V : Li + 1;
  const int bidx = p.dense ? p.blank : 0;
  int eidx[H];
  unsigned bvalid = 0, skipm = 0;
#pragma unroll
  for (int h = 0; h < H; h++) {
    const int li = lane_g * H + h;  // label index of cell s0+2h+1
    if (s0 + 2 * h < S) bvalid |= 1u << h;
    const bool lv = li < Li;
    eidx[h] = lv ? (p.dense ? s_lab[li] : 1 + li) : zero_slot;
    if (lv) {
      const int lab = s_lab[li];
      bool sk;
      if (!BWD) sk = li >= 1 && lab != p.blank && lab != s_lab[li - 1];
      else sk = li + 1 < Li && lab != p.blank && s_lab[li + 1] != lab;
      if (sk) skipm |= 1u << h;
    }
  }

  double x[K];
#pragma unroll
  for (int j = 0; j < K; j++) x[j] = 0.0;
  int e = kNegExp, sh = 0;
  {  // virtual frame before the first one: all mass on the entry cell
    const int entry = BWD ? (S - 1) : 0;
    if (lane_g == entry / K) {
#pragma unroll
      for (int j = 0; j < K; j++) if (j == entry % K) x[j] = 1.0;
      e = 0;
    }
    // publish the virtual frame's boundary cells for step 0 (matters when the entry cell sits on
    // a warp edge, e.g. the backward entry S-1 landing on lane 0 of a warp)
    if (Wi > 1) {
      if (!BWD) { if (lane == 31) { s_bnd[w].x0 = x[K - 1]; s_bnd[w].e = e; } }
      else { if (lane == 0) { s_bnd[w].x0 = x[0]; s_bnd[w].x1 = x[1]; s_bnd[w].e = e; } }
    }
  }

  const int tm = Ti / 2;
  const int nstore = BWD ? (Ti - tm) : tm;   // frames this sweep stores; the rest it combines
  const size_t urow = (size_t)b * p.T;
  uint32_t* hv = p.hv;
  int* he = p.he;

  double val[K];
  int en;
  int i = 0;
  // ---------------- first half: sweep and store ----------------
  for (; i < nstore; ++i) {
    const int t = BWD ? (Ti - 1 - i) : i;
    if ((i & chunk_mask) == 0) mbar_wait(&s_full[(i >> cs) & (kNumChunks - 1)], ((i >> cs) / kNumChunks) & 1);
    if (Wi > 1) chain_barrier(nthr);
    lattice_step<K, BWD>(x, e, sh, s_E + (size_t)(i & ring_mask) * p.lstride, eidx, bidx, bvalid, skipm,
                         s_bnd + (i & 1) * 32, s_bnd + ((i + 1) & 1) * 32, w, Wi, lane, val, en);
    if (((i + 1) & chunk_mask) == 0 || i + 1 == Ti) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[(i >> cs) & (kNumChunks - 1)]);
    }
    if (lane_active) {
      uint32_t pk[K];
#pragma unroll
      for (int j = 0; j < K; j++) pk[j] = pack_hi32(val[j]);
      VecIO<K>::st_u32(hv + (urow + t) * p.cells + s0, pk);
      he[(urow + t) * p.lanes + lane_g] = en;
    }
  }
  // ---------------- the halves meet ----------------
  // Each lattice warp publishes its stored half (its lanes' global stores, ordered by the warp
  // barrier, released at cluster scope by lane 0's remote arrive) on the PEER CTA's mbarrier and
  // acquires the peer's half on its own.
  __threadfence();
  __syncwarp();
  if (lane == 0) mbar_arrive_peer(s_meet, BWD ? 0u : 1u);
  mbar_wait_cluster(s_meet, 0);
  const int ncomb = Ti - nstore;
  if (ncomb == 0) return;

  // first combine frame: also yields Z = sum_s alpha(t,s) * beta(t,s)
  double invz;
  int Ez;
  {
    const int t = BWD ? (Ti - 1 - i) : i;
    if ((i & chunk_mask) == 0) mbar_wait(&s_full[(i >> cs) & (kNumChunks - 1)], ((i >> cs) / kNumChunks) & 1);
    if (Wi > 1) chain_barrier(nthr);
    lattice_step<K, BWD>(x, e, sh, s_E + (size_t)(i & ring_mask) * p.lstride, eidx, bidx, bvalid, skipm,
                         s_bnd + (i & 1) * 32, s_bnd + ((i + 1) & 1) * 32, w, Wi, lane, val, en);
    if (((i + 1) & chunk_mask) == 0 || i + 1 == Ti) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[(i >> cs) & (kNumChunks - 1)]);
    }
    uint32_t ov[K];
    int oe = kNegExp;
#pragma unroll
    for (int j = 0; j < K; j++) ov[j] = 0;
    if (lane_active) {
      VecIO<K>::ld_u32(hv + (urow + t) * p.cells + s0, ov);
      oe = __ldcg(he + (urow + t) * p.lanes + lane_g);
    }
    double prod[K], lsum = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) { prod[j] = val[j] * unpack_hi32(ov[j]); lsum += prod[j]; }
    const int El = en + oe;
    double z;
    block_sum_scaled(lsum, El, s_redd, s_redi, w, lane, Wi, &z, &Ez);
    invz = 1.0 / z;   // z == 0 (no path survives): NaN posteriors; the utterance gets flagged below
    if (lane_active) {
      const double c = pow2i(El - Ez) * invz;
      float po[K];
#pragma unroll
      for (int j = 0; j < K; j++) po[j] = (float)(prod[j] * c);
      VecIO<K>::st_f32(p.post + (urow + t) * p.cells + s0, po);
    }
    ++i;
  }
  // ---------------- second half: sweep and combine, other half prefetched PF frames ahead -----
  uint32_t ovb[PF][K];
  int oeb[PF];
#pragma unroll
  for (int u = 0; u < PF; u++) {
    oeb[u] = kNegExp;
#pragma unroll
    for (int j = 0; j < K; j++) ovb[u][j] = 0;
    const int iu = i + u;
    if (lane_active && iu < Ti) {
      const int t = BWD ? (Ti - 1 - iu) : iu;
      VecIO<K>::ld_u32(hv + (urow + t) * p.cells + s0, ovb[u]);
      oeb[u] = __ldcg(he + (urow + t) * p.lanes + lane_g);
    }
  }
  for (; i < Ti; i += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int iu = i + u;
      if (iu < Ti) {
        const int t = BWD ? (Ti - 1 - iu) : iu;
        if ((iu & chunk_mask) == 0) mbar_wait(&s_full[(iu >> cs) & (kNumChunks - 1)], ((iu >> cs) / kNumChunks) & 1);
        if (Wi > 1) chain_barrier(nthr);
        lattice_step<K, BWD>(x, e, sh, s_E + (size_t)(iu & ring_mask) * p.lstride, eidx, bidx, bvalid, skipm,
                             s_bnd + (iu & 1) * 32, s_bnd + ((iu + 1) & 1) * 32, w, Wi, lane, val, en);
        if (((iu + 1) & chunk_mask) == 0 || iu + 1 == Ti) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[(iu >> cs) & (kNumChunks - 1)]);
        }
        if (lane_active) {
          const double c = pow2i(en + oeb[u] - Ez) * invz;
          float po[K];
#pragma unroll
          for (int j = 0; j < K; j++) po[j] = (float)(val[j] * unpack_hi32(ovb[u][j]) * c);
          VecIO<K>::st_f32(p.post + (urow + t) * p.cells + s0, po);
          const int in = iu + PF;
          if (in < Ti) {
            const int tn = BWD ? (Ti - 1 - in) : in;
            VecIO<K>::ld_u32(hv + (urow + tn) * p.cells + s0, ovb[u]);
            oeb[u] = __ldcg(he + (urow + tn) * p.lanes + lane_g);
          }
        }
      }
    }
  }
  // loss = -log(alpha[S-1][T-1] + alpha[S-2][T-1]) (ctc_loss.cpp:63-70), taken from the LIVE fp64
  // forward state (not from the 21-bit stored half), so it is accurate to fp64 rounding.
  // Emissions were normalised per row, so for log-prob input the row normalisers (all ~0 for
  // true log-probabilities) are added back.
  if (!BWD) {
    double tail = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) if (s0 + j == S - 1 || s0 + j == S - 2) tail += x[j];
    double z;
    int ez;
    block_sum_scaled(tail, e, s_redd, s_redi, w, lane, Wi, &z, &ez);
    if (w == 0 && lane == 0) {
      double loss = INFINITY;
      if (z > 0.0) {
        loss = -(log(z) + (double)ez * 0.69314718055994530942);
        if (!p.from_logits) loss -= *s_lsesum;
      } else {
        p.flags[b] = kFlagInfeasible;
      }
      store_from_double(p.losses, p.dtype, b, loss);
    }
  }
}

// ---- kernel -----------------------------------------------------------------------------------
// Block = (NW lattice + kProducerWarps producer) warps; the plan keeps NW <= kMaxLatticeWarps.
constexpr int kLatticeMaxThreads = 32 * (kMaxLatticeWarps + kProducerWarps);

template <int K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kLatticeMaxThreads, 1)
ctc_lattice_kernel(const LatticeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ring = kNumChunks << p.chunk_log2;
  double* s_E = reinterpret_cast<double*>(smem_raw);
  unsigned char* s_raw = reinterpret_cast<unsigned char*>(s_E + (size_t)ring * p.lstride);
  Boundary* s_bnd = reinterpret_cast<Boundary*>(s_raw + (size_t)ring * p.rowlen_max * (p.dtype == E2E_F64 ? 32 : 16));
  double* s_redd = reinterpret_cast<double*>(s_bnd + 64);
  double* s_lsesum = s_redd + 32;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_lsesum + 1);
  uint64_t* s_empty = s_full + 8;
  uint64_t* s_meet = s_empty + 7;   // rings use at most 4 slots of the 8 reserved
  int* s_redi = reinterpret_cast<int*>(s_empty + 8);
  int* s_misc = s_redi + 32;  // [0] argument-check bits, [1] adjacent repeats
  int* s_lab = s_misc + 4;

  const int b = blockIdx.x >> 1;
  const bool bwd = cluster_ctarank() == 1;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int NW = (blockDim.x >> 5) - kProducerWarps;

  const long long Ti_ll = load_index(p.in_len, p.len_is64, b);
  const long long Li_ll = load_index(p.tgt_len, p.len_is64, b);
  int bad = 0;
  if (Ti_ll < 1 || Ti_ll > p.T) bad |= kBadFrames;
  if (Li_ll < 0 || Li_ll > p.Lmax) bad |= kBadTargetLen;
  const int Ti = (int)Ti_ll, Li = bad ? 0 : (int)Li_ll;
  if (tid < 4) s_misc[tid] = 0;
  __syncthreads();
  int rep = 0, badlab = 0;
  for (int i = tid; i < Li; i += blockDim.x) {
    const long long v = load_index(p.targets, p.tgt_is64, (long long)b * p.ts_b + i);
    if (v < 0 || v >= p.V) badlab = kBadLabel;
    s_lab[i] = (int)v;
  }
  __syncthreads();
  for (int i = tid + 1; i < Li; i += blockDim.x) rep += (s_lab[i] == s_lab[i - 1]);
  if (rep) atomicAdd(&s_misc[1], rep);
  if (badlab) atomicOr(&s_misc[0], badlab);
  __syncthreads();
  bad |= s_misc[0];
  rep = s_misc[1];
  if (bad) {  // undefined behaviour in the reference: reject (NaN loss, NaN gradient block, status)
    if (!bwd && tid == 0) {
      atomicOr(p.status, bad);
      p.flags[b] = kFlagInvalid;
      store_from_double(p.losses, p.dtype, b, (double)NAN);
    }
    return;
  }
  if (Ti < Li + rep) {  // no alignment exists: loss = +inf, gradient block all NaN
    if (!bwd && tid == 0) {
      p.flags[b] = kFlagInfeasible;
      store_from_double(p.losses, p.dtype, b, (double)INFINITY);
    }
    return;
  }
  const int S = 2 * Li + 1;
  const int Wi = ((S + K - 1) / K + 31) / 32;
  if (tid == 0) {
    if (!bwd) p.flags[b] = 0;
    for (int c = 0; c < kNumChunks; c++) {
      mbar_init(&s_full[c], kProducerWarps * 32);
      mbar_init(&s_empty[c], Wi);
    }
    mbar_init(s_meet, Wi);
    *s_lsesum = 0.0;
  }
  for (int f = tid; f < ring; f += blockDim.x) s_E[(size_t)f * p.lstride + (p.dense ? p.V : Li + 1)] = 0.0;
  for (int q = tid; q < 64; q += blockDim.x) { s_bnd[q].x0 = 0.0; s_bnd[q].x1 = 0.0; s_bnd[q].e = kNegExp; }
  __syncthreads();
  // Both CTAs of the pair took the same early-exit decisions above, so both reach this point:
  // the peer's mbarriers exist before anyone arrives on them remotely.
  cluster_arrive();
  cluster_wait();

  if (w >= NW) {  // emission producers
    const int ptid = tid - NW * 32;
    const int npt = kProducerWarps * 32;
    if (p.dtype == E2E_F64) {
      if (bwd) run_producer<true, true>(p, b, Ti, Li, s_lab, s_E, s_raw, s_full, s_empty, s_lsesum, ptid, npt);
      else run_producer<false, true>(p, b, Ti, Li, s_lab, s_E, s_raw, s_full, s_empty, s_lsesum, ptid, npt);
    } else {
      if (bwd) run_producer<true, false>(p, b, Ti, Li, s_lab, s_E, s_raw, s_full, s_empty, s_lsesum, ptid, npt);
      else run_producer<false, false>(p, b, Ti, Li, s_lab, s_E, s_raw, s_full, s_empty, s_lsesum, ptid, npt);
    }
    return;
  }
  if (w >= Wi) return;
  if (bwd) run_chain<K, true>(p, b, Ti, Li, s_lab, s_E, s_full, s_empty, s_meet, s_bnd, s_redd, s_redi, s_lsesum, w, lane, Wi);
  else run_chain<K, false>(p, b, Ti, Li, s_lab, s_E, s_full, s_empty, s_meet, s_bnd, s_redd, s_redi, s_lsesum, w, lane, Wi);
}

template <int K>
int launch_k(const LatticeParams& lp, const LossPlan& p, cudaStream_t s) {
  E2E_CUDA_TRY(cudaFuncSetAttribute(ctc_lattice_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  const unsigned threads = 32u * (unsigned)(p.NW + kProducerWarps);
  KernelTimer timer(kKernelLattice, s);
  ctc_lattice_kernel<K><<<2u * (unsigned)lp.B, threads, p.smem, s>>>(lp);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace

int launch_lattice(const e2e_ctc_desc& d, const LossPlan& p, const void* logits, const void* targets,
                   const void* in_len, const void* tgt_len, void* losses, char* ws, cudaStream_t s) {
  LatticeParams lp;
  lp.logits = logits; lp.dtype = d.dtype; lp.sb = d.logits_stride_b; lp.st = d.logits_stride_t;
  lp.stats = ws + p.off_stats;
  lp.targets = targets; lp.tgt_is64 = d.targets_itype == E2E_I64; lp.ts_b = d.targets_stride_b;
  lp.in_len = in_len; lp.tgt_len = tgt_len; lp.len_is64 = d.lengths_itype == E2E_I64;
  lp.B = d.batch; lp.T = d.max_frames; lp.V = d.alphabet; lp.Lmax = d.max_targets;
  lp.blank = d.blank_idx; lp.from_logits = d.from_logits;
  lp.losses = losses;
  lp.status = reinterpret_cast<int*>(ws + p.off_status);
  lp.flags = reinterpret_cast<int*>(ws + p.off_flags);
  lp.hv = reinterpret_cast<uint32_t*>(ws + p.off_hv);
  lp.he = reinterpret_cast<int*>(ws + p.off_he);
  lp.post = reinterpret_cast<float*>(ws + p.off_post);
  lp.cells = p.cells; lp.lanes = p.lanes; lp.chunk_log2 = p.chunk_log2; lp.lstride = p.lstride;
  lp.dense = p.dense; lp.rowlen_max = p.rowlen;
  switch (p.K) {
    case 2: return launch_k<2>(lp, p, s);
    case 4: return launch_k<4>(lp, p, s);
    case 8: return launch_k<8>(lp, p, s);
  }
  set_error("lattice: unsupported cells-per-lane %d", p.K);
  return E2E_ERR_UNSUPPORTED;
}

}  // namespace e2e
