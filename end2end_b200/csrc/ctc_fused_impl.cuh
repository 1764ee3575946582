// K2 -- the CTC lattice kernel: row log-softmax, alpha/beta recursion, posteriors and (small alphabets) the
// gradient write, in one launch.  ONE kernel design for every shape (it replaces round 1's three).
//
// Replaces CTCLossEngine::compute_2d (src/losses/ctc_loss.cpp:15-118): extended targets (:25-31),
// alpha (:33-61), loss (:63-70), beta (:72-100), alpha+beta / gradient (:102-117), plus F.log_softmax
// (pytorch_end2end/modules/ctc_loss.py:40) and the exp(logits) term of the gradient (:117).
//
// Design:
//  * One 2-CTA cluster per utterance: CTA 0 sweeps alpha forward in time, CTA 1 sweeps beta backward,
//    concurrently.  The beta recursion is the alpha recursion of the REVERSED label sequence over REVERSED
//    time, so both CTAs run the same code on reversed data.  Each sweep stores its first half of the frames
//    to a global stash (L2), the two meet once through a global flag, and in its second half each multiplies
//    its own cells with the other sweep's stashed row: the dependent chain is T frames, not 2T, and every
//    posterior alpha*beta/Z is produced exactly once.
//  * ONE LATTICE WARP PER SWEEP.  Lane l owns NBU consecutive blocks of four cells (blank,label,blank,label)
//    in registers: cells [16*NBU*l .. +4*NBU).  The s-1 / s-2 transitions are register moves inside a lane and
//    ONE 64-bit shuffle per frame across lanes -- no shared-memory hand-off, no barrier, no polling on the
//    recurrence (round 1 spread a sweep over four warps and paid ~200 cycles per cross-warp hop).  NBU is
//    chosen per utterance (ceil((2L+1)/128)), so short utterances of a batch sweep fewer cells.
//  * Arithmetic: LINEAR-domain fp64 with one block exponent per 4-cell block (value = x * 2^e): a cell update
//    is DADD (+ predicated DADD/DFMA for the repeat-label skip) + DMUL, nothing transcendental on the chain;
//    four cells x 149 bits (the smallest fp32 emission) stay inside the fp64 range whatever the input.
//    Exponents are re-centred once per 4-frame group from a snapshot two frames earlier (folded into the
//    emission multipliers): a block WITH mass takes its own maximum but never less than the previous block's
//    new exponent - 192; a block WITHOUT mass takes the nearest massive block's exponent -- one prefix maximum
//    over the blocks in lattice order (in-lane sequential + one warp scan).
//  * Warp roles around the lattice warp, decoupled by shared-memory rings and monotonic progress words
//    (st.release / ld.acquire at CTA scope; no block barrier in the frame loops):
//      producers : fused row log-softmax, ONE LANE PER FRAME over blocks of 32 frames -> E ring (doubles).
//                  Small alphabets (dense mode): whole rows, indexed by symbol.  Large alphabets (gather mode):
//                  only blank + the utterance's labels, gathered with the row statistics K1 wrote.
//      combiners : drain the val ring (top 32 bits of each fp64 cell + block exponents), frames round-robin.
//                  First half: copy rows to the global stash.  Second half: cp.async-prefetch the other sweep's
//                  row, multiply, normalise by Z = sum_s alpha*beta, then dense mode: group label posteriors by
//                  symbol (counting sort built once; no atomics, bitwise reproducible) and write the gradient
//                  row scale*(softmax - posterior); gather mode: write compact per-label posteriors for K3.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace e2e {
namespace {

#ifndef FZ_REGS4
#define FZ_REGS4 128
#endif
constexpr int kFzG = 4;         // frames between exponent re-centrings
constexpr int kFzSlack = 192;   // a massive block's exponent is at most this far below its predecessor's

struct FzParams {
  const void* logits; int dtype; long long sb, st;
  void* grads; long long gsb, gst; double scale;
  const void* stats;                      // gather mode: row {max, logsumexp} from K1
  const double* emis; int emis_stride;    // gather mode: compact emission rows from K1: [label 0 .. Lmax-1 | blank] per frame
  float* post; int post_stride, cells;    // gather mode: compact posteriors [B*T][post_stride], blank total at cells/2
  const void* targets; int tgt_is64; long long ts_b;
  const void* in_len; const void* tgt_len; int len_is64;
  int B, T, V, Lmax, blank, from_logits;
  void* losses;
  int* status; int* flags; int* meet;
  uint32_t* stash;                        // [B*T][roww]: 4*32*NBU cell words in lattice order, then 32*NBU block exponents
  int roww;                               // words per stash row (sized for the kernel class)
  FzLayout L;
};

// control block (shared memory)
// Hand-offs that publish DATA (emission blocks, val-ring frames) go through mbarriers: arrive has release and
// try_wait acquire semantics without a memory fence instruction, and a waiting warp is suspended instead of
// spinning (a st.release / ld.acquire pair on a shared-memory word costs a MEMBAR per store: ~500 cycles per
// frame on the lattice warp, measured).  Hand-offs that only free a ring slot for REUSE (write-after-read) use
// plain volatile progress words: the reader has consumed its loads (data dependence) before it stores them.
struct FzCtl {
  unsigned long long full[32];   // val-ring slot s holds frame i (i % RV == s): phase i / RV
  unsigned long long fullE[16];  // emission-ring block slot: phase (block / (R / PB))
  unsigned long long sc_full[2]; // exponent update n (from frame 4n+1, applied at frame 4n+4) is in scale-table slot n % 2
  volatile int lat_prog;         // frames [0, value) swept (their emission rows are no longer read by the lattice)
  int pad1;
  int zero;                      // no path survives / NaN input
  int pad0;
  volatile int comb_done[8];     // combiner q: the next frame it will take (all its earlier frames are done)
  double tail_x[2]; int tail_e[2]; int tail_on[2];
  double lse[4];
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fz_cp_async_cg16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fz_smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void fz_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fz_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int fz_ld_acquire_gpu(const int* p) {
  int r;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void fz_st_release(volatile int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(fz_smem_u32(const_cast<int*>(p))), "r"(v) : "memory");
}
__device__ __forceinline__ int fz_ld_acquire(const volatile int* p) {
  int r;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(r) : "r"(fz_smem_u32(const_cast<int*>(p))) : "memory");
  return r;
}
__device__ __forceinline__ void fz_mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fz_mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(fz_smem_u32(bar)) : "memory");
}
// the same arrive, predicated inside the asm so that the caller stays branch-free
__device__ __forceinline__ void fz_mbar_arrive_if(unsigned long long* bar, int on) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 st;\n\tsetp.ne.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(fz_smem_u32(bar)), "r"(on) : "memory");
}
__device__ __forceinline__ void fz_mbar_wait(unsigned long long* bar, int parity) {
  uint32_t done = 0;
  const uint32_t a = fz_smem_u32(bar);
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ double fz_hi2d(uint32_t h) { return __hiloint2double((int)h, 0); }

// p(t, v) relative to the row's log-sum-exp, as a double.  Raw logits: the exponent argument is formed as
// torch's fp32 log_softmax does ((x - max) - logsum, fp32), so the emission equals exp(double(lp32)) of the
// reference up to one fp32 exp rounding.  Log-prob input: the argument x - (m + ls) is split into an fp32 head
// and a tail; a -inf log-prob (a masked symbol) is an exact zero emission, as in the reference's log_sum_exp.
__device__ __noinline__ double fz_emission_lp(float x, float m, float ls) {
  const double d = (double)x - ((double)m + (double)ls);
  const float hi = (float)d;
  if (!(hi > -INFINITY)) return hi != hi ? (double)hi : 0.0;   // -inf -> 0, NaN -> NaN
  const float lo = (float)(d - (double)hi);
  return (double)expf(hi) * (1.0 + (double)lo);
}
__device__ __forceinline__ double fz_emission(float x, float m, float ls, int from_logits) {
  if (from_logits) return (double)expf((x - m) - ls);
  return fz_emission_lp(x, m, ls);
}

#ifdef FZ_DBG
__device__ long long g_fz_dbg[128];
#define FZ_CLK() clock64()
#else
#define FZ_CLK() 0ll
#endif

struct FzView {
  int* lab; double* sc; double* E; unsigned char* val; uint32_t* stage; float* post; FzCtl* ctl;
};
__device__ __forceinline__ FzView fz_carve(unsigned char* base, const FzLayout& L) {
  FzView v;
  v.lab = reinterpret_cast<int*>(base + L.off_lab);
  v.sc = reinterpret_cast<double*>(base + L.off_occ);   // scale table: 2 slots x {f[GT], fb[GT] doubles, en[GT] ints}
  v.E = reinterpret_cast<double*>(base + L.off_E);
  v.val = base + L.off_val;
  v.stage = reinterpret_cast<uint32_t*>(base + L.off_stage);
  v.post = reinterpret_cast<float*>(base + L.off_post);
  v.ctl = reinterpret_cast<FzCtl*>(base + L.off_ctl);
  return v;
}

// val-ring frame: NBU rows of 32 uint4 (the top words of a block's four cells), row stride padded so that the
// combiners' lattice-order reads (block g = lane + 32u lives in row g % NBU, lane g / NBU) spread over the
// banks; then the block exponents, [lane][NBP].
template <int NBU> struct FzGeom {
  static constexpr int NBP = (NBU + 3) & ~3;
  static constexpr int P = 512 + 16 * ((8 + NBU - 1) / NBU);   // bytes per row
  static constexpr int G = 32 * NBU;                           // blocks per sweep
};

__device__ __forceinline__ int fz_min_done(const volatile int* a, int n) {
  int m = a[0];
  for (int k = 1; k < n; k++) m = min(m, a[k]);
  return m;
}

// ---- producers: fused row log-softmax -> E ring ---------------------------------------------------
__device__ __forceinline__ float fz_load_logit(const void* base, int dtype, long long idx) {
  if (dtype == E2E_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  const unsigned short r = __ldg(reinterpret_cast<const unsigned short*>(base) + idx);
  return dtype == E2E_BF16 ? __uint_as_float((uint32_t)r << 16) : __half2float(__ushort_as_half(r));
}

// dense mode, alphabets of <= 32 symbols (32 / 16-bit logits): ONE LANE PER FRAME over passes of 32 frames.  A lane loads
// its whole row into registers (all loads independent and in flight together), reduces it serially -- no cross-lane
// shuffles -- and writes the row's emissions; the HBM round trip (~2000 cycles) is paid once
// per 32 frames instead of once per four (the warp-per-frame version cost 775 cycles per frame per producer, and the
// lattice warp waited for emissions 27 % of its time).  A pass covers 32 / PB emission blocks; each is published on its
// own mbarrier.  The ring must hold two passes (R >= 64).
__device__ __noinline__ double fz_produce_lanes(const FzParams& p, const FzView& sv, long long xbase, int Ti, int i0, int lane, bool BWD) {
  const FzLayout& L = p.L;
  const int i = i0 + lane;
  if (i >= Ti) return 0.0;
  const long long ro = xbase + (long long)(BWD ? (Ti - 1 - i) : i) * p.st;
  double* Erow = sv.E + (size_t)(i & (L.R - 1)) * L.es;
  const int V = p.V;
  float m = -INFINITY, s = 0.f;
  bool nan = false;
  float xv[32];
#pragma unroll
  for (int v = 0; v < 32; v++) xv[v] = v < V ? fz_load_logit(p.logits, p.dtype, ro + v) : -INFINITY;
#pragma unroll
  for (int v = 0; v < 32; v++) { nan |= xv[v] != xv[v]; m = fmaxf(m, xv[v]); }
#pragma unroll
  for (int v = 0; v < 32; v++) s += expf(xv[v] - m);   // exp(-inf) = 0 for the padding columns
  float ls = logf(s);
  if (nan) { m = NAN; ls = NAN; }                       // a NaN in the row poisons it, as in torch's log_softmax
#pragma unroll
  for (int v = 0; v < 32; v++) if (v < V) Erow[v] = fz_emission(xv[v], m, ls, p.from_logits);
  const double mls = (double)m + (double)ls;
  Erow[V] = 0.0;                                       // the column padding cells read
  Erow[V + 1] = p.from_logits ? 1.0 : exp(mls);        // turns the emission back into exp(x)
  return mls;
}

__device__ __forceinline__ double fz_producer_dense(const FzParams& p, const FzView& sv, int b, int Ti, int pw, int lane, bool BWD) {
  constexpr int PASS = 32;
  const FzLayout& L = p.L;
  const int npass = (Ti + PASS - 1) / PASS, per = PASS / L.PB;
  const long long xbase = (long long)b * p.sb;
  double lse = 0.0;
  for (int ps = pw; ps < npass; ps += L.NP) {
    const int i0 = ps * PASS;
    const int need = i0 + PASS - L.R;   // frames below `need` must have left the ring (write-after-read: plain progress words)
    if (need > 0) {
      while (sv.ctl->lat_prog < need || fz_min_done(sv.ctl->comb_done, L.NC) < need) { if (L.nap) __nanosleep(L.nap); }
    }
    lse += fz_produce_lanes(p, sv, xbase, Ti, i0, lane, BWD);
    __syncwarp();
    if (lane < per && i0 + lane * L.PB < Ti) fz_mbar_arrive(&sv.ctl->fullE[(ps * per + lane) & L.neb_mask]);
  }
  return warp_sum(lse);   // fixed order: deterministic loss for log-prob input
}

// dense mode: one WARP per frame, lanes over the symbols (coalesced row loads, warp-shuffle max / sum), four frames
// per iteration so that the shuffle chains of independent rows overlap; the next group's loads are issued before
// this group is reduced (HBM latency off the chain).  Deliberately a ROLLED loop in a separate function: the SM's
// L1.5 instruction cache is 32 KB for all warps of all roles (B300_MICROARCH.md), and a fully unrolled block of
// eight frames alone was ~15 KB.  VPL = symbols per lane.
template <int VPL>
__device__ __noinline__ double fz_produce_dense(const void* logits, int dtype, long long st, int V, int from_logits,
                                                double* E, int Rm, int es, long long xbase, int Ti, int i0, int i1,
                                                int lane, bool BWD) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int F = 4;
  double lse = 0.0;
  float xn[F][VPL];
  auto load = [&](int ib) {
#pragma unroll
    for (int f = 0; f < F; f++) {
      const int i = min(ib + f, i1 - 1);   // a short last group repeats its last frame (the same values are rewritten)
      const long long ro = xbase + (long long)(BWD ? (Ti - 1 - i) : i) * st;
#pragma unroll
      for (int u = 0; u < VPL; u++) {
        const int v = lane + 32 * u;
        xn[f][u] = v < V ? fz_load_logit(logits, dtype, ro + v) : -INFINITY;
      }
    }
  };
  load(i0);
#pragma unroll 1
  for (int ib = i0; ib < i1; ib += F) {
    float xv[F][VPL], m[F], s[F];
    bool nan[F];
#pragma unroll
    for (int f = 0; f < F; f++)
#pragma unroll
      for (int u = 0; u < VPL; u++) xv[f][u] = xn[f][u];
    if (ib + F < i1) load(ib + F);
#pragma unroll
    for (int f = 0; f < F; f++) {
      float mm = xv[f][0];
      bool nn = xv[f][0] != xv[f][0];
#pragma unroll
      for (int u = 1; u < VPL; u++) { mm = fmaxf(mm, xv[f][u]); nn |= xv[f][u] != xv[f][u]; }
      m[f] = mm; nan[f] = nn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int f = 0; f < F; f++) m[f] = fmaxf(m[f], __shfl_xor_sync(FULL, m[f], o));
#pragma unroll
    for (int f = 0; f < F; f++) {
      float ss = 0.f;
#pragma unroll
      for (int u = 0; u < VPL; u++) ss += expf(xv[f][u] - m[f]);   // exp(-inf) = 0 for the padding lanes
      s[f] = ss;
      nan[f] = __any_sync(FULL, nan[f]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int f = 0; f < F; f++) s[f] += __shfl_xor_sync(FULL, s[f], o);
#pragma unroll
    for (int f = 0; f < F; f++) {
      const int i = min(ib + f, i1 - 1);
      double* Erow = E + (size_t)(i & Rm) * es;
      float mf = m[f], ls = logf(s[f]);
      if (nan[f]) { mf = NAN; ls = NAN; }   // a NaN in the row poisons it, as in torch's log_softmax
#pragma unroll
      for (int u = 0; u < VPL; u++) {
        const int v = lane + 32 * u;
        if (v < V) Erow[v] = fz_emission(xv[f][u], mf, ls, from_logits);
      }
      if (lane == 0) {
        const double mls = (double)mf + (double)ls;
        Erow[V] = 0.0;                                     // the column padding cells read
        Erow[V + 1] = from_logits ? 1.0 : exp(mls);        // turns the emission back into exp(x)
        if (ib + f < i1) lse += mls;
      }
    }
  }
  return lse;
}

// dense mode, fp64 input: one warp per frame, everything in double
__device__ __noinline__ double fz_produce_dense_f64(const FzParams& p, const FzView& sv, long long xbase, int Ti,
                                                    int i0, int lane, bool BWD) {
  const FzLayout& L = p.L;
  const int V = p.V;
  const int i1 = min(i0 + L.PB, Ti);
  double lse = 0.0;
  for (int i = i0; i < i1; i++) {
    const double* xr = reinterpret_cast<const double*>(p.logits) + xbase + (long long)(BWD ? (Ti - 1 - i) : i) * p.st;
    double* Erow = sv.E + (size_t)(i & (L.R - 1)) * L.es;
    double m = -INFINITY, s = 0.0;
    bool nan = false;
    for (int v = lane; v < V; v += 32) { const double x = __ldg(xr + v); nan |= x != x; m = x > m ? x : m; }
    m = warp_max(m);
    for (int v = lane; v < V; v += 32) s += exp(__ldg(xr + v) - m);
    s = warp_sum(s);
    nan = __any_sync(0xffffffffu, nan);
    double ls = log(s);
    if (nan) { m = NAN; ls = NAN; }
    for (int v = lane; v < V; v += 32) Erow[v] = exp((__ldg(xr + v) - m) - ls);   // exp(-inf) = 0
    if (lane == 0) {
      Erow[V] = 0.0;
      Erow[V + 1] = p.from_logits ? 1.0 : exp(m + ls);
      lse += m + ls;
    }
  }
  return lse;
}

// gather mode: K1 left every frame's emissions compact in global memory ([label 0 .. label Lmax-1 | blank], doubles:
// ctc_rowstats.cu) while it had the row in hand, so staging a block is a coalesced copy of (L_i + 1) doubles per
// frame, all the block's frames in flight per lane: E row = [label 0 .. label Li-1 | zeros ... | blank at column
// 64*NB] in THIS sweep's label order (the backward sweep runs over the reversed labels).
__device__ __forceinline__ double fz_produce_gather(const FzParams& p, const FzView& sv, int b, long long xbase, int Ti, int Li,
                                                    int i0, int lane, bool BWD) {
  constexpr int F = 8;   // == PB in gather mode
  const FzLayout& L = p.L;
  const int bcol = 64 * L.NB;
  const int nf = min(F, Ti - i0);
  double lse = 0.0;
  const double* er[F];
#pragma unroll
  for (int f = 0; f < F; f++) {
    const int i = i0 + min(f, nf - 1), t = BWD ? (Ti - 1 - i) : i;
    const long long row = (long long)b * p.T + t;
    er[f] = p.emis + row * p.emis_stride;
    if (lane == 0 && f < nf && !p.from_logits) {   // log-prob input: the row normalisers go back into the loss
      if (p.dtype == E2E_F64) lse += reinterpret_cast<const double*>(p.stats)[2 * row] + reinterpret_cast<const double*>(p.stats)[2 * row + 1];
      else { const float2 st = __ldg(reinterpret_cast<const float2*>(p.stats) + row); lse += (double)st.x + (double)st.y; }
    }
  }
  for (int k = lane; k <= Li; k += 32) {
    const bool isb = k == Li;
    const int src = isb ? p.Lmax : (BWD ? Li - 1 - k : k);
    const int col = isb ? bcol : k;
    double x[F];
#pragma unroll
    for (int f = 0; f < F; f++) x[f] = __ldcg(er[f] + src);
#pragma unroll
    for (int f = 0; f < F; f++) {
      const int i = i0 + min(f, nf - 1);
      sv.E[(size_t)(i & (L.R - 1)) * L.es + col] = x[f];
    }
  }
  (void)xbase;
  return lse;
}

template <bool GATHER>
__device__ void fz_producer(const FzParams& p, const FzView& sv, int b, int Ti, int Li, int pw, int lane, bool BWD) {
  const FzLayout& L = p.L;
  const int PB = L.PB;
  const int nblocks = (Ti + PB - 1) / PB;
  const long long xbase = (long long)b * p.sb;
  double lse = 0.0;
  if (!GATHER && p.dtype != E2E_F64 && p.V <= 32) {
    // alphabets of <= 32 symbols, 32 / 16-bit logits: the lane-per-frame row producer (its own pass loop)
    lse = fz_producer_dense(p, sv, b, Ti, pw, lane, BWD);
    if (lane == 0) sv.ctl->lse[pw] = lse;
    return;
  }
  long long dbg_wait = 0, dbg_work = 0;
  for (int bi = pw; bi < nblocks; bi += L.NP) {
    const int need = bi * PB + PB - L.R;   // frames below `need` must have left the ring
    const long long dbg_a = FZ_CLK();
    if (need > 0) {   // write-after-read on the ring rows: plain progress words
      // dense mode: the combiners read the emission rows too (the softmax term of the gradient); gather mode: only the lattice does
      while (sv.ctl->lat_prog < need || (!GATHER && fz_min_done(sv.ctl->comb_done, L.NC) < need)) { if (L.nap) __nanosleep(L.nap); }
    }
    const long long dbg_b = FZ_CLK();
    if (GATHER) lse += fz_produce_gather(p, sv, b, xbase, Ti, Li, bi * PB, lane, BWD);
    else if (p.dtype == E2E_F64) lse += fz_produce_dense_f64(p, sv, xbase, Ti, bi * PB, lane, BWD);
    else {   // 33 .. 128 symbols: one warp per frame (a lane-per-frame pass over rows this long thrashes L1: 176 against 103 us on a c3 bucket)
      const int i0 = bi * PB, i1 = min(i0 + PB, Ti);
      if (p.V <= 64) lse += fz_produce_dense<2>(p.logits, p.dtype, p.st, p.V, p.from_logits, sv.E, L.R - 1, L.es, xbase, Ti, i0, i1, lane, BWD);
      else lse += fz_produce_dense<4>(p.logits, p.dtype, p.st, p.V, p.from_logits, sv.E, L.R - 1, L.es, xbase, Ti, i0, i1, lane, BWD);
    }
    __syncwarp();
    if (lane == 0) fz_mbar_arrive(&sv.ctl->fullE[bi & L.neb_mask]);
    dbg_wait += dbg_b - dbg_a; dbg_work += FZ_CLK() - dbg_b;
  }
#ifdef FZ_DBG
  if (blockIdx.x < 4 && lane == 0 && pw < 2) { g_fz_dbg[64 + blockIdx.x * 4 + pw * 2] = dbg_wait; g_fz_dbg[64 + blockIdx.x * 4 + pw * 2 + 1] = dbg_work; }
#endif
  (void)dbg_wait; (void)dbg_work;
  if (lane == 0) sv.ctl->lse[pw] = lse;   // lane 0 summed its frames in order: deterministic loss for log-prob input
}

// ---- the lattice warp -----------------------------------------------------------------------------
template <int NBU, bool BWD, bool GATHER, bool SCALER>
__device__ __forceinline__ void fz_lattice(const FzParams& p, const FzView& sv, int Ti, int Li, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = kFzG;
  using GM = FzGeom<NBU>;
  constexpr int NBP = GM::NBP, P = GM::P;
  const FzLayout& L = p.L;
  const int S = 2 * Li + 1;
  const int g0 = lane * NBU;           // first block of this lane, in lattice order
  const int CF = L.CF;

  // per-lane lattice constants: emission columns of the lane's label cells, repeat-skip permission bits
  int ecol[NBU][2];
  unsigned skipm = 0;
#pragma unroll
  for (int j = 0; j < NBU; j++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int li = 2 * (g0 + j) + h;
      const bool lv = li < Li;
      const int lab = sv.lab[li];        // padded with blank past L_i
      ecol[j][h] = (lv ? lab : p.V) * 8; // byte offset; column V is the zero column
      const bool sk = lv && li >= 1 && lab != p.blank && lab != sv.lab[li - 1];
      skipm |= sk ? (1u << (2 * j + h)) : 0u;
    }
  const int bcol = GATHER ? 64 * L.NB : p.blank;

  double x[NBU][4];
  int e[NBU], en_next[NBU];
  double fb[NBU];
#pragma unroll
  for (int j = 0; j < NBU; j++) {
#pragma unroll
    for (int c = 0; c < 4; c++) x[j][c] = (g0 + j == 0 && c == 0) ? 1.0 : 0.0;   // a virtual frame before the first: all mass on cell 0
    e[j] = 0; en_next[j] = 0;
    fb[j] = (j == 0 && lane == 0) ? 0.0 : 1.0;
  }
  int nb_en0 = 0;

  unsigned char* const vbase = sv.val;
  const int lane16 = lane * 16;
  long long dbg_we = 0, dbg_wc = 0;
  const long long dbg_t0 = FZ_CLK();
  const int pb_log2 = L.pb_log2, RVm = L.RV - 1, Rm = L.R - 1, es = L.es, vframe = L.vframe;
  const int neb_mask = L.neb_mask, neb_log2 = L.neb_log2;
  // before the first frame of a chunk: its emission block is in the ring (data: mbarrier), and the val-ring
  // slots the chunk overwrites have been drained by the combiners (reuse: plain progress words)
  auto chunk_wait = [&](int i0) {
    const long long t0 = FZ_CLK();
    if ((i0 & ((1 << pb_log2) - 1)) == 0) {
      const int bi = i0 >> pb_log2;
      fz_mbar_wait(&sv.ctl->fullE[bi & neb_mask], (bi >> neb_log2) & 1);
    }
    const long long t1 = FZ_CLK();
#ifndef FZ_EXP_NOCHUNK
    const int needv = i0 + CF - L.RV;        // frames below this must have left the val ring
    if (needv > 0) { while (fz_min_done(sv.ctl->comb_done, L.NC) < needv) {} }
#endif
    dbg_we += t1 - t0; dbg_wc += FZ_CLK() - t1;
  };
  // emissions of one frame: the blank column and the lane's 2*NBU label columns
  auto load_em = [&](const double* Erow, double& mb, double (&ml)[NBU][2]) {
    mb = Erow[bcol];
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      if (GATHER) {
        ml[j][0] = Erow[2 * (g0 + j)];
        ml[j][1] = Erow[2 * (g0 + j) + 1];
      } else {
        ml[j][0] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(Erow) + ecol[j][0]);
        ml[j][1] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(Erow) + ecol[j][1]);
      }
    }
  };
  // ---- snapshot: where every block's scale should move (applied at the next group's first frame) ----
  // en_g = max_{g' <= g} (own_g' + D_g') - D_g with D_g = kFzSlack * #{blocks with mass <= g}, in lattice order
  // g = lane*NBU + j: a block WITH mass takes its own maximum but never less than the previous massive block's
  // new exponent - kFzSlack (incoming cells are scaled by at most 2^kFzSlack per hop: no overflow when a large
  // mass follows a tiny front trickle); a block WITHOUT mass takes the nearest massive block's exponent exactly,
  // so the front always runs into a scale that is at most a few frames stale (counting blocks instead of massive
  // blocks inflates the cells the front reaches by 2^kFzSlack per block and overflows alpha*beta: found by the fuzz).
  // In-lane running maximum + one exclusive warp scan of the lane totals.
  auto snapshot = [&]() {
    const unsigned lt_mask = (1u << lane) - 1u;
    int own[NBU]; bool alive[NBU];
    int before = 0;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      const int mhi = max(max(__double2hiint(x[j][0]), __double2hiint(x[j][1])), max(__double2hiint(x[j][2]), __double2hiint(x[j][3])));
      alive[j] = mhi != 0;
      own[j] = e[j] + ((mhi >> 20) - 1023);
      before += __popc(__ballot_sync(FULL, alive[j]) & lt_mask);
    }
    int Dj[NBU], vin[NBU];
    int c = before, run = 2 * kNegExp;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      c += alive[j] ? 1 : 0;
      Dj[j] = kFzSlack * c;
      const int v = alive[j] ? own[j] + Dj[j] : 2 * kNegExp;
      run = max(run, v);
      vin[j] = run;
    }
    int tot = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, tot, d);
      if (lane >= d) tot = max(tot, t);
    }
    int excl = __shfl_up_sync(FULL, tot, 1);
    if (lane == 0) excl = 2 * kNegExp;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      const int pre = max(excl, vin[j]);
      en_next[j] = pre < kNegExp ? e[j] : pre - Dj[j];   // nothing with mass up to here: keep the scale
    }
    nb_en0 = __shfl_up_sync(FULL, en_next[NBU - 1], 1);
  };

  // One group of G = 4 frames as straight-line code (an in-order warp needs long basic blocks: the same loop with
  // one frame per iteration runs 1.6x slower, profiles/microbench/lattice_loop.cu).  The snapshot is taken before
  // the third frame and applied at the next group's first frame (en_next == e until the first snapshot, so the
  // first apply is the identity).  The next frame's emissions are loaded BEFORE this frame's val-ring stores: the
  // compiler cannot hoist shared-memory loads above possibly-aliasing stores by itself.
  double mb_n, ml_n[NBU][2];
  auto group = [&](auto full_c, int i0) {
    constexpr bool FULLG = decltype(full_c)::value;
    const double* const Erow0 = sv.E + (size_t)(i0 & Rm) * es;          // R and RV are multiples of G: no wrap inside a group
    unsigned char* const vf0 = vbase + (size_t)(i0 & RVm) * vframe;
#pragma unroll
    for (int k = 0; k < G; k++) {
      const int i = i0 + k;
      if (FULLG || i < Ti) {
        const double mb = mb_n;
        double ml[NBU][2];
#pragma unroll
        for (int j = 0; j < NBU; j++) { ml[j][0] = ml_n[j][0]; ml[j][1] = ml_n[j][1]; }
        if (k + 1 < G) {
          if (FULLG || i + 1 < Ti) load_em(Erow0 + (size_t)(k + 1) * es, mb_n, ml_n);
        } else if (i + 1 < Ti) {
          if (((i + 1) & (CF - 1)) == 0) chunk_wait(i + 1);
          load_em(sv.E + (size_t)((i + 1) & Rm) * es, mb_n, ml_n);
        }
        if (!SCALER && k == 2) snapshot();
        // re-centring: every fourth frame the blocks move to new exponents; the factors come from the scaler warp
        // (update n = i/4 - 1, computed from frame 4n+1 while frames 4n+2, 4n+3 were swept) or from the lane's own
        // snapshot (large lattices: no spare warp); they are folded into this frame's emission multipliers
        const bool apply = k == 0 && (!SCALER || i0 > 0);
        double mbj[NBU];
#pragma unroll
        for (int j = 0; j < NBU; j++) mbj[j] = mb;
        double fbn[NBU];
        if (apply) {
          if (SCALER) {
            const int n = (i0 >> 2) - 1;
            fz_mbar_wait(&sv.ctl->sc_full[n & 1], (n >> 1) & 1);
            const double* tf = sv.sc + (size_t)(n & 1) * (5 * GM::G / 2) + g0;   // slot: f[GT], fb[GT] doubles, en[GT] ints
            const double* tb = tf + GM::G;
            const int* te = reinterpret_cast<const int*>(tf - g0 + 2 * GM::G) + g0;
#pragma unroll
            for (int j = 0; j < NBU; j++) {
              const double f = tf[j];
              fbn[j] = tb[j];
              en_next[j] = te[j];
              mbj[j] *= f; ml[j][0] *= f; ml[j][1] *= f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < NBU; j++) {
              const double f = pow2i(e[j] - en_next[j]);
              mbj[j] *= f; ml[j][0] *= f; ml[j][1] *= f;
              fbn[j] = j > 0 ? pow2i(en_next[j - 1] - en_next[j]) : (lane == 0 ? 0.0 : pow2i(nb_en0 - en_next[0]));
            }
          }
        }
        // boundary cell from the previous lane (scaled by fb of the OLD exponents)
        const double top = __shfl_up_sync(FULL, x[NBU - 1][3], 1);
        unsigned char* const vf = vf0 + (size_t)k * vframe;
        // blocks top down: block j reads block j-1's last cell of the previous frame before it is overwritten
#pragma unroll
        for (int j = NBU - 1; j >= 0; j--) {
          const double in = j > 0 ? x[j - 1][3] : top;
          const double x0 = x[j][0], x1 = x[j][1], x2 = x[j][2], x3 = x[j][3];
          double s3 = x3 + x2;
          if (skipm & (1u << (2 * j + 1))) s3 += x1;
          const double s2 = x2 + x1;
          double s1 = x1 + x0;
          if (skipm & (1u << (2 * j))) s1 = fma(in, fb[j], s1);
          const double s0 = fma(in, fb[j], x0);
          x[j][3] = s3 * ml[j][1];
          x[j][2] = s2 * mbj[j];
          x[j][1] = s1 * ml[j][0];
          x[j][0] = s0 * mbj[j];
          // drop the frame into the val ring: alpha with its emission (forward), beta before its emission (backward)
          uint4 wd;
          if (BWD) wd = make_uint4((uint32_t)__double2hiint(s0), (uint32_t)__double2hiint(s1), (uint32_t)__double2hiint(s2), (uint32_t)__double2hiint(s3));
          else wd = make_uint4((uint32_t)__double2hiint(x[j][0]), (uint32_t)__double2hiint(x[j][1]), (uint32_t)__double2hiint(x[j][2]), (uint32_t)__double2hiint(x[j][3]));
          *reinterpret_cast<uint4*>(vf + j * P + lane16) = wd;
        }
        {
          int* const ve = reinterpret_cast<int*>(vf + NBU * P) + lane * NBP;
          const bool newe = !BWD && apply;
          int ev[NBP];
#pragma unroll
          for (int j = 0; j < NBP; j++) ev[j] = j < NBU ? (newe ? en_next[j] : e[j]) : 0;
#pragma unroll
          for (int u = 0; u < NBP / 4; u++) reinterpret_cast<int4*>(ve)[u] = make_int4(ev[4 * u], ev[4 * u + 1], ev[4 * u + 2], ev[4 * u + 3]);
        }
        if (apply) {
#pragma unroll
          for (int j = 0; j < NBU; j++) { fb[j] = fbn[j]; e[j] = en_next[j]; }
        }
        // publish the frame to its combiner (predicated: no branch); once per group tell the producers how far the sweep is
#ifndef FZ_EXP_NOPUB
        __syncwarp();
        fz_mbar_arrive_if(&sv.ctl->full[(i0 & RVm) + k], lane == 0);
#endif
        if (k == G - 1 && lane == 0) sv.ctl->lat_prog = i + 1;
      }
    }
  };

  chunk_wait(0);
  load_em(sv.E, mb_n, ml_n);
  int i0 = 0;
  for (; i0 + G <= Ti; i0 += G) group(std::true_type{}, i0);
  if (i0 < Ti) group(std::false_type{}, i0);
#ifdef FZ_DBG
  if (blockIdx.x < 4 && lane == 0) {
    long long* d = g_fz_dbg + blockIdx.x * 8;
    d[0] = FZ_CLK() - dbg_t0; d[1] = dbg_we; d[2] = dbg_wc; d[3] = Ti; d[4] = NBU;
  }
#endif
  (void)dbg_we; (void)dbg_wc; (void)dbg_t0;
  // exit cells S-1 and S-2 of the last frame: Z = their sum (ctc_loss.cpp:63-70)
#pragma unroll
  for (int j = 0; j < NBU; j++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int m = 4 * (g0 + j) + c;
      if (m == S - 1) { sv.ctl->tail_x[0] = x[j][c]; sv.ctl->tail_e[0] = e[j]; sv.ctl->tail_on[0] = 1; }
      if (m == S - 2) { sv.ctl->tail_x[1] = x[j][c]; sv.ctl->tail_e[1] = e[j]; sv.ctl->tail_on[1] = 1; }
    }
}

// ---- scaler warp: the block-exponent snapshot, off the lattice warp's dependent chain ---------------------
// Every fourth frame (4n+1) the scaler reads the frame's val-ring row -- the top words of the cells and their block
// exponents, exactly what the snapshot needs -- and computes where every block's scale should move:
//   en_g = max_{g' <= g} (own_g' + D_g') - D_g,  D_g = kFzSlack * #{blocks with mass <= g}, in lattice order:
// a block WITH mass takes its own maximum but never less than the previous massive block's new exponent - kFzSlack
// (incoming cells are scaled by at most 2^kFzSlack per hop: no overflow when a large mass follows a tiny front
// trickle); a block WITHOUT mass takes the nearest massive block's exponent exactly, so the front always runs into
// a scale that is at most a few frames stale.  It leaves the factors f = 2^(e - en), fb = 2^(en[g-1] - en[g]) and
// en in the scale table; the lattice warp folds them into the emission multipliers of frame 4n+4.  Values drift for
// at most 6 frames between a snapshot and the next re-centring: 6 x 149 bits (the smallest fp32 emission) < 1022.
template <int NBU>
__device__ __forceinline__ void fz_scaler(const FzParams& p, const FzView& sv, int Ti, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  using GM = FzGeom<NBU>;
  constexpr int NBP = GM::NBP, P = GM::P, GT = GM::G;
  const FzLayout& L = p.L;
  const int rv_mask = L.RV - 1, rv_log2 = L.rv_log2;
  const int g0 = lane * NBU;
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll 1
  for (int n = 0; 4 * n + 4 < Ti; n++) {
    const int a = 4 * n + 1;
    fz_mbar_wait(&sv.ctl->full[a & rv_mask], (a >> rv_log2) & 1);
    const unsigned char* vf = sv.val + (size_t)(a & rv_mask) * L.vframe;
    const int* ve = reinterpret_cast<const int*>(vf + NBU * P) + lane * NBP;
    int e[NBU], own[NBU]; bool alive[NBU];
    int before = 0;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      const uint4 wd = *reinterpret_cast<const uint4*>(vf + j * P + lane * 16);
      e[j] = ve[j];
      const int mhi = max(max((int)wd.x, (int)wd.y), max((int)wd.z, (int)wd.w));
      alive[j] = mhi != 0;
      own[j] = e[j] + ((mhi >> 20) - 1023);
      before += __popc(__ballot_sync(FULL, alive[j]) & lt_mask);
    }
    int Dj[NBU], vin[NBU], en[NBU];
    int c = before, run = 2 * kNegExp;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      c += alive[j] ? 1 : 0;
      Dj[j] = kFzSlack * c;
      const int v = alive[j] ? own[j] + Dj[j] : 2 * kNegExp;
      run = max(run, v);
      vin[j] = run;
    }
    int tot = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, tot, d);
      if (lane >= d) tot = max(tot, t);
    }
    int excl = __shfl_up_sync(FULL, tot, 1);
    if (lane == 0) excl = 2 * kNegExp;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      const int pre = max(excl, vin[j]);
      en[j] = pre < kNegExp ? e[j] : pre - Dj[j];   // nothing with mass up to here: keep the scale
    }
    const int nb_en0 = __shfl_up_sync(FULL, en[NBU - 1], 1);
    double* tf = sv.sc + (size_t)(n & 1) * (5 * GT / 2) + g0;
    double* tb = tf + GT;
    int* te = reinterpret_cast<int*>(tf - g0 + 2 * GT) + g0;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      tf[j] = pow2i(e[j] - en[j]);
      tb[j] = j > 0 ? pow2i(en[j - 1] - en[j]) : (lane == 0 ? 0.0 : pow2i(nb_en0 - en[0]));
      te[j] = en[j];
    }
    __syncwarp();
    if (lane == 0) fz_mbar_arrive(&sv.ctl->sc_full[n & 1]);
  }
}

// ---- combiner warps ---------------------------------------------------------------------------------
template <int NBU, bool GATHER>
__device__ __forceinline__ void fz_combiner(const FzParams& p, const FzView& sv, int b, int Ti, int Li, int q, int lane, bool BWD) {
  constexpr unsigned FULL = 0xffffffffu;
  using GM = FzGeom<NBU>;
  constexpr int NBP = GM::NBP, P = GM::P, GT = GM::G;
  constexpr int ROWW = 5 * GT;                 // words of a stash row in use: 4*GT cell words + GT exponents
  const FzLayout& L = p.L;
  const int NC = L.NC, PF = L.PF;
  const int S = 2 * Li + 1;
  const int tm = Ti / 2;
  const int nstore = BWD ? (Ti - tm) : tm;          // frames this sweep stores; the rest it combines
  const int nstore_peer = Ti - nstore;
  auto frame_t = [&](int i) { return BWD ? (Ti - 1 - i) : i; };
  uint32_t* const stash_b = p.stash + (size_t)b * p.T * p.roww;
  uint32_t* const stage = sv.stage + (size_t)q * PF * (L.srow >> 2);
  uint32_t* const acc = reinterpret_cast<uint32_t*>(sv.post) + (size_t)q * L.prow;   // dense: per-symbol posterior sums of one frame (2^-31 fixed point)

  const int rv_mask = L.RV - 1, rv_log2 = L.rv_log2;
  auto wait_val = [&](int i) { fz_mbar_wait(&sv.ctl->full[i & rv_mask], (i >> rv_log2) & 1); };
  auto done = [&](int i) {
    __syncwarp();
    if (lane == 0) sv.ctl->comb_done[q] = i + NC;
  };
  auto prefetch = [&](int i2, int slot) {   // the other sweep's stored row of my frame i2 -> staging slot
    if (i2 < Ti) {
      uint32_t* dst = stage + (size_t)slot * (L.srow >> 2);
      const uint32_t* src = stash_b + (size_t)frame_t(i2) * p.roww;
      for (int u = lane; u < ROWW / 4; u += 32) fz_cp_async_cg16(dst + 4 * u, src + 4 * u);
    }
  };

  int i = q;
  // ---- first half: val ring -> global stash (lattice order) ----
  for (; i < nstore; i += NC) {
    wait_val(i);
#ifdef FZ_DBG
    if (L.dbg & 4) { done(i); continue; }
#endif
    const unsigned char* vf = sv.val + (size_t)(i & rv_mask) * L.vframe;
    const int* ve = reinterpret_cast<const int*>(vf + NBU * P);
    uint32_t* row = stash_b + (size_t)frame_t(i) * p.roww;
#pragma unroll
    for (int u = 0; u < NBU; u++) {
      const int g = lane + 32 * u, l = g / NBU, j = g - l * NBU;
      reinterpret_cast<uint4*>(row)[g] = *reinterpret_cast<const uint4*>(vf + j * P + l * 16);
      row[4 * GT + g] = (uint32_t)ve[l * NBP + j];
    }
    if (i + NC >= nstore) {   // my last stored row: publish to the other CTA
      __syncwarp();
      if (lane == 0) { __threadfence(); atomicAdd(p.meet + 2 * b + (BWD ? 1 : 0), 1); }
    }
    done(i);
  }
  if (i >= Ti) return;
#ifdef FZ_DBG
  if (L.dbg & 4) { for (; i < Ti; i += NC) { wait_val(i); done(i); } return; }   // timing experiment: helpers idle
#endif
  // ---- meet: the other sweep's stored rows must be visible ----
  {
    const int want = min(NC, nstore_peer);
    const int* flag = p.meet + 2 * b + (BWD ? 0 : 1);
    while (fz_ld_acquire_gpu(flag) < want) __nanosleep(64);
  }
  for (int u = 0; u < PF; u++) { prefetch(i + u * NC, u); fz_cp_async_commit(); }

  // My cell m = 4g + c is the other sweep's cell mp = S-1-m.  (S-1) = 4*qq + r with r in {0, 2} (S is odd):
  //   r == 0: c = 0 pairs with word 0 of the other's block qq-g; c = 1,2,3 with words 3,2,1 of block qq-g-1;
  //   r == 2: c = 0,1,2 pair with words 2,1,0 of block qq-g; c = 3 with word 3 of block qq-g-1.
  // Blocks below 0 are past the other sweep's lattice: zero.
  const int qq = (S - 1) >> 2, r = (S - 1) & 3;
  const float sc = (float)p.scale;
  const int srow_w = L.srow >> 2;
  // products alpha*beta of my block g (unscaled) + the exponent sums of its two partner blocks (branch-free)
  auto block = [&](const unsigned char* vf, const int* ve, const uint32_t* orow, int g, double (&pr)[4], int& EA, int& EB) {
    const int l = g / NBU, j = g - l * NBU;
    const uint4 mine = *reinterpret_cast<const uint4*>(vf + j * P + l * 16);
    const int em = ve[l * NBP + j];
    const int ga = qq - g, gb = ga - 1;
    uint4 oa = reinterpret_cast<const uint4*>(orow)[max(ga, 0)];
    uint4 ob = reinterpret_cast<const uint4*>(orow)[max(gb, 0)];
    const int ea = (int)orow[4 * GT + max(ga, 0)], eb = (int)orow[4 * GT + max(gb, 0)];
    if (ga < 0) oa = make_uint4(0u, 0u, 0u, 0u);
    if (gb < 0) ob = make_uint4(0u, 0u, 0u, 0u);
    EA = em + ea; EB = em + eb;
    const uint32_t w0 = r == 0 ? oa.x : oa.z, w1 = r == 0 ? ob.w : oa.y, w2 = r == 0 ? ob.z : oa.x, w3 = r == 0 ? ob.y : ob.w;
    pr[0] = fz_hi2d(mine.x) * fz_hi2d(w0);
    pr[1] = fz_hi2d(mine.y) * fz_hi2d(w1);
    pr[2] = fz_hi2d(mine.z) * fz_hi2d(w2);
    pr[3] = fz_hi2d(mine.w) * fz_hi2d(w3);
  };
  // which partner exponent cell c uses: r == 0: c=0 -> A, else B;  r == 2: c<3 -> A, c=3 -> B
  auto usesA = [&](int c) { return r == 0 ? c == 0 : c < 3; };

  // ---- Z = sum_s alpha(t,s) * beta(t,s), the same for every frame: taken once per combiner warp, at its first frame.
  // 2^31 / Z = cz * 2^kz with cz in [1,2): the power of two moves into Ez, so the per-cell scale 2^(El - Ez) * cz
  // stays finite whatever stale exponent a massless block carries (0 * finite = 0).  Z == 0 or NaN: the lattice
  // tail flags the utterance and the block is overwritten with NaN.
  double cz;
  int Ez;
  {
    wait_val(i);
    if (PF == 2) fz_cp_async_wait<1>(); else fz_cp_async_wait<3>();
    __syncwarp();
    const uint32_t* orow = stage;
    const unsigned char* vf = sv.val + (size_t)(i & rv_mask) * L.vframe;
    const int* ve = reinterpret_cast<const int*>(vf + NBU * P);
    int emax = 4 * kNegExp;
#pragma unroll 1
    for (int u = 0; u < NBU; u++) {
      double pr[4]; int EA, EB;
      block(vf, ve, orow, lane + 32 * u, pr, EA, EB);
#pragma unroll
      for (int c = 0; c < 4; c++) if (pr[c] > 0.0) emax = max(emax, usesA(c) ? EA : EB);
    }
    emax = warp_max_int(emax);
    double tot = 0.0;
#pragma unroll 1
    for (int u = 0; u < NBU; u++) {
      double pr[4]; int EA, EB;
      block(vf, ve, orow, lane + 32 * u, pr, EA, EB);
      const double fA = pow2i(EA - emax), fB = pow2i(EB - emax);
#pragma unroll
      for (int c = 0; c < 4; c++) if (pr[c] > 0.0) tot += pr[c] * (usesA(c) ? fA : fB);
    }
    tot = warp_sum(tot);
    const double rz = 2147483648.0 / tot;
    const int kz = ((__double2hiint(rz) >> 20) & 0x7ff) - 1023;
    cz = __hiloint2double((__double2hiint(rz) & 0x800fffff) | 0x3ff00000, __double2loint(rz));
    Ez = emax - kz;
  }

  // ---- second half: one frame per iteration; the block loop is ROLLED (instruction-cache footprint: the
  // combiner loop shares the SM's 32 KB L1.5 with the lattice and producer loops)
  for (int k = 0; i < Ti; i += NC, ++k) {
    if (k > 0) {
      wait_val(i);
      if (PF == 2) fz_cp_async_wait<1>(); else fz_cp_async_wait<3>();
      __syncwarp();
    }
    const uint32_t* orow = stage + (size_t)(k % PF) * srow_w;
    const unsigned char* vf = sv.val + (size_t)(i & rv_mask) * L.vframe;
    const int* ve = reinterpret_cast<const int*>(vf + NBU * P);
    // posteriors scaled by 2^31: blank cells are summed as fixed point (one warp-wide integer add); label cells are
    // added to their symbol's accumulator with integer shared-memory atomics (dense: integer adds commute, so the
    // gradient is bitwise reproducible) or go straight to the compact global row (gather)
    uint32_t bsum = 0u;
    const int t = frame_t(i);
    float* const prow = GATHER ? p.post + ((size_t)b * p.T + t) * (size_t)p.post_stride : nullptr;
#pragma unroll 1
    for (int u = 0; u < NBU; u++) {
      const int g = lane + 32 * u;
      double pr[4]; int EA, EB;
      block(vf, ve, orow, g, pr, EA, EB);
      const double sA = pow2i(EA - Ez) * cz, sB = pow2i(EB - Ez) * cz;
      const double p0 = pr[0] * (usesA(0) ? sA : sB), p1 = pr[1] * (usesA(1) ? sA : sB);
      const double p2 = pr[2] * (usesA(2) ? sA : sB), p3 = pr[3] * (usesA(3) ? sA : sB);
      bsum += __double2uint_rn(p0) + __double2uint_rn(p2);
      const int li0 = 2 * g, li1 = 2 * g + 1;
      if (GATHER) {   // the FORWARD label index in the compact row
        if (li0 < Li) prow[BWD ? Li - 1 - li0 : li0] = (float)p1 * (1.f / 2147483648.f);
        if (li1 < Li) prow[BWD ? Li - 1 - li1 : li1] = (float)p3 * (1.f / 2147483648.f);
      } else {        // the label's symbol (lab[] is padded with blank past L_i: those cells carry zero)
        atomicAdd(acc + sv.lab[li0], __double2uint_rn(p1));
        atomicAdd(acc + sv.lab[li1], __double2uint_rn(p3));
      }
    }
    const uint32_t qb = __reduce_add_sync(FULL, bsum);
    __syncwarp();
    if (GATHER) {
      if (lane == 0) prow[p.cells / 2] = (float)qb * (1.f / 2147483648.f);
    } else {
      // gradient row: scale * (softmax - posterior); log-prob input: exp(lp) - posterior (the engine contract)
      const double* Erow = sv.E + (size_t)(i & (L.R - 1)) * L.es;
      const long long gbase = (long long)b * p.gsb + (long long)t * p.gst;
      const double rsd = Erow[p.V + 1];
      const float rs = (float)rsd;
      for (int v = lane; v < p.V; v += 32) {
        const uint32_t a = acc[v] + (v == p.blank ? qb : 0u);
        acc[v] = 0u;
        if (p.dtype == E2E_F64) {
          reinterpret_cast<double*>(p.grads)[gbase + v] = p.scale * (Erow[v] * rsd - (double)a * (1.0 / 2147483648.0));
        } else {
          const float gv = sc * ((float)Erow[v] * rs - (float)a * (1.f / 2147483648.f));
          if (p.dtype == E2E_F32) reinterpret_cast<float*>(p.grads)[gbase + v] = gv;
          else if (p.dtype == E2E_BF16) reinterpret_cast<__nv_bfloat16*>(p.grads)[gbase + v] = __float2bfloat16_rn(gv);
          else reinterpret_cast<__half*>(p.grads)[gbase + v] = __float2half_rn(gv);
        }
      }
    }
    __syncwarp();
    prefetch(i + PF * NC, k % PF);
    fz_cp_async_commit();
    done(i);
  }
  fz_cp_async_wait<0>();
}

// ---- role dispatch ------------------------------------------------------------------------------------
template <int NBU, bool GATHER>
__device__ __forceinline__ void fz_roles_nbu(const FzParams& p, const FzView& sv, int b, int Ti, int Li, int role, int idx, int lane, bool BWD) {
  constexpr bool SCALER = false;      // experiment: the scaler's chain (~600 cycles) does not fit the 2-frame lag; the lattice warp keeps the snapshot
  if (role == 0) {
    if (BWD) fz_lattice<NBU, true, GATHER, SCALER>(p, sv, Ti, Li, lane);
    else fz_lattice<NBU, false, GATHER, SCALER>(p, sv, Ti, Li, lane);
  } else if (role == 1) {
    fz_combiner<NBU, GATHER>(p, sv, b, Ti, Li, idx, lane, BWD);
  } else if (role == 4) {
    if (SCALER) fz_scaler<NBU>(p, sv, Ti, lane);
  }
}

// block rows per lane this utterance sweeps (its own 2L+1 cells, not the batch maximum), from the set the
// kernel class NB is instantiated for
template <int NB>
__device__ __forceinline__ int fz_pick_nbu(int S) {
  const int need = (S + 127) >> 7;
  if (NB <= 4) return need < 1 ? 1 : need;
  return need <= 6 ? 6 : (need <= 8 ? 8 : 10);
}

template <int NB, bool GATHER>
__device__ __forceinline__ void fz_roles(const FzParams& p, const FzView& sv, int b, int Ti, int Li, int w, int lane, bool BWD) {
  const FzLayout& L = p.L;
  const int role = L.role[w], idx = L.ridx[w];
  if (role == 2) {
    fz_producer<GATHER>(p, sv, b, Ti, Li, idx, lane, BWD);
    if (!GATHER) {
      // padding frames t >= T_i: exp(lp) for log-prob input (the engine contract, ctc_loss.cpp:105-117),
      // 0 for fused-logits input (what the reference's log_softmax backward leaves there)
      for (int r = Ti + (BWD ? 1 : 0) + 2 * idx; r < p.T; r += 2 * L.NP) {
        const long long xo = (long long)b * p.sb + (long long)r * p.st;
        const long long go = (long long)b * p.gsb + (long long)r * p.gst;
        for (int v = lane; v < p.V; v += 32) {
          double gq = 0.0;
          if (!p.from_logits) gq = p.dtype == E2E_F64 ? exp(load_as_double(p.logits, p.dtype, xo + v)) : (double)expf(load_as_float(p.logits, p.dtype, xo + v));
          store_from_double(p.grads, p.dtype, go + v, p.scale * gq);
        }
      }
    }
    return;
  }
  if (role == 3) return;
  const int nbu = fz_pick_nbu<NB>(2 * Li + 1);
  if (NB == 1) { fz_roles_nbu<1, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); }
  else if (NB == 2) {
    if (nbu == 1) fz_roles_nbu<1, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD);
    else fz_roles_nbu<2, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD);
  } else if (NB == 4) {
    switch (nbu) {
      case 1: fz_roles_nbu<1, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
      case 2: fz_roles_nbu<2, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
      case 3: fz_roles_nbu<3, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
      default: fz_roles_nbu<4, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
    }
  } else {
    switch (nbu) {
      case 6: fz_roles_nbu<6, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
      case 8: fz_roles_nbu<8, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
      default: fz_roles_nbu<10, GATHER>(p, sv, b, Ti, Li, role, idx, lane, BWD); break;
    }
  }
}

template <int NB> struct FzBounds {
  static constexpr int kThreads = NB <= 4 ? 256 : 160;
  static constexpr int kMaxRegs = NB <= 2 ? 80 : (NB <= 4 ? FZ_REGS4 : 200);   // 3 / 2 / 2 CTAs per SM
};

template <int NB, bool GATHER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FzBounds<NB>::kThreads) __maxnreg__(FzBounds<NB>::kMaxRegs)
ctc_fused_kernel(const FzParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FzView sv = fz_carve(smem_raw, p.L);
  __shared__ int pre[4];

  const int b = blockIdx.x >> 1;
  const bool bwd = (blockIdx.x & 1) != 0;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;

  const long long Ti_ll = load_index(p.in_len, p.len_is64, b);
  const long long Li_ll = load_index(p.tgt_len, p.len_is64, b);
  int bad = 0;
  if (Ti_ll < 1 || Ti_ll > p.T) bad |= kBadFrames;
  if (Li_ll < 0 || Li_ll > p.Lmax) bad |= kBadTargetLen;
  const int Ti = (int)Ti_ll, Li = bad ? 0 : (int)Li_ll;
  if (tid < 4) pre[tid] = 0;
  __syncthreads();
  int rep = 0, badlab = 0;
  // the backward CTA sweeps the reversed label sequence
  for (int i = tid; i < Li; i += blockDim.x) {
    const long long v = load_index(p.targets, p.tgt_is64, (long long)b * p.ts_b + i);
    if (v < 0 || v >= p.V) badlab = kBadLabel;
    sv.lab[bwd ? (Li - 1 - i) : i] = (int)v;
  }
  for (int i = Li + tid; i < 64 * NB + 1; i += blockDim.x) sv.lab[i] = p.blank;   // cells past the lattice carry zero mass
  __syncthreads();
  for (int i = tid + 1; i < Li; i += blockDim.x) rep += (sv.lab[i] == sv.lab[i - 1]);
  if (rep) atomicAdd(&pre[1], rep);
  if (badlab) atomicOr(&pre[0], badlab);
  __syncthreads();
  bad |= pre[0];
  rep = pre[1];
  const long long gfill_base = (long long)b * p.gsb;
  if (bad || Ti < Li + rep) {
    // out-of-range lengths / labels (undefined behaviour in the reference): NaN loss + status bits;
    // no alignment exists (T < L + repeats): loss = +inf.  Either way the gradient block is all NaN
    // (-inf - (-inf) in the reference, ctc_loss.cpp:116-117), padding rows included (gather mode: K3 writes it).
    if (tid == 0 && !bwd) {
      if (bad) atomicOr(p.status, bad);
      p.flags[b] = bad ? kFlagInvalid : kFlagInfeasible;
      store_from_double(p.losses, p.dtype, b, bad ? (double)NAN : (double)INFINITY);
    }
    if (!GATHER && p.grads != nullptr) {
      for (int r = 2 * w + (bwd ? 1 : 0); r < p.T; r += 2 * nwarps)
        for (int v = lane; v < p.V; v += 32) store_from_double(p.grads, p.dtype, gfill_base + (long long)r * p.gst + v, (double)NAN);
    }
    return;
  }
  {  // control block and (gather mode) the emission ring start at zero: label columns past L_i are never written
    uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw + p.L.off_ctl);
    for (int k = tid; k < (int)(sizeof(FzCtl) >> 2); k += blockDim.x) z[k] = 0u;
    if (GATHER) {
      uint32_t* ez = reinterpret_cast<uint32_t*>(sv.E);
      const int n = p.L.R * p.L.es * 2;
      for (int k = tid; k < n; k += blockDim.x) ez[k] = 0u;
    }
    for (int k = tid; k < p.L.NC * p.L.prow; k += blockDim.x) sv.post[k] = 0.f;
  }
  __syncthreads();
  if (tid < p.L.NC) sv.ctl->comb_done[tid] = tid;
  if (tid < 32) fz_mbar_init(&sv.ctl->full[tid], 1);
  else if (tid < 48) fz_mbar_init(&sv.ctl->fullE[tid - 32], 1);
  else if (tid < 50) fz_mbar_init(&sv.ctl->sc_full[tid - 48], 1);
  __syncthreads();

#ifdef FZ_DBG
  const long long dbg_k0 = clock64();
#endif
  fz_roles<NB, GATHER>(p, sv, b, Ti, Li, w, lane, bwd);
#ifdef FZ_DBG
  if (blockIdx.x < 4 && lane == 0 && w < 8) g_fz_dbg[32 + blockIdx.x * 8 + w] = clock64() - dbg_k0;
#endif
  __syncthreads();

  // loss = -log(alpha[S-1][T-1] + alpha[S-2][T-1]) (ctc_loss.cpp:63-70) from the live fp64 state of the
  // forward sweep (the backward sweep's exit cells give the same Z: it only needs the zero test).
  // Emissions were normalised per row, so for log-prob input the row normalisers are added back.
  if (tid == 0) {
    FzCtl* c = sv.ctl;
    const double x0 = c->tail_on[0] ? c->tail_x[0] : 0.0, x1 = c->tail_on[1] ? c->tail_x[1] : 0.0;
    const int e0 = (c->tail_on[0] && x0 > 0.0) ? c->tail_e[0] : 4 * kNegExp;
    const int e1 = (c->tail_on[1] && x1 > 0.0) ? c->tail_e[1] : 4 * kNegExp;
    const int emax = max(e0, e1);
    double z = 0.0;
    if (x0 > 0.0) z += x0 * pow2i(e0 - emax);
    if (x1 > 0.0) z += x1 * pow2i(e1 - emax);
    if (x0 != x0 || x1 != x1) z = NAN;
    double loss = INFINITY;
    if (z > 0.0) {
      loss = -(log(z) + (double)emax * 0.69314718055994530942);
      if (!p.from_logits) { for (int k = 0; k < p.L.NP; k++) loss -= c->lse[k]; }
    } else {   // no path survives (exact-zero emissions): +inf; NaN input: NaN.  NaN gradient block either way
      if (z != z) loss = NAN;
      c->zero = 1;
    }
    if (!bwd) {
      p.flags[b] = z > 0.0 ? 0 : kFlagInfeasible;
      store_from_double(p.losses, p.dtype, b, loss);
    }
  }
  __syncthreads();
  if (!GATHER && sv.ctl->zero) {
    // each CTA overwrites the rows it wrote: its combined frames and its share of the padding rows
    const int tm = Ti / 2;
    for (int r = w; r < p.T; r += nwarps) {
      const bool mine = r < Ti ? (bwd ? r < tm : r >= tm) : (((r - Ti) & 1) == (bwd ? 1 : 0));
      if (!mine) continue;
      for (int v = lane; v < p.V; v += 32) store_from_double(p.grads, p.dtype, gfill_base + (long long)r * p.gst + v, (double)NAN);
    }
  }
}

template <int NB, bool GATHER>
int launch_fused_k(const FzParams& fp, cudaStream_t s) {
  // the dynamic shared-memory opt-in is PER DEVICE: cached per device ordinal (it only ever grows)
  static int attr_smem[64];
  int dev = 0;
  E2E_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || fp.L.total > attr_smem[dev]) {
    E2E_CUDA_TRY(cudaFuncSetAttribute(ctc_fused_kernel<NB, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, fp.L.total));
    if (dev >= 0 && dev < 64) attr_smem[dev] = fp.L.total;
  }
  const unsigned threads = 32u * (unsigned)fp.L.nwarps;
  KernelTimer timer(kKernelLattice, s);
  ctc_fused_kernel<NB, GATHER><<<2u * (unsigned)fp.B, threads, (size_t)fp.L.total, s>>>(fp);
  E2E_CUDA_TRY(cudaGetLastError());
#ifdef FZ_DBG
  {
    long long h[128];
    cudaStreamSynchronize(s);
    cudaMemcpyFromSymbol(h, g_fz_dbg, sizeof(h));
    for (int c = 0; c < 4; c++) {
      fprintf(stderr, "[fz dbg] cta %d lattice loop %lld cyc (wait E %lld, wait comb %lld) T %lld NBU %lld | role cycles:", c, h[c * 8], h[c * 8 + 1], h[c * 8 + 2], h[c * 8 + 3], h[c * 8 + 4]);
      for (int w = 0; w < 8; w++) fprintf(stderr, " %lld", h[32 + c * 8 + w]);
      fprintf(stderr, " | producers (ring wait, work):");
      for (int q = 0; q < 2; q++) fprintf(stderr, " (%lld, %lld)", h[64 + c * 4 + q * 2], h[64 + c * 4 + q * 2 + 1]);
      fprintf(stderr, "\n");
    }
  }
#endif
  return E2E_OK;
}

}  // namespace
}  // namespace e2e
