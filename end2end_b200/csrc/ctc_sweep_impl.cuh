// K2 "sweep" kernel -- the alpha/beta recursion over the blank-extended label lattice with ONE WARP PER
// SWEEP, and in dense mode (small alphabets) the whole loss path (row log-softmax, lattice, gradient
// write) in one kernel.
//
// Replaces CTCLossEngine::compute_2d (src/losses/ctc_loss.cpp:15-118): extended targets (:25-31),
// alpha (:33-61), loss (:63-70), beta (:72-100), alpha+beta / gradient (:102-117); in dense mode
// also F.log_softmax (pytorch_end2end/modules/ctc_loss.py:40) and the exp(logits) term (:117).
//
// Design:
//  * CTA = one utterance = two warps.  Warp 0 sweeps alpha forward (t = 0..T-1), warp 1 sweeps beta
//    backward (t = T-1..0), concurrently.  Each stores its first half of the frames (top 32 bits of
//    every fp64 cell + the lane's block exponent) to a global stash, the two meet once in the middle
//    (one named barrier), and in its second half each multiplies its LIVE fp64 state with the other's
//    stashed row: every posterior alpha*beta/Z is produced exactly once, nothing is swept twice, and
//    only half of the lattice is ever written to memory.
//  * A lane owns K consecutive cells (even K: cells alternate blank,label) in registers; the s-1/s-2
//    transitions cross lanes with one (alpha) / two (beta) warp shuffles per frame; the repeat-label
//    skip is a per-lane bit mask.  No shared-memory traffic and no barrier on the recurrence.
//  * Arithmetic is LINEAR-domain fp64 with a per-lane block exponent (value = x * 2^e): a cell update
//    is DADD (+ predicated DADD) + DMUL, no MUFU on the chain, error ~1e-16 per step (an fp32 log-space
//    recursion fails the 1e-5 parity budget).  The block exponents are re-centred every
//    second frame from a snapshot taken two frames earlier, so the integer/shuffle work of the
//    renormalisation overlaps the fp64 pipe instead of extending the dependent chain.
//  * Emissions: per chunk of <= 32 frames the warp cp.async-stages the raw logits one chunk ahead and
//    then converts them with ONE LANE PER FRAME (no cross-lane reductions): dense mode computes the row
//    max / sum-exp / softmax of the whole row (the fused log_softmax), gather mode (large alphabets)
//    converts only blank + the utterance's labels using the row statistics of K1.
//  * Second half: the other sweep's stashed row is prefetched PF frames ahead with cp.async (L2 only),
//    posteriors are summed per symbol with integer shared-memory atomics (dense: bitwise reproducible)
//    and the gradient row softmax - posterior is written by the same warp; gather mode writes compact
//    per-label posteriors for K3.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace e2e {
namespace {

struct SweepParams {
  const void* logits; int dtype; long long sb, st;
  void* grads; long long gsb, gst; double scale;   // dense (fused) mode: gradient output
  const void* stats;                               // gather mode: row {max, logsumexp} from K1
  const void* targets; int tgt_is64; long long ts_b;
  const void* in_len; const void* tgt_len; int len_is64;
  int B, T, V, Lmax, blank, from_logits;
  void* losses;
  int* status; int* flags;
  const int* order;  // [B] utterance handled by each CTA (null: its own index)
  uint32_t* stash;   // [B*T][32][WORDS]  first-half lattice state
  float* post;       // gather mode: [B*T][post_stride] compact posteriors (labels..., blank total at cells/2)
  int dense, post_stride, cells;
  int cf;            // frames per emission chunk (<= 32)
  int es;            // emission row stride (elements, odd)
  int rawrow;        // raw staging row stride in bytes (multiple of 4, odd number of words)
  int vpad;          // dense: u32 accumulators per parity buffer
  // shared-memory layout (bytes): labels, then two identical per-warp blocks
  int off_lab, off_warp, warp_bytes, w_E, w_raw, w_stat, w_rs, w_acc, w_stage;
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BYTES>
__device__ __forceinline__ void sw_cp_async_ca(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sw_smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void sw_cp_async_cg16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sw_smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void sw_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sw_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sw_named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double sw_unpack_hi32(uint32_t h) { return __hiloint2double((int)h, 0); }

// stash prefetch depth (frames in flight per warp): deep for the narrow variants (they run with few
// warps per SM, so the L2 latency must be covered by the warp itself), shallow for the wide ones
template <int K> struct SweepCfg {
  static constexpr int H = K / 2;
  static constexpr int WORDS = (K + 1 + 3) & ~3;
  static constexpr int ROWW = 32 * WORDS;
  static constexpr int PF = K >= 24 ? kSweepPFWide : kSweepPFSmall;
};

template <int K, typename ET>
struct SweepState {
  double x[K];      // cells s0 .. s0+K-1 of the current frame, value = x * 2^e
  int e;            // block exponent
  double fb;        // 2^(neighbour's exponent - e): scale of the cells shuffled in from the neighbour lane
  int en_next;      // pending re-centring (from the last snapshot)
  double f_next, fb_next;
};

// ---- emissions of one chunk: raw logits -> probabilities, one lane per frame ---------------------
// Raw staging row f holds the frame's logits (dense: the whole row, word-aligned copy whose first element
// sits `shift` halfwords in for 16-bit types; gather: [0] blank, [1+k] label k, one 4/8-byte word each).
template <typename ET>
struct ChunkCtx {
  ET* E; unsigned char* raw; unsigned char* stat; float* rs; int* lab;
};

template <bool BWD, typename ET>
__device__ __forceinline__ void issue_chunk(const SweepParams& p, const ChunkCtx<ET>& cx, const char* lbase, int b,
                                            int Ti, int Li, int c, int lane) {
  constexpr bool F64 = sizeof(ET) == 8;
  const int i0 = c * p.cf;
  if (i0 >= Ti) return;
  const int nf = min(p.cf, Ti - i0);
  const int esz = F64 ? 8 : (p.dtype == E2E_F32 ? 4 : 2);
  const long long st_bytes = p.st * esz;
  if (p.dense) {
    for (int f = 0; f < nf; ++f) {
      const int t = BWD ? (Ti - 1 - (i0 + f)) : (i0 + f);
      const char* rowp = lbase + (long long)t * st_bytes;
      unsigned char* rslot = cx.raw + (size_t)f * p.rawrow;
      if (F64) {
        for (int v = lane; v < p.V; v += 32) sw_cp_async_ca<8>(rslot + v * 8, rowp + (size_t)v * 8);
      } else {
        const char* a = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(rowp) & ~(uintptr_t)3);
        const int nwords = (int)((rowp - a) + (size_t)p.V * esz + 3) >> 2;
        for (int w = lane; w < nwords; w += 32) sw_cp_async_ca<4>(rslot + w * 4, a + (size_t)w * 4);
      }
    }
  } else {
    for (int k = lane; k <= Li; k += 32) {
      const size_t soff = (size_t)(k == 0 ? p.blank : cx.lab[k - 1]) * esz;
      for (int f = 0; f < nf; ++f) {
        const int t = BWD ? (Ti - 1 - (i0 + f)) : (i0 + f);
        const char* src = lbase + (long long)t * st_bytes + soff;
        unsigned char* slot = cx.raw + (size_t)f * p.rawrow + (size_t)k * (F64 ? 8 : 4);
        if (F64) sw_cp_async_ca<8>(slot, src);
        else sw_cp_async_ca<4>(slot, reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3));
      }
    }
    if (lane < nf) {   // row statistics written by K1
      const int t = BWD ? (Ti - 1 - (i0 + lane)) : (i0 + lane);
      const long long row = (long long)b * p.T + t;
      if (F64) sw_cp_async_ca<16>(cx.stat + lane * 16, reinterpret_cast<const char*>(p.stats) + row * 16);
      else sw_cp_async_ca<8>(cx.stat + lane * 16, reinterpret_cast<const char*>(p.stats) + row * 8);
    }
  }
}

__device__ __forceinline__ float sw_half_to_float(uint32_t word, int dtype, bool upper) {
  const uint32_t h = upper ? (word >> 16) : (word & 0xffffu);
  return dtype == E2E_BF16 ? __uint_as_float(h << 16) : __half2float(__ushort_as_half((unsigned short)h));
}

// returns this lane's contribution to sum_t (max_t + logsumexp_t) (log-prob input only)
template <bool BWD, typename ET, int H>
__device__ __forceinline__ double convert_chunk(const SweepParams& p, const ChunkCtx<ET>& cx, const char* lbase,
                                                int Ti, int Li, int c, int lane) {
  constexpr bool F64 = sizeof(ET) == 8;
  const int i0 = c * p.cf;
  const int nf = min(p.cf, Ti - i0);
  double lse_part = 0.0;
  // Dense 32/16-bit rows: a chunk of cf < 32 frames would leave lanes idle, so 32 / cf lanes SHARE a frame, each taking a
  // contiguous range of the symbols (max and sum are combined with shuffles in a fixed order).  Other modes: one lane per frame.
  const bool split = p.dense && !F64 && p.cf < 32;
  const int lpf = split ? 32 / p.cf : 1;
  const int f = split ? (lane & (p.cf - 1)) : lane;
  const int sub = split ? lane / p.cf : 0;
  if (f >= nf) return lse_part;
  const unsigned grp = __activemask();   // every lane of a frame's group takes the same branches
  const int t = BWD ? (Ti - 1 - (i0 + f)) : (i0 + f);
  const int esz = F64 ? 8 : (p.dtype == E2E_F32 ? 4 : 2);
  const unsigned char* rslot = cx.raw + (size_t)f * p.rawrow;
  ET* Erow = cx.E + (size_t)f * p.es;
  if (p.dense) {
    if (F64) {
      const double* xr = reinterpret_cast<const double*>(rslot);
      double m = -INFINITY; bool nan = false;
      for (int v = 0; v < p.V; ++v) { const double x = xr[v]; nan |= x != x; m = x > m ? x : m; }
      double s = 0.0;
      for (int v = 0; v < p.V; ++v) s += exp(xr[v] - m);
      double ls = log(s);
      if (nan) { m = NAN; ls = NAN; }
      for (int v = 0; v < p.V; ++v) Erow[v] = (ET)exp((xr[v] - m) - ls);
      if (!p.from_logits) { cx.rs[f] = (float)0; reinterpret_cast<double*>(cx.stat)[2 * f] = exp(m + ls); lse_part = m + ls; }
    } else {
      const char* rowp = lbase + (long long)t * p.st * esz;
      const int shift = (int)((reinterpret_cast<uintptr_t>(rowp) & 3) >> 1);      // 16-bit types: first element's halfword
      const uint32_t* wr = reinterpret_cast<const uint32_t*>(rslot);
      float* Ef = reinterpret_cast<float*>(Erow);
      const bool f32 = p.dtype == E2E_F32;
      // One lane walks a whole row: four independent max / sum chains (the serial one was latency-bound: a chunk of
      // 32 frames x 96 symbols took ~23k cycles, a third of BASELINE config 3's step).  The sum keeps a FIXED order
      // -- four strided partial sums, then (s0 + s1) + (s2 + s3) -- so results stay bitwise reproducible.
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY; bool nan = false;
      // this lane's symbols [va, vb): the whole row, or its share of it (boundaries at multiples of four)
      const int va = sub == 0 ? 0 : ((p.V * sub / lpf) & ~3), vb = sub + 1 == lpf ? p.V : ((p.V * (sub + 1) / lpf) & ~3);
      const int V4 = va + ((vb - va) & ~3);
      if (f32) {
        for (int v = va; v < V4; v += 4) {
          const float x0 = __uint_as_float(wr[v]), x1 = __uint_as_float(wr[v + 1]), x2 = __uint_as_float(wr[v + 2]), x3 = __uint_as_float(wr[v + 3]);
          nan |= (x0 != x0) | (x1 != x1) | (x2 != x2) | (x3 != x3);
          m0 = x0 > m0 ? x0 : m0; m1 = x1 > m1 ? x1 : m1; m2 = x2 > m2 ? x2 : m2; m3 = x3 > m3 ? x3 : m3;
        }
        for (int v = V4; v < vb; ++v) { const float x = __uint_as_float(wr[v]); nan |= x != x; m0 = x > m0 ? x : m0; }
      } else {
        // 16-bit logits: unpack WORD by word (two symbols per 32-bit load, two instructions per symbol) into the emission
        // row as floats; the exp pass below reads them back.  (Unpacking symbol by symbol, twice, was half of all the
        // instructions BASELINE config 3 executed.)  Symbol v sits in halfword shift + v of the staged row.
        const int k0 = ((shift + va) >> 1) & ~1, nw = (shift + vb + 1) >> 1;
        const bool bf = p.dtype == E2E_BF16;
        for (int k = k0; k < nw; k += 2) {
          const uint32_t w0 = wr[k], w1 = k + 1 < nw ? wr[k + 1] : 0u;
          float a0, a1, a2, a3;
          if (bf) { a0 = __uint_as_float(w0 << 16); a1 = __uint_as_float(w0 & 0xffff0000u); a2 = __uint_as_float(w1 << 16); a3 = __uint_as_float(w1 & 0xffff0000u); }
          else {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
            a0 = f0.x; a1 = f0.y; a2 = f1.x; a3 = f1.y;
          }
          const int v0 = 2 * k - shift;
          if (v0 >= va && v0 < vb) { Ef[v0] = a0; nan |= a0 != a0; m0 = a0 > m0 ? a0 : m0; }
          if (v0 + 1 >= va && v0 + 1 < vb) { Ef[v0 + 1] = a1; nan |= a1 != a1; m1 = a1 > m1 ? a1 : m1; }
          if (v0 + 2 >= va && v0 + 2 < vb) { Ef[v0 + 2] = a2; nan |= a2 != a2; m2 = a2 > m2 ? a2 : m2; }
          if (v0 + 3 >= va && v0 + 3 < vb) { Ef[v0 + 3] = a3; nan |= a3 != a3; m3 = a3 > m3 ? a3 : m3; }
        }
      }
      auto elem = [&](int v) -> float { return f32 ? __uint_as_float(wr[v]) : Ef[v]; };
      float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      for (int o = p.cf; o < 32; o <<= 1) {      // the frame's other lanes (none when a lane owns the whole row)
        m = fmaxf(m, __shfl_xor_sync(grp, m, o));
        nan |= __shfl_xor_sync(grp, (int)nan, o) != 0;
      }
      // exp(x - max) once per symbol: kept in the emission row, normalised below
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int v = va; v < V4; v += 4) {
        const float e0 = expf(elem(v) - m), e1 = expf(elem(v + 1) - m), e2 = expf(elem(v + 2) - m), e3 = expf(elem(v + 3) - m);
        Ef[v] = e0; Ef[v + 1] = e1; Ef[v + 2] = e2; Ef[v + 3] = e3;
        s0 += e0; s1 += e1; s2 += e2; s3 += e3;
      }
      for (int v = V4; v < vb; ++v) { const float ev = expf(elem(v) - m); Ef[v] = ev; s0 += ev; }
      float s = (s0 + s1) + (s2 + s3);
      if (lpf == 2) {         // partial sums of the frame's lanes, added in lane order: the same bits on every lane
        const float q0 = __shfl_sync(grp, s, f), q1 = __shfl_sync(grp, s, f + p.cf);
        s = q0 + q1;
      } else if (lpf == 4) {
        const float q0 = __shfl_sync(grp, s, f), q1 = __shfl_sync(grp, s, f + p.cf), q2 = __shfl_sync(grp, s, f + 2 * p.cf), q3 = __shfl_sync(grp, s, f + 3 * p.cf);
        s = (q0 + q1) + (q2 + q3);
      }
      float inv = 1.f / s;
      if (nan) inv = NAN;
      for (int v = va; v < vb; ++v) Ef[v] *= inv;
      if (sub != 0) return lse_part;       // the frame's first lane reports the row normaliser
      if (!p.from_logits) {
        const double mls = (double)m + (double)logf(s);
        cx.rs[f] = nan ? NAN : (float)exp(mls);
        lse_part = nan ? (double)NAN : mls;
      }
    }
  } else {
    // gather mode: columns [0] blank, [1 + h*32 + lane'] label lane'*H + h
    if (F64) {
      const double2 st = *reinterpret_cast<const double2*>(cx.stat + f * 16);
      const double* xr = reinterpret_cast<const double*>(rslot);
      for (int k = 0; k <= Li; ++k) {
        const int col = k == 0 ? 0 : (1 + ((k - 1) % H) * 32 + (k - 1) / H);
        Erow[col] = (ET)exp((xr[k] - st.x) - st.y);
      }
      if (!p.from_logits) lse_part = st.x + st.y;
    } else {
      const float2 st = *reinterpret_cast<const float2*>(cx.stat + f * 16);
      const uint32_t* wr = reinterpret_cast<const uint32_t*>(rslot);
      const char* rowp = lbase + (long long)t * p.st * esz;
      for (int k = 0; k <= Li; ++k) {
        const int sym = k == 0 ? p.blank : cx.lab[k - 1];
        float x;
        if (p.dtype == E2E_F32) x = __uint_as_float(wr[k]);
        else x = sw_half_to_float(wr[k], p.dtype, (reinterpret_cast<uintptr_t>(rowp + (size_t)sym * 2) & 2) != 0);
        const int col = k == 0 ? 0 : (1 + ((k - 1) % H) * 32 + (k - 1) / H);
        float ev;
        if (p.from_logits) ev = expf((x - st.x) - st.y);
        else ev = expf((float)((double)x - ((double)st.x + (double)st.y)));
        reinterpret_cast<float*>(Erow)[col] = ev;
      }
      if (!p.from_logits) lse_part = (double)st.x + (double)st.y;
    }
  }
  return lse_part;
}

// ---- block-exponent snapshot: where the lane's scale should move ----------------------------------
template <int K, bool BWD, typename ET>
__device__ __forceinline__ void snapshot(SweepState<K, ET>& s, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  int mhi = 0;
#pragma unroll
  for (int j = 0; j < K; j++) mhi = max(mhi, __double2hiint(s.x[j]));
  int ec = mhi == 0 ? kNegExp : s.e + ((mhi >> 20) - 1023);
  // an all-zero lane takes the exponent of the nearest live lane on the side its mass will come from
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = BWD ? __shfl_down_sync(FULL, ec, d) : __shfl_up_sync(FULL, ec, d);
    const bool has = BWD ? (lane + d < 32) : (lane >= d);
    if (has && ec == kNegExp) ec = t;
  }
  int bec = BWD ? __shfl_down_sync(FULL, ec, 1) : __shfl_up_sync(FULL, ec, 1);
  const bool edge = BWD ? (lane == 31) : (lane == 0);
  if (edge) bec = kNegExp;
  // Own exponent, but at most kExpSlack below the exponent of the lane the mass comes from, so that incoming
  // cells are scaled by at most 2^kExpSlack.  (A plain max(ec, bec) flushed a lane whose own mass is more than
  // 2^1022 below its neighbour's -- the only feasible path of a tight alignment under very peaky emissions.)
  constexpr int kExpSlack = 900;
  const int en = max(ec, bec - kExpSlack);
  const int ben = BWD ? __shfl_down_sync(FULL, en, 1) : __shfl_up_sync(FULL, en, 1);
  s.en_next = en;
  s.f_next = pow2i(s.e - en);
  s.fb_next = edge ? 0.0 : pow2i(ben - en);
}

// ---- one sweep --------------------------------------------------------------------------------------
template <int K, bool BWD, typename ET>
__device__ void run_sweep(const SweepParams& p, unsigned char* smem, int b, int Ti, int Li, int lane, int* zero_flag) {
  using C = SweepCfg<K>;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int H = C::H, WORDS = C::WORDS, ROWW = C::ROWW, PF = C::PF;
  constexpr bool F64 = sizeof(ET) == 8;
  const int S = 2 * Li + 1;
  const int s0 = lane * K;

  unsigned char* wbase = smem + p.off_warp + (BWD ? p.warp_bytes : 0);
  ChunkCtx<ET> cx;
  cx.E = reinterpret_cast<ET*>(wbase + p.w_E);
  cx.raw = wbase + p.w_raw;
  cx.stat = wbase + p.w_stat;
  cx.rs = reinterpret_cast<float*>(wbase + p.w_rs);
  cx.lab = reinterpret_cast<int*>(smem + p.off_lab);
  uint32_t* acc = reinterpret_cast<uint32_t*>(wbase + p.w_acc);
  uint32_t* stage = reinterpret_cast<uint32_t*>(wbase + p.w_stage) + (size_t)lane * WORDS;   // + slot*ROWW

  const int esz = F64 ? 8 : (p.dtype == E2E_F32 ? 4 : 2);
  const char* lbase = reinterpret_cast<const char*>(p.logits) + (long long)b * p.sb * esz;

  // per-lane lattice constants
  const int zero_col = p.dense ? p.V : 0;   // gather mode: invalid label cells read their own (never written, zero) column
  const int bcol = p.dense ? p.blank : 0;
  int ecol[H];
  unsigned skipm = 0;
#pragma unroll
  for (int h = 0; h < H; h++) {
    const int li = lane * H + h;
    const bool lv = li < Li;
    const int lab = cx.lab[li];      // padded with blank past L_i
    ecol[h] = p.dense ? (lv ? lab : zero_col) : (1 + h * 32 + lane);
    bool sk = false;
    if (lv) {
      if (!BWD) sk = li >= 1 && lab != p.blank && lab != cx.lab[li - 1];
      else sk = li + 1 < Li && lab != p.blank && cx.lab[li + 1] != lab;
    }
    skipm |= sk ? (1u << h) : 0u;
  }

  SweepState<K, ET> s;
#pragma unroll
  for (int j = 0; j < K; j++) {
    // forward: a virtual frame before the first one with all mass on cell 0;
    // backward: beta of the last frame (ctc_loss.cpp:74-77): 1 on the final blank and the final label
    const bool on = BWD ? (s0 + j == S - 1 || s0 + j == S - 2) : (s0 + j == 0);
    s.x[j] = on ? 1.0 : 0.0;
  }
  s.e = 0;
  s.fb = (BWD ? lane == 31 : lane == 0) ? 0.0 : 1.0;
  s.en_next = 0; s.f_next = 1.0; s.fb_next = s.fb;

  const int tm = Ti / 2;
  const int nstore = BWD ? (Ti - tm) : tm;   // frames this sweep stores; the rest it combines
  uint32_t* const stash_u = p.stash + (size_t)b * p.T * ROWW + (size_t)lane * WORDS;   // + t*ROWW
  auto frame_t = [&](int i) { return BWD ? (Ti - 1 - i) : i; };
  // The reference's window (ctc_loss.cpp:44-45): at frame t only cells [max(0, S - 2(T - t)), min(2t + 2, S)) can carry
  // alpha * beta != 0 -- below it beta is exactly 0, above it alpha is.  A lane whose K cells all lie outside the window
  // neither stores its part of the row nor fetches the other sweep's part (the combine step takes zeros): for long
  // utterances about half of the stash traffic.
  // (Wide variants only: with few cells per lane the stash is small and the test costs more than it saves -- c3, K = 4:
  // 173 -> 188 us with it, c5, K = 40: 13.7 -> 11.9 ms.)
  constexpr bool kBand = K >= 16;
  auto lane_live = [&](int t) { return !kBand || (s0 < min(2 * t + 2, S) && s0 + K > S - 2 * (Ti - t)); };
  auto prefetch = [&](int i2) {   // the other sweep's stored row of iteration i2 -> staging slot i2 % PF
    if (i2 < Ti && lane_live(frame_t(i2))) {
      uint32_t* dst = stage + (size_t)(i2 % PF) * ROWW;
      const uint32_t* src = stash_u + (size_t)frame_t(i2) * ROWW;
#pragma unroll
      for (int u = 0; u < WORDS / 4; u++) sw_cp_async_cg16(dst + 4 * u, src + 4 * u);
    }
  };

  const int cf = p.cf;
  double lse_part = 0.0;
  bool have_z = false;
  double invz = 0.0;
  int Ez = 0;

  issue_chunk<BWD, ET>(p, cx, lbase, b, Ti, Li, 0, lane);
  sw_cp_async_commit();

  int fch = cf;   // frame index inside the current chunk (cf: a new chunk starts)
  int c = -1;
  for (int i = 0; i < Ti; ++i) {
    if (fch == cf) {
      ++c; fch = 0;
      // the chunk's copies were committed at least one chunk (>= PF frames) ago, except for the very first one
      if (i == 0) sw_cp_async_wait<0>(); else sw_cp_async_wait<PF - 1>();
      __syncwarp();
      lse_part += convert_chunk<BWD, ET, H>(p, cx, lbase, Ti, Li, c, lane);
      __syncwarp();
      issue_chunk<BWD, ET>(p, cx, lbase, b, Ti, Li, c + 1, lane);   // lands while this chunk is swept
    }
    const ET* Erow = cx.E + (size_t)fch * p.es;
    const bool snap = F64 || (i & 1) == 0;
    const bool apply = F64 || (i & 1) == 1;

    if (!BWD) {
      // emissions of this frame, with the pending re-centring folded in
      double mulb = (double)Erow[bcol];
      double mull[H];
#pragma unroll
      for (int h = 0; h < H; h++) mull[h] = (double)Erow[ecol[h]];
      if (snap) snapshot<K, BWD, ET>(s, lane);
      if (apply) {
        mulb *= s.f_next;
#pragma unroll
        for (int h = 0; h < H; h++) mull[h] *= s.f_next;
      }
      const double bxs = __shfl_up_sync(FULL, s.x[K - 1], 1) * s.fb;
      // cell j gathers j, j-1 and (label cells, when allowed) j-2 of the previous frame; in place, top down
#pragma unroll
      for (int j = K - 1; j >= 2; j--) {
        double so = s.x[j] + s.x[j - 1];
        if ((j & 1) && (skipm & (1u << (j >> 1)))) so += s.x[j - 2];
        s.x[j] = so * ((j & 1) ? mull[j >> 1] : mulb);
      }
      {
        double so = s.x[1] + s.x[0];
        if (skipm & 1u) so += bxs;
        s.x[1] = so * mull[0];
        s.x[0] = (s.x[0] + bxs) * mulb;
      }
      if (apply) { s.e = s.en_next; s.fb = s.fb_next; }
    } else {
      if (i > 0) {
        // state is gamma = beta * emission of frame t+1; beta_t(j) gathers j, j+1 and (label cells) j+2
        const double bys0 = __shfl_down_sync(FULL, s.x[0], 1) * s.fb;
        const double bys1 = __shfl_down_sync(FULL, s.x[1], 1) * s.fb;
#pragma unroll
        for (int j = 0; j < K - 2; j++) {
          double so = s.x[j] + s.x[j + 1];
          if ((j & 1) && (skipm & (1u << (j >> 1)))) so += s.x[j + 2];
          s.x[j] = so;
        }
        s.x[K - 2] = s.x[K - 2] + s.x[K - 1];
        {
          double so = s.x[K - 1] + bys0;
          if (skipm & (1u << (H - 1))) so += bys1;
          s.x[K - 1] = so;
        }
      }
      if (snap) snapshot<K, BWD, ET>(s, lane);
    }

    // ---- store (first half) or combine with the other sweep's stored row (second half) ----
    const int t = frame_t(i);
    if (i < nstore) {
      uint32_t* dst = stash_u + (size_t)t * ROWW;
      uint32_t wds[WORDS];
#pragma unroll
      for (int j = 0; j < K; j++) wds[j] = (uint32_t)__double2hiint(s.x[j]);
      wds[K] = (uint32_t)s.e;
#pragma unroll
      for (int j = K + 1; j < WORDS; j++) wds[j] = 0u;
      if (lane_live(t)) {
#pragma unroll
        for (int u = 0; u < WORDS / 4; u++)
          reinterpret_cast<uint4*>(dst)[u] = make_uint4(wds[4 * u], wds[4 * u + 1], wds[4 * u + 2], wds[4 * u + 3]);
      }
      if (i == nstore - 1) {
        // meet: publish my stored rows, wait for the other sweep's (both warps pass here exactly once)
        __threadfence();
        sw_named_barrier(1, 64);
        for (int u = 0; u < PF; u++) { prefetch(i + 1 + u); sw_cp_async_commit(); }
      } else {
        sw_cp_async_commit();   // one group per frame keeps the wait depth uniform
      }
    } else {
      if (nstore == 0 && i == 0) {   // a one-frame utterance: the forward sweep stores nothing
        sw_named_barrier(1, 64);
        for (int u = 0; u < PF; u++) { prefetch(i + u); sw_cp_async_commit(); }
      }
      sw_cp_async_wait<PF - 1>();
      const uint32_t* orow = stage + (size_t)(i % PF) * ROWW;
      uint32_t ow[WORDS];
      const bool olive = lane_live(t);   // the other sweep stored this lane's part of frame t
#pragma unroll
      for (int u = 0; u < WORDS / 4; u++) {
        const uint4 q = olive ? reinterpret_cast<const uint4*>(orow)[u] : make_uint4(0u, 0u, 0u, 0u);
        ow[4 * u] = q.x; ow[4 * u + 1] = q.y; ow[4 * u + 2] = q.z; ow[4 * u + 3] = q.w;
      }
      prefetch(i + PF);
      sw_cp_async_commit();
      const int El = s.e + (int)ow[K];
      double bs = 0.0, pr[H];
#pragma unroll
      for (int u = 0; u < H; u++) {
        bs = fma(s.x[2 * u], sw_unpack_hi32(ow[2 * u]), bs);          // blank cell 2u
        pr[u] = s.x[2 * u + 1] * sw_unpack_hi32(ow[2 * u + 1]);      // label cell 2u+1
      }
      if (!have_z) {
        // Z = sum_s alpha(t,s) * beta(t,s), the same for every frame t: taken once per sweep
        double lsum = bs;
#pragma unroll
        for (int u = 0; u < H; u++) lsum += pr[u];
        const int emax = warp_max_int(lsum > 0.0 ? El : 4 * kNegExp);
        const double tot = warp_sum(lsum > 0.0 ? lsum * pow2i(El - emax) : 0.0);
        // 1/Z (times 2^31 in dense mode: fixed-point posteriors) as a mantissa in [1,2) with its power of two
        // folded into the reference exponent: the per-lane scale 2^(El - Ez) * invz then stays FINITE whatever
        // stale exponent a lane without mass carries (0 * finite = 0; an infinite scale made 0 * inf = NaN, and
        // the float-to-integer conversion of NaN put a spurious posterior of 1 on such a lane's symbols --
        // tight alignments under very peaky emissions, found by scratch/gpu_fuzz.py).
        // tot == 0 or NaN (no path survives): garbage posteriors; the forward sweep flags the utterance and the
        // block is overwritten with NaN.
        const double rz = (p.dense ? 2147483648.0 : 1.0) / tot;
        const int kz = ((__double2hiint(rz) >> 20) & 0x7ff) - 1023;
        invz = __hiloint2double((__double2hiint(rz) & 0x800fffff) | 0x3ff00000, __double2loint(rz));
        Ez = emax - kz;
        have_z = true;
      }
      if (p.dense) {
        // posteriors summed per symbol with integer shared-memory atomics (fixed point 2^-31: the sum per
        // symbol is <= 1; integer adds commute, so the gradient is bitwise reproducible)
        uint32_t* arow = acc + (size_t)(i & 1) * p.vpad;
        const double ccl = pow2i(El - Ez) * invz;
#pragma unroll
        for (int u = 0; u < H; u++) atomicAdd(arow + ecol[u], __double2uint_rn(pr[u] * ccl));   // cells past L_i add 0 to the spare slot V
        const uint32_t qb = __reduce_add_sync(FULL, __double2uint_rn(bs * ccl));
        __syncwarp();
        // gradient row: scale * (softmax - posterior) (ctc_loss.cpp:116-117 + log_softmax backward);
        // log-prob input: exp(lp) - posterior (the engine contract)
        const long long gbase = (long long)b * p.gsb + (long long)t * p.gst;
        if (F64) {
          const double rs = p.from_logits ? 1.0 : reinterpret_cast<const double*>(cx.stat)[2 * fch];
          double* g = reinterpret_cast<double*>(p.grads) + gbase;
          for (int v = lane; v < p.V; v += 32) {
            const uint32_t a = arow[v] + (v == p.blank ? qb : 0u);
            arow[v] = 0u;
            g[v] = p.scale * ((double)Erow[v] * rs - (double)a * (1.0 / 2147483648.0));
          }
        } else {
          const float rs = p.from_logits ? 1.f : cx.rs[fch], sc = (float)p.scale;
          for (int v = lane; v < p.V; v += 32) {
            const uint32_t a = arow[v] + (v == p.blank ? qb : 0u);
            arow[v] = 0u;
            const float gv = sc * ((float)Erow[v] * rs - (float)a * (1.f / 2147483648.f));
            if (p.dtype == E2E_F32) reinterpret_cast<float*>(p.grads)[gbase + v] = gv;
            else if (p.dtype == E2E_BF16) reinterpret_cast<__nv_bfloat16*>(p.grads)[gbase + v] = __float2bfloat16_rn(gv);
            else reinterpret_cast<__half*>(p.grads)[gbase + v] = __float2half_rn(gv);
          }
        }
      } else {
        // compact posterior row for the gradient kernel: [label 0 .. label cells/2-1 | blank total]
        float* prow = p.post + ((size_t)b * p.T + t) * (size_t)p.post_stride;
        const double cc = pow2i(El - Ez) * invz;
        float lp[H];
#pragma unroll
        for (int u = 0; u < H; u++) lp[u] = (float)(pr[u] * cc);
        float* dst = prow + (size_t)lane * H;
        if (H % 4 == 0) {
#pragma unroll
          for (int u = 0; u < H / 4; u++)
            reinterpret_cast<float4*>(dst)[u] = make_float4(lp[4 * u], lp[4 * u + 1], lp[4 * u + 2], lp[4 * u + 3]);
        } else if (H % 2 == 0) {
#pragma unroll
          for (int u = 0; u < H / 2; u++) reinterpret_cast<float2*>(dst)[u] = make_float2(lp[2 * u], lp[2 * u + 1]);
        } else {
#pragma unroll
          for (int u = 0; u < H; u++) dst[u] = lp[u];
        }
        const double bsum = warp_sum(bs * cc);
        if (lane == 0) prow[p.cells / 2] = (float)bsum;
      }
    }

    if (BWD) {
      // gamma_t = beta_t * emission of frame t (pending re-centring folded in)
      double mulb = (double)Erow[bcol];
      double mull[H];
#pragma unroll
      for (int h = 0; h < H; h++) mull[h] = (double)Erow[ecol[h]];
      if (apply) {
        mulb *= s.f_next;
#pragma unroll
        for (int h = 0; h < H; h++) mull[h] *= s.f_next;
      }
#pragma unroll
      for (int j = 0; j < K; j++) s.x[j] *= (j & 1) ? mull[j >> 1] : mulb;
      if (apply) { s.e = s.en_next; s.fb = s.fb_next; }
    }
    ++fch;
  }
  sw_cp_async_wait<0>();

  // loss = -log(alpha[S-1][T-1] + alpha[S-2][T-1]) (ctc_loss.cpp:63-70), from the live fp64 forward
  // state.  Emissions were normalised per row, so for log-prob input the row normalisers (all ~0
  // for true log-probabilities) are added back.
  if (!BWD) {
    double tail = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) if (s0 + j == S - 1 || s0 + j == S - 2) tail += s.x[j];
    const int emax = warp_max_int(tail > 0.0 ? s.e : 4 * kNegExp);
    const double z = warp_sum(tail > 0.0 ? tail * pow2i(s.e - emax) : (tail != tail ? tail : 0.0));
    const double lse_tot = warp_sum(lse_part);
    if (lane == 0) {
      double loss = INFINITY;
      if (z > 0.0) {
        loss = -(log(z) + (double)emax * 0.69314718055994530942);
        if (!p.from_logits) loss -= lse_tot;
      } else {   // no path survives (exact-zero emissions): +inf; NaN input: NaN.  NaN gradient block either way
        if (z != z) loss = NAN;
        p.flags[b] = kFlagInfeasible;
        *zero_flag = 1;
      }
      store_from_double(p.losses, p.dtype, b, loss);
    }
  }
}

// ---- kernel -----------------------------------------------------------------------------------
template <int K, bool F64>
__global__ void __launch_bounds__(64, K <= 8 ? E2E_SWEEP_MINBLK : (K >= 32 ? E2E_SWEEP_WIDE_MINBLK : 1))
ctc_sweep_kernel(const SweepParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using ET = typename std::conditional<F64, double, float>::type;
  constexpr int H = K / 2;
  int* lab = reinterpret_cast<int*>(smem_raw + p.off_lab);
  __shared__ int misc[4];

  const int b = p.order ? p.order[blockIdx.x] : (int)blockIdx.x;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;

  const long long Ti_ll = load_index(p.in_len, p.len_is64, b);
  const long long Li_ll = load_index(p.tgt_len, p.len_is64, b);
  int bad = 0;
  if (Ti_ll < 1 || Ti_ll > p.T) bad |= kBadFrames;
  if (Li_ll < 0 || Li_ll > p.Lmax) bad |= kBadTargetLen;
  const int Ti = (int)Ti_ll, Li = bad ? 0 : (int)Li_ll;
  if (tid < 4) misc[tid] = 0;
  __syncthreads();
  int rep = 0, badlab = 0;
  for (int i = tid; i < Li; i += blockDim.x) {
    const long long v = load_index(p.targets, p.tgt_is64, (long long)b * p.ts_b + i);
    if (v < 0 || v >= p.V) badlab = kBadLabel;
    lab[i] = (int)v;
  }
  for (int i = Li + tid; i < 32 * H + 1; i += blockDim.x) lab[i] = p.blank;   // cells past the lattice carry zero mass
  __syncthreads();
  for (int i = tid + 1; i < Li; i += blockDim.x) rep += (lab[i] == lab[i - 1]);
  if (rep) atomicAdd(&misc[1], rep);
  if (badlab) atomicOr(&misc[0], badlab);
  __syncthreads();
  bad |= misc[0];
  rep = misc[1];
  const long long gfill_base = (long long)b * p.gsb;
  if (bad || Ti < Li + rep) {
    // out-of-range lengths / labels (undefined behaviour in the reference): NaN loss + status bits;
    // no alignment exists (T < L + repeats): loss = +inf.  Either way the gradient block is all NaN
    // (-inf - (-inf) in the reference, ctc_loss.cpp:116-117), padding rows included.
    if (tid == 0) {
      if (bad) atomicOr(p.status, bad);
      p.flags[b] = bad ? kFlagInvalid : kFlagInfeasible;
      store_from_double(p.losses, p.dtype, b, bad ? (double)NAN : (double)INFINITY);
    }
    if (p.dense && p.grads != nullptr) {
      for (int r = w; r < p.T; r += 2)
        for (int v = lane; v < p.V; v += 32) store_from_double(p.grads, p.dtype, gfill_base + (long long)r * p.gst + v, (double)NAN);
    }
    return;
  }
  if (tid == 0) p.flags[b] = 0;
  {  // emission buffers: every column the conversion never writes must read as zero; accumulators start at zero
    uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw + p.off_warp);
    const int n = (2 * p.warp_bytes) >> 2;
    for (int k = tid; k < n; k += blockDim.x) z[k] = 0u;
  }
  __syncthreads();

  if (w == 0) run_sweep<K, false, ET>(p, smem_raw, b, Ti, Li, lane, &misc[2]);
  else run_sweep<K, true, ET>(p, smem_raw, b, Ti, Li, lane, &misc[2]);

  if (p.dense && p.grads != nullptr) {
    __syncthreads();
    if (misc[2]) {
      // Z == 0 although an alignment exists (exact-zero emissions): the reference yields +inf / NaN
      for (int r = w; r < p.T; r += 2)
        for (int v = lane; v < p.V; v += 32) store_from_double(p.grads, p.dtype, gfill_base + (long long)r * p.gst + v, (double)NAN);
    } else {
      // padding frames t >= T_i: exp(lp) for log-prob input (the engine contract, ctc_loss.cpp:105-117),
      // 0 for fused-logits input (what the reference's log_softmax backward leaves there)
      for (int r = Ti + w; r < p.T; r += 2) {
        const long long xo = (long long)b * p.sb + (long long)r * p.st;
        const long long go = gfill_base + (long long)r * p.gst;
        for (int v = lane; v < p.V; v += 32) {
          double g = 0.0;
          if (!p.from_logits)
            g = p.dtype == E2E_F64 ? exp(load_as_double(p.logits, p.dtype, xo + v)) : (double)expf(load_as_float(p.logits, p.dtype, xo + v));
          store_from_double(p.grads, p.dtype, go + v, p.scale * g);
        }
      }
    }
  }
}

template <int K, bool F64>
int launch_sweep_k(const SweepParams& sp, size_t smem, cudaStream_t s) {
  // the dynamic shared-memory opt-in is PER DEVICE: cached per device ordinal (it only ever grows)
  static int attr_smem[64];
  int dev = 0;
  E2E_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || (int)smem > attr_smem[dev]) {
    E2E_CUDA_TRY(cudaFuncSetAttribute(ctc_sweep_kernel<K, F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_smem[dev] = (int)smem;
  }
  KernelTimer timer(kKernelLattice, s);
  ctc_sweep_kernel<K, F64><<<(unsigned)sp.B, 64, smem, s>>>(sp);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace
}  // namespace e2e
