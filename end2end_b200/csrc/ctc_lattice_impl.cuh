// K2 -- alpha/beta recursion over the blank-extended label lattice (S = 2L+1 cells), and in
// "dense" mode the whole loss path (row log-softmax statistics, lattice, gradient write) in ONE kernel.
//
// Replaces CTCLossEngine::compute_2d (src/losses/ctc_loss.cpp:15-118): extended targets (:25-31),
// alpha (:33-61), loss (:63-70), beta (:72-100), alpha+beta / gradient (:102-117); in dense mode
// also F.log_softmax (pytorch_end2end/modules/ctc_loss.py:40) and the exp(logits) term (:117).
//
// Design (B200-first):
//  * One 2-CTA thread-block cluster per utterance.  CTA rank 0 runs the forward (alpha) sweep
//    t = 0..T-1, rank 1 the backward (beta) sweep t = T-1..0, concurrently on two SMs; every frame
//    is "stored" by the sweep that reaches it first and "combined" by the other one, so the
//    dependent chain is T frames, not 2T, and every posterior is produced exactly once.
//  * Warp specialisation inside a CTA, three roles decoupled by mbarrier rings of frame chunks:
//      producers : stage per-frame emissions p(t, symbol) into a shared-memory ring (cp.async
//                  gathers kNumChunks chunks ahead; dense mode also computes the row max / log-sum-exp
//                  from the staged row, i.e. the fused log_softmax);
//      lattice   : the recurrence and nothing else.  One warp covers up to 32*K cells (K <= 40
//                  cells per lane, so S <= 1280 needs no block barrier at all; two warps with a
//                  named barrier beyond that).  It reads emissions with LDS, exchanges the s-1/s-2
//                  neighbours with warp shuffles and drops its state (top 32 bits of each fp64
//                  cell + the lane's block exponent) into a shared-memory ring.  It never touches
//                  global memory and never waits for the other sweep.
//      combiners : drain that ring.  First half: copy the rows to the global stash (L2).  After
//                  the two sweeps' combiners have met (one cluster-scope mbarrier hand-off), second
//                  half: cp.async-prefetch the other sweep's stashed row, multiply, normalise by
//                  Z = sum_s alpha*beta, and either (gather mode) write compact per-label posteriors
//                  for the gradient kernel, or (dense mode) sum them per symbol with integer
//                  shared-memory atomics and write the gradient row softmax - posterior themselves.
//  * Arithmetic is LINEAR-domain fp64 with a per-lane block exponent (value = x * 2^e): a cell
//    update is 1 DADD + 1 DMUL (+1 DFMA for label cells, the repeat-label skip is a 0/1 multiplier),
//    no MUFU on the chain, error ~1e-16 per step (an fp32 log-space recursion fails the 1e-5 parity
//    budget).  Renormalisation is lane-local integer work lagged by one frame.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace e2e {
namespace {

struct LatticeParams {
  const void* logits; int dtype; long long sb, st;
  void* grads; long long gsb, gst; double scale;   // dense (fused) mode: gradient output
  const void* stats;                               // gather mode: row {max, logsumexp} from K1
  const void* targets; int tgt_is64; long long ts_b;
  const void* in_len; const void* tgt_len; int len_is64;
  int B, T, V, Lmax, blank, from_logits;
  void* losses;
  int* status; int* flags;
  uint32_t* stash;   // [B*T][lanes][words]  first-half lattice state
  float* post;       // gather mode: [B*T][post_stride] compact posteriors (labels..., blank at cells/2)
  int np, nc, pfd;   // producer warps, combiner warps, stashed rows each combiner warp keeps in flight
  long long* trace;  // debugging: per-chunk clock64 stamps of utterances 0/1 (E2E_CTC_TRACE=1), else NULL
  LatticeSmem sm;
  int chunk_log2, lstride, dense, rowlen_max, post_stride, vpad;
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
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_ca(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES)
               : "memory");
}
__device__ __forceinline__ void cp_async_cg16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kTraceChunks = 128, kTraceEvents = 4;
__device__ __forceinline__ void trace_stamp(const LatticeParams& p, int b, bool bwd, int role, int c, int ev) {
  if (p.trace != nullptr && b < 2 && c < kTraceChunks)
    p.trace[((((size_t)b * 2 + (bwd ? 1 : 0)) * 3 + role) * kTraceChunks + c) * kTraceEvents + ev] = clock64();
}
__device__ __forceinline__ double unpack_hi32(uint32_t h) { return __hiloint2double((int)h, 0); }

// ---- shared-memory view -------------------------------------------------------------------------
struct SmemView {
  int* lab; int* misc; double* lsesum;
  double* E; unsigned char* raw; unsigned char* rstat; uint32_t* val; uint32_t* stage; uint32_t* acc;
  Boundary* bnd; double* redd; int* redi;
  uint64_t* full; uint64_t* latdone; uint64_t* empty; uint64_t* meet;
};
__device__ __forceinline__ SmemView carve(unsigned char* base, const LatticeSmem& L) {
  SmemView v;
  v.lab = reinterpret_cast<int*>(base + L.lab);
  v.misc = reinterpret_cast<int*>(base + L.misc);
  v.lsesum = reinterpret_cast<double*>(base + L.misc + 64);
  v.E = reinterpret_cast<double*>(base + L.E);
  v.raw = base + L.raw;
  v.rstat = base + L.rstat;
  v.val = reinterpret_cast<uint32_t*>(base + L.val);
  v.stage = reinterpret_cast<uint32_t*>(base + L.stage);
  v.acc = reinterpret_cast<uint32_t*>(base + L.acc);
  v.bnd = reinterpret_cast<Boundary*>(base + L.bnd);
  v.redd = reinterpret_cast<double*>(base + L.red);
  v.redi = reinterpret_cast<int*>(base + L.red + 128);
  v.full = reinterpret_cast<uint64_t*>(base + L.bars);
  v.latdone = v.full + kNumChunks;
  v.empty = v.latdone + kNumChunks;
  v.meet = v.empty + kNumChunks;
  return v;
}

// ---- emissions --------------------------------------------------------------------------------
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

__device__ __forceinline__ float raw_to_float(uint32_t raw, int dtype, bool upper_half) {
  if (dtype == E2E_F32) return __uint_as_float(raw);
  const uint32_t half = upper_half ? (raw >> 16) : (raw & 0xffffu);
  return dtype == E2E_BF16 ? __uint_as_float(half << 16) : __half2float(__ushort_as_half((unsigned short)half));
}

// Emission-ring columns.  gather mode: [0] blank, [1] zero, [2 + h*lanes + lane] label lane*H+h
// (h-major so that the 32 lanes of a lattice warp read consecutive doubles); dense mode: [v] symbol v,
// [V] zero, [V+1] exp(row max + logsumexp) (the factor that turns the emission back into exp(x)).
//
// Producer warps: stage the emissions of `chunk` frames at a time into the E ring.  Every item's raw
// logit is fetched with cp.async kNumChunks chunks ahead of its conversion, so the L2/HBM latency never
// reaches the lattice warps.  A single warp issues at most ~0.5 instructions per cycle on this SM, so the
// conversion is spread over np warps and kept short:
//   dense mode : one warp per frame (frames round-robin over the producer warps), <= 4 symbols per lane
//                held in registers: warp-shuffle max / sum-exp, then the emissions -- the fused row
//                log-softmax (pytorch_end2end/modules/ctc_loss.py:40);
//   gather mode: a thread owns emission columns k = ptid, ptid + npt, ... of every frame of the chunk.
template <bool BWD, bool F64, int H, int LANES>
__device__ void run_producer(const LatticeParams& p, const SmemView& sv, int b, int Ti, int Li, int pw, int plane) {
  constexpr int RAWSZ = F64 ? 8 : 4;
  constexpr unsigned FULL = 0xffffffffu;
  const int ptid = pw * 32 + plane, npt = p.np * 32;
  const int cs = p.chunk_log2, CF = 1 << cs;
  const int rowlen = p.dense ? p.V : Li + 1;
  const int nchunks = (Ti + CF - 1) >> cs;
  const int esz = p.dtype == E2E_F32 ? 4 : (p.dtype == E2E_F64 ? 8 : 2);
  const char* lbase = reinterpret_cast<const char*>(p.logits) + (long long)b * p.sb * esz;
  const long long st_bytes = p.st * esz;
  const int ring_mask = (kNumChunks << cs) - 1;
  const bool want_lse = !BWD && !p.from_logits;
  const int rowraw = p.rowlen_max * RAWSZ;                        // bytes of raw staging per frame

  auto frame_t = [&](int i) { return BWD ? (Ti - 1 - i) : i; };
  auto issue = [&](int c) {
    if (c < nchunks) {
      const int nf = min(CF, Ti - (c << cs));
      unsigned char* chunk_raw = sv.raw + (size_t)(c & (kNumChunks - 1)) * CF * rowraw;
      if (p.dense) {       // a frame's row is contiguous: lanes over symbols, frames over warps
        for (int f = pw; f < nf; f += p.np) {
          const char* rowp = lbase + (long long)frame_t((c << cs) + f) * st_bytes;
          unsigned char* rslot = chunk_raw + (size_t)f * rowraw;
          for (int v = plane; v < p.V; v += 32) {
            if (F64) cp_async_ca<8>(rslot + v * 8, rowp + (size_t)v * 8);
            else     cp_async_ca<4>(rslot + v * 4, reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(rowp + (size_t)v * esz) & ~(uintptr_t)3));
          }
        }
      } else {
        for (int k = ptid; k < rowlen; k += npt) {
          const size_t soff = (size_t)(k == 0 ? p.blank : sv.lab[k - 1]) * esz;
          for (int f = 0; f < nf; ++f) {
            const char* src = lbase + (long long)frame_t((c << cs) + f) * st_bytes + soff;
            unsigned char* slot = chunk_raw + (size_t)f * rowraw + (size_t)k * RAWSZ;
            if (F64) cp_async_ca<8>(slot, src);
            else     cp_async_ca<4>(slot, reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3));
          }
        }
        if (ptid < nf) {   // row statistics of the chunk's frames (written by K1)
          const long long row = (long long)b * p.T + frame_t((c << cs) + ptid);
          unsigned char* dst = sv.rstat + (size_t)(((c << cs) + ptid) & ring_mask) * 16;
          if (F64) cp_async_ca<16>(dst, reinterpret_cast<const char*>(p.stats) + row * 16);
          else     cp_async_ca<8>(dst, reinterpret_cast<const char*>(p.stats) + row * 8);
        }
      }
    }
    cp_async_commit();   // one group per chunk, also when empty, so the wait depth stays uniform
  };
  // 16-bit inputs: the staged 32-bit word holds the element in its low or high half
  auto cvt32 = [&](uint32_t raw, const char* elem_addr) -> float {
    if (p.dtype == E2E_F32) return __uint_as_float(raw);
    return raw_to_float(raw, p.dtype, (reinterpret_cast<uintptr_t>(elem_addr) & 2) != 0);
  };

  for (int c = 0; c < kNumChunks; ++c) issue(c);
  double lsesum = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const int slot_c = c & (kNumChunks - 1);
    cp_async_wait<kNumChunks - 1>();                             // this thread's copies of chunk c have landed
    if (ptid == 0) trace_stamp(p, b, BWD, 0, c, 0);
    mbar_wait(&sv.empty[slot_c], ((c / kNumChunks) & 1) ^ 1);   // lattice + combiners released the slot
    if (ptid == 0) trace_stamp(p, b, BWD, 0, c, 1);
    const int nf = min(CF, Ti - (c << cs));
    const unsigned char* chunk_raw = sv.raw + (size_t)slot_c * CF * rowraw;
    if (p.dense) {
      // every lane reads exactly the raw slots it fetched itself: no barrier needed
      for (int f = pw; f < nf; f += p.np) {
        const int fr = ((c << cs) + f) & ring_mask;
        const unsigned char* rslot = chunk_raw + (size_t)f * rowraw;
        const char* rowp = lbase + (long long)frame_t((c << cs) + f) * st_bytes;
        double* Erow = sv.E + (size_t)fr * p.lstride;
        if (F64) {
          double xv[4], m = -INFINITY; bool nan = false;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            xv[q] = -INFINITY;
            if (32 * q < p.V) {
              const int v = plane + 32 * q;
              if (v < p.V) xv[q] = *reinterpret_cast<const double*>(rslot + v * 8);
              nan |= xv[q] != xv[q]; m = xv[q] > m ? xv[q] : m;
            }
          }
          m = warp_max(m);
          double s = 0.0;
#pragma unroll
          for (int q = 0; q < 4; q++) if (32 * q < p.V) s += exp(xv[q] - m);      // exp(-inf) = 0 for the padding lanes
          s = warp_sum(s);
          double ls = log(s);
          if (__any_sync(FULL, nan)) { m = NAN; ls = NAN; }
#pragma unroll
          for (int q = 0; q < 4; q++) if (32 * q < p.V) { const int v = plane + 32 * q; if (v < p.V) Erow[v] = emission_f64(xv[q], m, ls); }
          if (plane == 0) { if (!p.from_logits) Erow[p.V + 1] = exp(m + ls); lsesum += m + ls; }
        } else {
          float xv[4], m = -INFINITY; bool nan = false;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            xv[q] = -INFINITY;
            if (32 * q < p.V) {
              const int v = plane + 32 * q;
              if (v < p.V) xv[q] = cvt32(*reinterpret_cast<const uint32_t*>(rslot + v * 4), rowp + (size_t)v * 2);
              nan |= xv[q] != xv[q]; m = xv[q] > m ? xv[q] : m;
            }
          }
          m = warp_max(m);
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 4; q++) if (32 * q < p.V) s += expf(xv[q] - m);
          s = warp_sum(s);
          float ls = logf(s);
          if (__any_sync(FULL, nan)) { m = NAN; ls = NAN; }
          if (p.from_logits) {   // (x - m) - ls is torch's fp32 log_softmax value; its exp, widened
#pragma unroll
            for (int q = 0; q < 4; q++) if (32 * q < p.V) { const int v = plane + 32 * q; if (v < p.V) Erow[v] = (double)expf((xv[q] - m) - ls); }
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++) if (32 * q < p.V) { const int v = plane + 32 * q; if (v < p.V) Erow[v] = emission_f32(xv[q], m, ls, 0); }
            if (plane == 0) { Erow[p.V + 1] = exp((double)m + (double)ls); lsesum += (double)m + (double)ls; }
          }
        }
        uint32_t* arow = sv.acc + (size_t)fr * p.vpad;
        for (int v = plane; v < p.vpad; v += 32) arow[v] = 0u;
      }
    } else {
      // threads read the row statistics fetched by OTHER producer threads' cp.async:
      // every thread has waited for its own copies, the barrier makes them visible group-wide
      named_barrier(3, npt);
      for (int k = ptid; k < rowlen; k += npt) {
        const int sym = k == 0 ? p.blank : sv.lab[k - 1];
        const int col = k == 0 ? 0 : (2 + ((k - 1) % H) * LANES + (k - 1) / H);
        for (int f = 0; f < nf; ++f) {
          const int fr = ((c << cs) + f) & ring_mask;
          const unsigned char* slot = chunk_raw + (size_t)f * rowraw + (size_t)k * RAWSZ;
          double em;
          if (F64) {
            const double2 st = *reinterpret_cast<const double2*>(sv.rstat + (size_t)fr * 16);
            em = emission_f64(*reinterpret_cast<const double*>(slot), st.x, st.y);
          } else {
            const float2 st = *reinterpret_cast<const float2*>(sv.rstat + (size_t)fr * 16);
            const float x = cvt32(*reinterpret_cast<const uint32_t*>(slot),
                                  lbase + (long long)frame_t((c << cs) + f) * st_bytes + (size_t)sym * 2);
            em = p.from_logits ? (double)expf((x - st.x) - st.y) : emission_f32(x, st.x, st.y, 0);
          }
          sv.E[(size_t)fr * p.lstride + col] = em;
        }
      }
      if (want_lse && ptid == 0) {
        // sum of the row normalisers (log-prob input only), fixed order => deterministic loss
        for (int f = 0; f < nf; ++f) {
          const unsigned char* st = sv.rstat + (size_t)(((c << cs) + f) & ring_mask) * 16;
          if (F64) lsesum += reinterpret_cast<const double*>(st)[0] + reinterpret_cast<const double*>(st)[1];
          else lsesum += (double)reinterpret_cast<const float*>(st)[0] + (double)reinterpret_cast<const float*>(st)[1];
        }
      }
    }
    if (want_lse && c == nchunks - 1 && plane == 0) sv.lsesum[pw] = lsesum;            // rides on the last hand-off
    if (ptid == 0) trace_stamp(p, b, BWD, 0, c, 2);
    mbar_arrive(&sv.full[slot_c]);
    if (!p.dense) named_barrier(3, npt);   // nobody still reads this chunk's statistics when they are refilled
    issue(c + kNumChunks);
  }
  cp_async_wait<0>();
}

// ---- sum of per-lane values v * 2^ex over the NW lattice warps of a sweep ------------------------
template <int NW>
__device__ __forceinline__ void sweep_sum_scaled(double v, int ex, const SmemView& sv, int w, int lane,
                                                 double* z, int* ez) {
  int emax = warp_max_int(v > 0.0 ? ex : 4 * kNegExp);
  if (NW > 1) {
    if (lane == 0) sv.redi[w] = emax;
    named_barrier(1, NW * 32);
    emax = sv.redi[0];
#pragma unroll
    for (int q = 1; q < NW; q++) emax = max(emax, sv.redi[q]);
  }
  double t = warp_sum(v > 0.0 ? v * pow2i(ex - emax) : 0.0);
  if (NW > 1) {
    if (lane == 0) sv.redd[w] = t;
    named_barrier(1, NW * 32);
    t = 0.0;
#pragma unroll
    for (int q = 0; q < NW; q++) t += sv.redd[q];
  }
  *z = t;
  *ez = emax;
}

// ---- the lattice warp(s) --------------------------------------------------------------------------
template <int K, int NW, bool BWD>
__device__ void run_lattice(const LatticeParams& p, const SmemView& sv, int b, int Ti, int Li, int w, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int H = K / 2;
  constexpr int WORDS = (K + 1 + 3) & ~3;
  constexpr int LANES = 32 * NW;
  const int S = 2 * Li + 1;
  const int lane_g = w * 32 + lane;
  const int s0 = lane_g * K;
  const int cs = p.chunk_log2, cmask = (1 << cs) - 1;
  const int ring_mask = (kNumChunks << cs) - 1;

  const int zero_col = p.dense ? p.V : 1;
  const int bcol = p.dense ? p.blank : 0;
  int ecol[H];
  double skipd[H];
#pragma unroll
  for (int h = 0; h < H; h++) {
    const int li = lane_g * H + h;  // label index of cell s0+2h+1
    const bool lv = li < Li;
    ecol[h] = p.dense ? (lv ? sv.lab[li] : zero_col) : (2 + h * LANES + lane_g);
    bool sk = false;
    if (lv) {
      const int lab = sv.lab[li];
      if (!BWD) sk = li >= 1 && lab != p.blank && lab != sv.lab[li - 1];
      else sk = li + 1 < Li && lab != p.blank && sv.lab[li + 1] != lab;
    }
    skipd[h] = sk ? 1.0 : 0.0;
  }
  int nbcol = zero_col;   // BWD: emission column of the next lane's first label cell
  if (BWD) {
    const int li_n = (lane_g + 1) * H;
    if (li_n < Li) nbcol = p.dense ? sv.lab[li_n] : (2 + lane_g + 1);
    if (!p.dense && lane_g + 1 >= LANES) nbcol = zero_col;
  }

  double x[K];
  int e = kNegExp, sh = 0;
  {
    bool any = false;
#pragma unroll
    for (int j = 0; j < K; j++) {
      // forward: a virtual frame before the first one with all mass on cell 0;
      // backward: beta of the last frame (ctc_loss.cpp:74-77): 1 on the final blank and the final label
      const bool on = BWD ? (s0 + j == S - 1 || s0 + j == S - 2) : (s0 + j == 0);
      x[j] = on ? 1.0 : 0.0;
      any |= on;
    }
    if (any) e = 0;
  }
  if (NW > 1) {   // publish the initial boundary cells
    Boundary* bw = sv.bnd + (BWD ? 1 : 0) * 8;
    if (!BWD) { if (lane == 31) { bw[w].x0 = x[K - 1]; bw[w].e = e; } }
    else { if (lane == 0) { bw[w].x0 = x[0]; bw[w].x1 = x[1]; bw[w].e = e; } }
  }

  for (int i = 0; i < Ti; ++i) {
    const int c = i >> cs;
    if ((i & cmask) == 0) {
      if (lane_g == 0) trace_stamp(p, b, BWD, 1, c, 0);
      mbar_wait(&sv.full[c & (kNumChunks - 1)], (c / kNumChunks) & 1);
      if (lane_g == 0) trace_stamp(p, b, BWD, 1, c, 1);
    }
    int en = e;
    if (!BWD || i > 0) {
      if (NW > 1) named_barrier(1, NW * 32);
      const Boundary* brd = sv.bnd + (i & 1) * 8;
      Boundary* bwr = sv.bnd + ((i + 1) & 1) * 8;
      const double* Erow = sv.E + (size_t)((BWD ? i - 1 : i) & ring_mask) * p.lstride;
      const double pb = Erow[bcol];
      double pl[H];
#pragma unroll
      for (int h = 0; h < H; h++) pl[h] = Erow[ecol[h]];
      if (!BWD) {
        double bx = __shfl_up_sync(FULL, x[K - 1], 1);
        int be = __shfl_up_sync(FULL, e, 1);
        if (lane == 0) {
          if (NW > 1 && w > 0) { bx = brd[w - 1].x0; be = brd[w - 1].e; }
          else be = kNegExp;
        }
        const int eo = e - sh;
        en = max(eo, be);
        const double fo = pow2i(e - en), fb = pow2i(be - en);
        const double bxs = bx * fb;
        const double pbf = pb * fo;
        double plf[H];
#pragma unroll
        for (int h = 0; h < H; h++) plf[h] = pl[h] * fo;
        // cell j gathers j, j-1 and (label cells, when allowed) j-2 of the previous frame; in place, top down
#pragma unroll
        for (int j = K - 1; j >= 2; j--) {
          double so = x[j] + x[j - 1];
          if (j & 1) so = fma(skipd[j >> 1], x[j - 2], so);
          x[j] = so * ((j & 1) ? plf[j >> 1] : pbf);
        }
        x[1] = pl[0] * fma(fo, x[1] + x[0], skipd[0] * bxs);
        x[0] = pb * fma(fo, x[0], bxs);
      } else {
        double by0 = __shfl_down_sync(FULL, x[0], 1);
        double by1 = __shfl_down_sync(FULL, x[1], 1);
        int be = __shfl_down_sync(FULL, e, 1);
        if (lane == 31) {
          if (NW > 1 && w + 1 < NW) { by0 = brd[w + 1].x0; by1 = brd[w + 1].x1; be = brd[w + 1].e; }
          else be = kNegExp;
        }
        const double pln = Erow[nbcol];
        const int eo = e - sh;
        en = max(eo, be);
        const double fo = pow2i(e - en), fb = pow2i(be - en);
        const double pbf = pb * fo;
        const double bu0 = (pb * fb) * by0, bu1 = (pln * fb) * by1;
        // state is beta BEFORE the emission of its frame (ctc_loss.cpp:84-85): multiply by the
        // emissions of frame t+1, then cell j gathers j, j+1 and (label cells, when allowed) j+2
#pragma unroll
        for (int j = 0; j < K; j++) x[j] *= (j & 1) ? (pl[j >> 1] * fo) : pbf;
#pragma unroll
        for (int j = 0; j < K - 2; j++) {
          double so = x[j] + x[j + 1];
          if (j & 1) so = fma(skipd[j >> 1], x[j + 2], so);
          x[j] = so;
        }
        x[K - 2] = x[K - 2] + x[K - 1];
        x[K - 1] = fma(skipd[H - 1], bu1, x[K - 1] + bu0);
      }
      // lagged lane-local renormalisation: next frame scales by 2^sh so the block maximum is in [1,2)
      int mhi = 0;
#pragma unroll
      for (int j = 0; j < K; j++) mhi = max(mhi, __double2hiint(x[j]));
      if (mhi == 0) { e = kNegExp; sh = 0; }
      else { e = en; sh = 1023 - (mhi >> 20); }
      if (NW > 1) {
        if (!BWD) { if (lane == 31) { bwr[w].x0 = x[K - 1]; bwr[w].e = e; } }
        else { if (lane == 0) { bwr[w].x0 = x[0]; bwr[w].x1 = x[1]; bwr[w].e = e; } }
      }
    }
    // drop the frame's state into the ring for the combiners: top 32 bits of every cell + exponent
    {
      uint32_t* ent = sv.val + ((size_t)(i & ring_mask) * LANES + lane_g) * WORDS;
      uint32_t wds[WORDS];
#pragma unroll
      for (int j = 0; j < K; j++) wds[j] = (uint32_t)__double2hiint(x[j]);
      wds[K] = (uint32_t)en;
#pragma unroll
      for (int j = K + 1; j < WORDS; j++) wds[j] = 0u;
#pragma unroll
      for (int u = 0; u < WORDS / 4; u++)
        reinterpret_cast<uint4*>(ent)[u] = make_uint4(wds[4 * u], wds[4 * u + 1], wds[4 * u + 2], wds[4 * u + 3]);
    }
    // hand the chunk to the combiners.  The backward sweep reads the emissions of frame i-1 at
    // iteration i, so it releases a chunk one iteration late.
    if (!BWD) {
      if ((i & cmask) == cmask || i == Ti - 1) { __syncwarp(); if (lane == 0) mbar_arrive(&sv.latdone[c & (kNumChunks - 1)]); }
    } else {
      if (i > 0 && (i & cmask) == 0) { __syncwarp(); if (lane == 0) mbar_arrive(&sv.latdone[(c - 1) & (kNumChunks - 1)]); }
    }
  }
  if (BWD) { __syncwarp(); if (lane == 0) mbar_arrive(&sv.latdone[((Ti - 1) >> cs) & (kNumChunks - 1)]); }

  // loss = -log(alpha[S-1][T-1] + alpha[S-2][T-1]) (ctc_loss.cpp:63-70), from the live fp64 forward
  // state.  Emissions were normalised per row, so for log-prob input the row normalisers (all ~0
  // for true log-probabilities) are added back.
  if (!BWD) {
    double tail = 0.0;
#pragma unroll
    for (int j = 0; j < K; j++) if (s0 + j == S - 1 || s0 + j == S - 2) tail += x[j];
    double z;
    int ez;
    sweep_sum_scaled<NW>(tail, e, sv, w, lane, &z, &ez);
    if (w == 0 && lane == 0) {
      double loss = INFINITY;
      if (z > 0.0) {
        loss = -(log(z) + (double)ez * 0.69314718055994530942);
        if (!p.from_logits) { for (int q = 0; q < kMaxProducerWarps; q++) loss -= sv.lsesum[q]; }
      } else {   // no path survives (exact-zero emissions): +inf; NaN input: NaN.  NaN gradient block either way
        if (z != z) loss = NAN;
        p.flags[b] = kFlagInfeasible;
        sv.misc[2] = 1;
      }
      store_from_double(p.losses, p.dtype, b, loss);
    }
  }
}

// ---- the combiner warps -----------------------------------------------------------------------------
// Frames round-robin over the nc combiner warps (warp q owns frames i = q mod nc).
template <int K, int NW, bool BWD>
__device__ void run_combiner(const LatticeParams& p, const SmemView& sv, int b, int Ti, int Li, int q, int lane) {
  constexpr int H = K / 2;
  constexpr int WORDS = (K + 1 + 3) & ~3;
  constexpr int LANES = 32 * NW;
  constexpr int CELLS = LANES * K;
  constexpr int ROWW = LANES * WORDS;          // u32 words per frame row
  const int nc = p.nc, pfd = p.pfd;
  const int cs = p.chunk_log2, CF = 1 << cs;
  const int ring_mask = (kNumChunks << cs) - 1;
  const int tm = Ti / 2;
  const int nstore = BWD ? (Ti - tm) : tm;   // frames this sweep stores; the rest it combines
  const int nchunks = (Ti + CF - 1) >> cs;
  uint32_t* const stash_u = p.stash + (size_t)b * p.T * ROWW + (size_t)lane * WORDS;   // + t*ROWW + g*32*WORDS
  uint32_t* const mystage = sv.stage + (size_t)q * pfd * ROWW + (size_t)lane * WORDS;
  const uint32_t* const myval = sv.val + (size_t)lane * WORDS;

  auto frame_t = [&](int i) { return BWD ? (Ti - 1 - i) : i; };
  auto prefetch = [&](int i2, int ord) {   // the other sweep's stored row of frame i2 -> staging slot ord % pfd
    if (i2 < Ti) {
      uint32_t* dst = mystage + (size_t)(ord % pfd) * ROWW;
      const uint32_t* src = stash_u + (size_t)frame_t(i2) * ROWW;
#pragma unroll
      for (int g = 0; g < NW; g++)
#pragma unroll
        for (int u = 0; u < WORDS / 4; u++) cp_async_cg16(dst + g * 32 * WORDS + 4 * u, src + g * 32 * WORDS + 4 * u);
    }
    cp_async_commit();
  };
  auto meet = [&]() {
    // publish this warp's stored rows (global stores, ordered by the warp barrier, released at
    // cluster scope by lane 0's remote arrive) on the PEER CTA's mbarrier; acquire the peer's on ours
    __threadfence();
    __syncwarp();
    if (lane == 0) mbar_arrive_peer(sv.meet, BWD ? 0u : 1u);
    mbar_wait_cluster(sv.meet, 0);
  };

  bool met = false, have_z = false;
  double invz = 0.0;
  int Ez = 0, ord = 0;                         // ord: ordinal of my next combine frame
  for (int c = 0; c < nchunks; ++c) {
    const int slot_c = c & (kNumChunks - 1), par = (c / kNumChunks) & 1;
    if (q == 0 && lane == 0) trace_stamp(p, b, BWD, 2, c, 0);
    mbar_wait(&sv.full[slot_c], par);
    mbar_wait(&sv.latdone[slot_c], par);
    if (q == 0 && lane == 0) trace_stamp(p, b, BWD, 2, c, 1);
    const int iend = min(Ti, (c + 1) << cs);
    for (int i = (c << cs) + ((q - (c << cs)) & (nc - 1)); i < iend; i += nc) {
      const int t = frame_t(i);
      const int fr = i & ring_mask;
      const uint32_t* vrow = myval + (size_t)fr * ROWW;
      if (i < nstore) {
        uint32_t* dst = stash_u + (size_t)t * ROWW;
#pragma unroll
        for (int g = 0; g < NW; g++)
#pragma unroll
          for (int u = 0; u < WORDS / 4; u++)
            reinterpret_cast<uint4*>(dst + g * 32 * WORDS)[u] = reinterpret_cast<const uint4*>(vrow + g * 32 * WORDS)[u];
        continue;
      }
      if (!met) {
        meet();
        met = true;
        for (int u = 0; u < pfd; u++) prefetch(i + u * nc, u);
      }
      if (pfd == 4) cp_async_wait<3>(); else if (pfd == 2) cp_async_wait<1>(); else cp_async_wait<0>();
      const uint32_t* orow = mystage + (size_t)(ord % pfd) * ROWW;
      if (!have_z) {
        // Z = sum_s alpha(t,s) * beta(t,s), the same for every frame t: taken once per warp
        double lsum[NW];
        int El[NW], emax = 4 * kNegExp;
#pragma unroll
        for (int g = 0; g < NW; g++) {
          const uint32_t* ve = vrow + g * 32 * WORDS;
          const uint32_t* oe = orow + g * 32 * WORDS;
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < K; j++) s = fma(unpack_hi32(ve[j]), unpack_hi32(oe[j]), s);
          lsum[g] = s;
          El[g] = (int)ve[K] + (int)oe[K];
          if (s > 0.0) emax = max(emax, El[g]);
        }
        emax = warp_max_int(emax);
        double tot = 0.0;
#pragma unroll
        for (int g = 0; g < NW; g++) tot += lsum[g] > 0.0 ? lsum[g] * pow2i(El[g] - emax) : 0.0;
        tot = warp_sum(tot);
        // 1/Z (times 2^31 in dense mode) as a mantissa in [1,2) with its power of two folded into the reference
        // exponent, so that the per-lane scale stays finite for lanes that carry a stale exponent and no mass
        // (0 * inf = NaN converted to a spurious posterior; see ctc_sweep_impl.cuh).  tot == 0 or NaN: garbage
        // posteriors; the forward sweep flags the utterance and the block is overwritten with NaN.
        const double rz = (p.dense ? 2147483648.0 : 1.0) / tot;
        const int kz = ((__double2hiint(rz) >> 20) & 0x7ff) - 1023;
        invz = __hiloint2double((__double2hiint(rz) & 0x800fffff) | 0x3ff00000, __double2loint(rz));
        Ez = emax - kz;
        have_z = true;
      }
      uint32_t* arow = sv.acc + (size_t)fr * p.vpad;
      if (p.dense) {
        // posteriors summed per symbol with integer shared-memory atomics (fixed point 2^-31: the sum
        // per symbol is <= 1; integer adds commute, so the gradient is bitwise reproducible)
        uint32_t qblank = 0u;
#pragma unroll
        for (int g = 0; g < NW; g++) {
          const uint32_t* ve = vrow + g * 32 * WORDS;
          const uint32_t* oe = orow + g * 32 * WORDS;
          const double ccl = pow2i((int)ve[K] + (int)oe[K] - Ez) * invz;
          const int* labp = sv.lab + (g * 32 + lane) * H;
          double bs = 0.0;
#pragma unroll
          for (int u = 0; u < H; u++) {
            const uint2 vv = *reinterpret_cast<const uint2*>(ve + 2 * u);
            const uint2 ov = *reinterpret_cast<const uint2*>(oe + 2 * u);
            bs = fma(unpack_hi32(vv.x), unpack_hi32(ov.x), bs);                       // blank cell 2u
            const double pl = unpack_hi32(vv.y) * unpack_hi32(ov.y) * ccl;           // label cell 2u+1
            atomicAdd(arow + labp[u], __double2uint_rn(pl));                          // labels past L_i are padded with blank, cells there are 0
          }
          qblank += __double2uint_rn(bs * ccl);
        }
        atomicAdd(arow + p.blank, qblank);
        // gradient row: scale * (softmax - posterior) (ctc_loss.cpp:116-117 + log_softmax backward);
        // log-prob input: exp(lp) - posterior (the engine contract)
        __syncwarp();
        const double* Erow = sv.E + (size_t)fr * p.lstride;
        const long long gbase = (long long)b * p.gsb + (long long)t * p.gst;
        if (p.dtype == E2E_F64) {
          const double rs = p.from_logits ? 1.0 : Erow[p.V + 1];
          double* g = reinterpret_cast<double*>(p.grads) + gbase;
          for (int v = lane; v < p.V; v += 32) g[v] = p.scale * (Erow[v] * rs - (double)arow[v] * (1.0 / 2147483648.0));
        } else {
          const float rs = p.from_logits ? 1.f : (float)Erow[p.V + 1], sc = (float)p.scale;
          for (int v = lane; v < p.V; v += 32) {
            const float gv = sc * ((float)Erow[v] * rs - (float)arow[v] * (1.f / 2147483648.f));
            if (p.dtype == E2E_F32) reinterpret_cast<float*>(p.grads)[gbase + v] = gv;
            else if (p.dtype == E2E_BF16) reinterpret_cast<__nv_bfloat16*>(p.grads)[gbase + v] = __float2bfloat16_rn(gv);
            else reinterpret_cast<__half*>(p.grads)[gbase + v] = __float2half_rn(gv);
          }
        }
      } else {
        // compact posterior row for the gradient kernel: [label 0 .. label cells/2-1 | blank total]
        float* prow = p.post + ((size_t)b * p.T + t) * (size_t)p.post_stride;
        double bsum = 0.0;
#pragma unroll
        for (int g = 0; g < NW; g++) {
          const uint32_t* ve = vrow + g * 32 * WORDS;
          const uint32_t* oe = orow + g * 32 * WORDS;
          const double cc = pow2i((int)ve[K] + (int)oe[K] - Ez) * invz;
          float lp[H];
          double bs = 0.0;
#pragma unroll
          for (int u = 0; u < H; u++) {
            const uint2 vv = *reinterpret_cast<const uint2*>(ve + 2 * u);
            const uint2 ov = *reinterpret_cast<const uint2*>(oe + 2 * u);
            bs = fma(unpack_hi32(vv.x), unpack_hi32(ov.x), bs);
            lp[u] = (float)(unpack_hi32(vv.y) * unpack_hi32(ov.y) * cc);
          }
          bsum = fma(bs, cc, bsum);
          float* dst = prow + (size_t)(g * 32 + lane) * H;
          if (H == 1) dst[0] = lp[0];
          else if (H == 2) *reinterpret_cast<float2*>(dst) = make_float2(lp[0], lp[1]);
          else {
#pragma unroll
            for (int u = 0; u < H / 4; u++)
              reinterpret_cast<float4*>(dst)[u] = make_float4(lp[4 * u], lp[4 * u + 1], lp[4 * u + 2], lp[4 * u + 3]);
          }
        }
        bsum = warp_sum(bsum);
        if (lane == 0) prow[CELLS / 2] = (float)bsum;
      }
      prefetch(i + pfd * nc, ord + pfd);
      ++ord;
    }
    __syncwarp();
    if (q == 0 && lane == 0) trace_stamp(p, b, BWD, 2, c, 2);
    if (lane == 0) mbar_arrive(&sv.empty[slot_c]);
  }
  if (!met) meet();
  cp_async_wait<0>();
}

// ---- kernel -----------------------------------------------------------------------------------
// Block = NW lattice warps, nc combiner warps, np producer warps.
template <int K, int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (NW + (K >= 24 ? 2 : kMaxCombinerWarps) + kMaxProducerWarps), 1)
ctc_lattice_kernel(const LatticeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int H = K / 2;
  constexpr int LANES = 32 * NW;
  const SmemView sv = carve(smem_raw, p.sm);

  const int b = blockIdx.x >> 1;
  const bool bwd = cluster_ctarank() == 1;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;

  const long long Ti_ll = load_index(p.in_len, p.len_is64, b);
  const long long Li_ll = load_index(p.tgt_len, p.len_is64, b);
  int bad = 0;
  if (Ti_ll < 1 || Ti_ll > p.T) bad |= kBadFrames;
  if (Li_ll < 0 || Li_ll > p.Lmax) bad |= kBadTargetLen;
  const int Ti = (int)Ti_ll, Li = bad ? 0 : (int)Li_ll;
  if (tid < 4) sv.misc[tid] = 0;
  __syncthreads();
  int rep = 0, badlab = 0;
  for (int i = tid; i < Li; i += blockDim.x) {
    const long long v = load_index(p.targets, p.tgt_is64, (long long)b * p.ts_b + i);
    if (v < 0 || v >= p.V) badlab = kBadLabel;
    sv.lab[i] = (int)v;
  }
  for (int i = Li + tid; i < LANES * H; i += blockDim.x) sv.lab[i] = p.blank;   // cells past the lattice carry zero mass
  __syncthreads();
  for (int i = tid + 1; i < Li; i += blockDim.x) rep += (sv.lab[i] == sv.lab[i - 1]);
  if (rep) atomicAdd(&sv.misc[1], rep);
  if (badlab) atomicOr(&sv.misc[0], badlab);
  __syncthreads();
  bad |= sv.misc[0];
  rep = sv.misc[1];
  const long long gfill_base = (long long)b * p.gsb;
  if (bad || Ti < Li + rep) {
    // out-of-range lengths / labels (undefined behaviour in the reference): NaN loss + status bits;
    // no alignment exists (T < L + repeats): loss = +inf.  Either way the gradient block is all NaN
    // (-inf - (-inf) in the reference, ctc_loss.cpp:116-117), padding rows included.
    if (!bwd && tid == 0) {
      if (bad) atomicOr(p.status, bad);
      p.flags[b] = bad ? kFlagInvalid : kFlagInfeasible;
      store_from_double(p.losses, p.dtype, b, bad ? (double)NAN : (double)INFINITY);
    }
    if (p.dense && p.grads != nullptr) {
      for (int r = (bwd ? 1 : 0) + 2 * w; r < p.T; r += 2 * (blockDim.x >> 5))
        for (int v = lane; v < p.V; v += 32) store_from_double(p.grads, p.dtype, gfill_base + (long long)r * p.gst + v, (double)NAN);
    }
    return;
  }
  if (tid == 0) {
    if (!bwd) p.flags[b] = 0;
    for (int c = 0; c < kNumChunks; c++) {
      mbar_init(&sv.full[c], p.np * 32);
      mbar_init(&sv.latdone[c], NW);
      mbar_init(&sv.empty[c], p.nc);
    }
    mbar_init(sv.meet, p.nc);
  }
  {  // emission ring: every column the producers never write must read as zero
    const int ring = kNumChunks << p.chunk_log2;
    const int n = ring * p.lstride;
    for (int k = tid; k < n; k += blockDim.x) sv.E[k] = 0.0;
    for (int q = tid; q < 16; q += blockDim.x) { sv.bnd[q].x0 = 0.0; sv.bnd[q].x1 = 0.0; sv.bnd[q].e = kNegExp; }
    if (tid < kMaxProducerWarps) sv.lsesum[tid] = 0.0;
  }
  __syncthreads();
  // Both CTAs of the pair took the same early-exit decisions above, so both reach this point:
  // the peer's mbarriers exist before anyone arrives on them remotely.
  cluster_arrive();
  cluster_wait();

  // Warp roles.  The SMSP arbiter favours the highest warp id among eligible warps, so the lattice
  // warps (the dependent chain) get the highest ids, the combiners the next, the producers the lowest.
  const int first_comb = p.np, first_lat = p.np + p.nc;
  if (w >= first_lat) {
    if (bwd) run_lattice<K, NW, true>(p, sv, b, Ti, Li, w - first_lat, lane);
    else run_lattice<K, NW, false>(p, sv, b, Ti, Li, w - first_lat, lane);
  } else if (w >= first_comb) {
    if (bwd) run_combiner<K, NW, true>(p, sv, b, Ti, Li, w - first_comb, lane);
    else run_combiner<K, NW, false>(p, sv, b, Ti, Li, w - first_comb, lane);
  } else {
    const int pw = w;
    if (p.dtype == E2E_F64) {
      if (bwd) run_producer<true, true, H, LANES>(p, sv, b, Ti, Li, pw, lane);
      else run_producer<false, true, H, LANES>(p, sv, b, Ti, Li, pw, lane);
    } else {
      if (bwd) run_producer<true, false, H, LANES>(p, sv, b, Ti, Li, pw, lane);
      else run_producer<false, false, H, LANES>(p, sv, b, Ti, Li, pw, lane);
    }
    if (p.dense && p.grads != nullptr) {
      // padding frames t >= T_i: exp(lp) for log-prob input (the engine contract, ctc_loss.cpp:105-117),
      // 0 for fused-logits input (what the reference's log_softmax backward leaves there)
      for (int r = Ti + (bwd ? 1 : 0) + 2 * pw; r < p.T; r += 2 * p.np) {
        const long long xo = (long long)b * p.sb + (long long)r * p.st;
        const long long go = gfill_base + (long long)r * p.gst;
        for (int v = lane; v < p.V; v += 32) {
          double g = 0.0;
          if (!p.from_logits) {
            g = p.dtype == E2E_F64 ? exp(load_as_double(p.logits, p.dtype, xo + v)) : (double)expf(load_as_float(p.logits, p.dtype, xo + v));
          }
          store_from_double(p.grads, p.dtype, go + v, p.scale * g);
        }
      }
    }
  }
  if (p.dense && p.grads != nullptr) {
    // Z == 0 although an alignment exists (exact-zero emissions): the reference yields +inf / NaN.
    // The forward CTA knows; once both CTAs are done it overwrites the block.
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    if (!bwd && sv.misc[2]) {
      for (int r = w; r < p.T; r += (blockDim.x >> 5))
        for (int v = lane; v < p.V; v += 32) store_from_double(p.grads, p.dtype, gfill_base + (long long)r * p.gst + v, (double)NAN);
    }
  }
}

template <int K, int NW>
int launch_k(const LatticeParams& lp, const LossPlan& p, cudaStream_t s) {
  static int attr_smem = -1;   // the attribute only ever grows
  if ((int)p.smem > attr_smem) {
    E2E_CUDA_TRY(cudaFuncSetAttribute(ctc_lattice_kernel<K, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    attr_smem = (int)p.smem;
  }
  const unsigned threads = 32u * (unsigned)(NW + p.nc + p.np);
  KernelTimer timer(kKernelLattice, s);
  ctc_lattice_kernel<K, NW><<<2u * (unsigned)lp.B, threads, p.smem, s>>>(lp);
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace
}  // namespace e2e
