// K8 -- LM-free CTC prefix beam search.
//
// Replaces src/decoders/ctc_decoder.cpp: decode (:153-198), decode_sentence (:353-441), get_next_prefix (:241-309, the
// branch without a language model), Prefix::next_step / get_prev_full_prob (:331-340), get_prev_full_prob_with_lmwt
// (:311-315) and Prefix::get_sentence (:225-239), plus the F.log_softmax of decoders/ctc_decoder.py:95-97 when the
// input is raw logits.
//
// One CTA per utterance; the frames are a dependent chain, the work inside a frame is parallel:
//   * the beam (<= beam_width prefixes: trie node, last symbol, word count, log p(blank), log p(not blank)) lives in
//     shared memory, double buffered;
//   * the reference makes a Prefix object for each of the beam x (V-1) extensions of a frame and throws all but
//     beam_width of them away (std::nth_element).  Here an extension is never materialised: its score is one add of
//     the frame's log-probability to a per-member base, recomputed wherever it is needed;
//   * the beam_width best of the beam members + extensions are found by an MSB-first radix select on an order-preserving
//     64-bit image of the fp64 score (8-bit digits, stops as soon as a digit bin holds exactly what is still needed);
//   * survivors are compacted in position order (members in beam order, then extensions by (member, symbol)) with
//     ballot / popc ranks and one scan over the per-member counts: bitwise reproducible.
//
// The reference finds an existing child through a weak_ptr in its parent (:244-246).  What that does is kept exactly:
// the prefix trie lives in the caller's workspace with a reference count per node (one for membership of the beam, one per
// living child = the shared_ptr holders `prefixes` and `Prefix::parent`), a child list per node, and
//   - an extension whose node is in the beam adds its mass to that member (is_new == false),
//   - an extension whose node was pruned but is still referenced by a descendant in the beam is swallowed: that prefix
//     cannot re-enter the beam while the descendant lives (a property of the reference, reproduced on purpose),
//   - otherwise the extension is a fresh prefix.
// Equal scores on both sides of the prune cut are resolved by libstdc++'s introselect in the reference; here the lower
// position wins and the utterance's tie counter is raised, so a caller can tell when the result depended on that.
//
// Arithmetic: fp64 log-space with the reference's two-argument log_sum_exp (src/utils/math_utils.h:8-16) and the score
// expression of :311-315 evaluated in its order without fused multiply-adds.
#include "common.cuh"

namespace e2e {
namespace {

struct __align__(32) BeamNode {   // 32 bytes: one sector
  int parent, chr, refs, slot, first_child, next_sib, depth, pad;
};

struct BeamParams {
  const void* logits; int dtype; long long sb, st;
  const void* in_len; int len_is64;
  long long* decoded;        // [B, T] zero padded
  long long* decoded_len;    // [B]
  long long* ties;           // [B] or null
  BeamNode* nodes;           // [B][node_cap]
  long long node_cap;
  int B, T, V, blank, beam, space, from_logits;
  double wip;
};

constexpr int kBeamThreads = 256;
constexpr int kBeamWarps = kBeamThreads / 32;

__device__ __forceinline__ double beam_lse(double a, double b) {   // math_utils.h:8-16
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  if (a > b) return __dadd_rn(a, log(__dadd_rn(1.0, exp(__dsub_rn(b, a)))));
  return __dadd_rn(b, log(__dadd_rn(1.0, exp(__dsub_rn(a, b)))));
}

// get_prev_full_prob_with_lmwt (:311-315) without a language model: lm_score = 0, lmwt = 0, num_oov_words = 0
__device__ __forceinline__ double beam_score(double full, int num_words, double wip) {
  return __dsub_rn(__dadd_rn(full, 0.0), __dmul_rn((double)num_words, wip));
}

// order-preserving image of a double; +0 and -0 coincide, NaN sorts below everything
__device__ __forceinline__ unsigned long long beam_key(double x) {
  if (x != x) return 0ull;
  x = __dadd_rn(x, 0.0);
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

struct BeamBuf {          // one copy of the beam
  double *pb, *pnb;
  int *node, *last, *nw, *dep;
};
struct BeamSmem {
  double* lp;            // [Vp]
  double *full, *npb, *npnb;
  unsigned long long* mkey;
  int *nwx;              // word count of an extension with a symbol other than space
  int *cgt, *ceq;        // per position group: survivors above the cut / in the cut class (members: [0,W), rows: [W,2W))
  unsigned* bitmap;      // [W][VW] extensions that are not fresh prefixes
  unsigned* hist;        // [256]
};

__device__ __forceinline__ double ext_value(const BeamSmem& S, const BeamBuf& C, int s, int c, int last_s) {
  // :381-390: a repeated character extends from the blank-ending mass only
  return __dadd_rn(S.lp[c], c == last_s ? C.pb[s] : S.full[s]);
}

template <typename T>
__global__ void __launch_bounds__(kBeamThreads) ctc_beam_kernel(const BeamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_red[kBeamWarps];
  __shared__ unsigned long long s_thr;
  __shared__ int s_sh, s_need, s_W, s_nodes, s_ties, s_total, s_done, s_keepm, s_keepall;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = p.V, WB = p.beam, VW = (V + 31) >> 5, Vp = (V + 1) & ~1;

  BeamSmem S;
  BeamBuf C, N;
  {
    unsigned char* q = smem_raw;
    auto take = [&](size_t bytes) { unsigned char* r = q; q += (bytes + 15) & ~(size_t)15; return r; };
    S.lp = reinterpret_cast<double*>(take(sizeof(double) * Vp));
    C.pb = reinterpret_cast<double*>(take(8 * WB)); C.pnb = reinterpret_cast<double*>(take(8 * WB));
    N.pb = reinterpret_cast<double*>(take(8 * WB)); N.pnb = reinterpret_cast<double*>(take(8 * WB));
    S.full = reinterpret_cast<double*>(take(8 * WB));
    S.npb = reinterpret_cast<double*>(take(8 * WB));
    S.npnb = reinterpret_cast<double*>(take(8 * WB));
    S.mkey = reinterpret_cast<unsigned long long*>(take(8 * WB));
    C.node = reinterpret_cast<int*>(take(4 * WB)); C.last = reinterpret_cast<int*>(take(4 * WB));
    C.nw = reinterpret_cast<int*>(take(4 * WB)); C.dep = reinterpret_cast<int*>(take(4 * WB));
    N.node = reinterpret_cast<int*>(take(4 * WB)); N.last = reinterpret_cast<int*>(take(4 * WB));
    N.nw = reinterpret_cast<int*>(take(4 * WB)); N.dep = reinterpret_cast<int*>(take(4 * WB));
    S.nwx = reinterpret_cast<int*>(take(4 * WB));
    S.cgt = reinterpret_cast<int*>(take(4 * 2 * WB));
    S.ceq = reinterpret_cast<int*>(take(4 * 2 * WB));
    S.bitmap = reinterpret_cast<unsigned*>(take((size_t)4 * WB * VW));
    S.hist = reinterpret_cast<unsigned*>(take(4 * 256));
  }

  long long Ti_ll = p.in_len ? load_index(p.in_len, p.len_is64, b) : (long long)p.T;
  const int Ti = (int)max(0LL, min(Ti_ll, (long long)p.T));
  BeamNode* nodes = p.nodes + (long long)b * p.node_cap;
  const T* x = reinterpret_cast<const T*>(p.logits) + (long long)b * p.sb;

  if (tid == 0) {   // get_initial_prefix (:201-209): the empty prefix with log p(blank) = 0
    BeamNode r; r.parent = -1; r.chr = -1; r.refs = 1; r.slot = 0; r.first_child = -1; r.next_sib = -1; r.depth = 0; r.pad = 0;
    nodes[0] = r;
    C.node[0] = 0; C.last[0] = -1; C.nw[0] = 0; C.dep[0] = 0; C.pb[0] = 0.0; C.pnb[0] = -INFINITY;
    s_W = 1; s_nodes = 1; s_ties = 0;
  }
  __syncthreads();

  for (int t = 0; t < Ti; t++) {
    const int W = s_W;
    // ---- the frame's log-probabilities (decoders/ctc_decoder.py:95-97 when the input is raw logits) ----------
    const T* row = x + (long long)t * p.st;
    if (p.from_logits) {
      if (sizeof(T) == 8) {   // fp64 input: the log-softmax in fp64
        double m = -INFINITY;
        for (int c = tid; c < V; c += kBeamThreads) m = fmax(m, (double)Elem<T>::load(row + c));
        m = warp_max(m);
        double* redd = reinterpret_cast<double*>(S.hist);   // 256 words: room for 8 doubles + broadcast
        if (lane == 0) redd[warp] = m;
        __syncthreads();
        m = redd[0];
        for (int w = 1; w < kBeamWarps; w++) m = fmax(m, redd[w]);
        __syncthreads();
        double sum = 0.0;
        for (int c = tid; c < V; c += kBeamThreads) sum += exp((double)Elem<T>::load(row + c) - m);
        sum = warp_sum(sum);
        if (lane == 0) redd[warp] = sum;
        __syncthreads();
        sum = 0.0;
        for (int w = 0; w < kBeamWarps; w++) sum += redd[w];
        const double ls = log(sum);
        for (int c = tid; c < V; c += kBeamThreads) S.lp[c] = ((double)Elem<T>::load(row + c) - m) - ls;
      } else {                // fp32 arithmetic in torch's operation order: (x - max) - log(sum exp(x - max))
        float m = -INFINITY;
        for (int c = tid; c < V; c += kBeamThreads) m = fmaxf(m, (float)Elem<T>::load(row + c));
        m = warp_max(m);
        if (lane == 0) s_red[warp] = m;
        __syncthreads();
        m = s_red[0];
        for (int w = 1; w < kBeamWarps; w++) m = fmaxf(m, s_red[w]);
        __syncthreads();
        float sum = 0.f;
        for (int c = tid; c < V; c += kBeamThreads) sum += expf((float)Elem<T>::load(row + c) - m);
        sum = warp_sum(sum);
        if (lane == 0) s_red[warp] = sum;
        __syncthreads();
        sum = 0.f;
        for (int w = 0; w < kBeamWarps; w++) sum += s_red[w];
        const float ls = logf(sum);
        for (int c = tid; c < V; c += kBeamThreads) {
          float v = ((float)Elem<T>::load(row + c) - m) - ls;
          if (sizeof(T) == 2) { T r; Elem<T>::store(&r, v); v = Elem<T>::get(r); }   // torch returns the input dtype
          S.lp[c] = (double)v;
        }
      }
    } else {
      for (int c = tid; c < V; c += kBeamThreads) S.lp[c] = (double)Elem<T>::load(row + c);
    }
    for (int i = tid; i < W * VW; i += kBeamThreads) S.bitmap[i] = 0u;
    __syncthreads();

    // ---- phase A: every member's own update (:372-376, :383-385) ----------------------------------------------
    if (tid < W) {
      const int s = tid;
      const double pbv = C.pb[s], pnbv = C.pnb[s];
      const double full = beam_lse(pnbv, pbv);                       // get_prev_full_prob (:331-333)
      S.full[s] = full;
      S.npb[s] = beam_lse(-INFINITY, __dadd_rn(S.lp[p.blank], full));
      const int last = C.last[s];
      S.npnb[s] = last >= 0 ? beam_lse(-INFINITY, __dadd_rn(S.lp[last], pnbv)) : -INFINITY;
      const int nws = C.nw[s];
      S.nwx[s] = nws + ((nws == 0 || last == p.space) ? 1 : 0);      // :252-257 for a symbol other than space
    }
    __syncthreads();
    // ---- phase B: living children (:244-246): merge into a member of the beam, or swallow --------------------
    if (tid < W) {
      const int s = tid, pn = C.node[s], last = C.last[s];
      int prev = -1;
      int z = nodes[pn].first_child;
      while (z >= 0) {
        const BeamNode zn = nodes[z];
        if (zn.refs <= 0) {                      // expired weak_ptr: unlink
          if (prev < 0) nodes[pn].first_child = zn.next_sib; else nodes[prev].next_sib = zn.next_sib;
        } else {
          const int c = zn.chr;
          atomicOr(&S.bitmap[s * VW + (c >> 5)], 1u << (c & 31));
          if (zn.slot >= 0) S.npnb[zn.slot] = beam_lse(S.npnb[zn.slot], ext_value(S, C, s, c, last));
          prev = z;
        }
        z = zn.next_sib;
      }
    }
    if (tid == 0) { s_total = 0; s_done = 0; }
    __syncthreads();
    // ---- phase C: member scores; how many prefixes are there after this frame (:392-399) ----------------------
    if (tid < W) S.mkey[tid] = beam_key(beam_score(beam_lse(S.npnb[tid], S.npb[tid]), C.nw[tid], p.wip));
    {
      int excl = 0;
      for (int i = tid; i < W * VW; i += kBeamThreads) excl += __popc(S.bitmap[i]);
      excl = warp_sum(excl);
      if (lane == 0 && excl) atomicAdd(&s_total, excl);
    }
    __syncthreads();
    const int total = W + W * (V - 1) - s_total;
    // ---- phase D: radix select of the beam_width best ------------------------------------------------------------
    unsigned long long thr = 0ull;
    int sh = 0, need = total;
    if (total > WB) {
      int remaining = WB;
      unsigned long long prefix = 0ull;
      for (sh = 56; sh >= 0; sh -= 8) {
        S.hist[tid] = 0u;                        // kBeamThreads == 256 bins
        __syncthreads();
        const unsigned long long pm = sh == 56 ? 0ull : (~0ull << (sh + 8));
        if (tid < W) {
          const unsigned long long k = S.mkey[tid];
          if ((k & pm) == (prefix & pm)) atomicAdd(&S.hist[(unsigned)(k >> sh) & 255u], 1u);
        }
        for (int s = warp; s < W; s += kBeamWarps) {
          const int last = C.last[s];
          const double base = S.full[s], baseb = C.pb[s];
          const double penx = __dmul_rn((double)S.nwx[s], p.wip), pens = __dmul_rn((double)C.nw[s], p.wip);
          for (int c0 = 0; c0 < V; c0 += 32) {
            const int c = c0 + lane;
            bool valid = c < V && c != p.blank;
            if (valid) valid = !((S.bitmap[s * VW + (c0 >> 5)] >> lane) & 1u);
            unsigned long long k = 0ull;
            if (valid) {
              const double v = __dadd_rn(S.lp[c], c == last ? baseb : base);
              k = beam_key(__dsub_rn(__dadd_rn(v, 0.0), c == p.space ? pens : penx));
              valid = (k & pm) == (prefix & pm);
            }
            const unsigned d = (unsigned)(k >> sh) & 255u;
            const unsigned act = __ballot_sync(0xffffffffu, valid);
            if (valid) {
              const unsigned m = __match_any_sync(act, d);
              if (lane == __ffs(m) - 1) atomicAdd(&S.hist[d], (unsigned)__popc(m));
            }
          }
        }
        __syncthreads();
        if (warp == 0) {                          // the bin that holds the remaining-th largest
          unsigned h[8]; unsigned mine = 0;
#pragma unroll
          for (int j = 0; j < 8; j++) { h[j] = S.hist[lane * 8 + j]; mine += h[j]; }
          unsigned above = 0;                     // entries in bins of higher lanes
          {
            unsigned v = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v += u; }
            above = v - mine;
          }
          const bool here = above < (unsigned)remaining && above + mine >= (unsigned)remaining;
          if (here) {
            unsigned acc = above; int d = 7;
            for (; d > 0; d--) { if (acc + h[d] >= (unsigned)remaining) break; acc += h[d]; }
            s_need = remaining - (int)acc;
            s_thr = prefix | ((unsigned long long)(lane * 8 + d) << sh);
            s_done = (h[d] == (unsigned)(remaining - (int)acc)) ? 1 : 0;
            if (sh == 0 && !s_done) s_ties++;    // equal scores on both sides of the cut
          }
        }
        __syncthreads();
        prefix = s_thr; remaining = s_need;
        if (s_done || sh == 0) break;
      }
      thr = prefix; need = remaining;
    }
    const unsigned long long tcls = thr >> sh;
    // ---- phase E: survivors per position group ---------------------------------------------------------------
    if (tid < W) {
      const unsigned long long kc = S.mkey[tid] >> sh;
      S.cgt[tid] = kc > tcls; S.ceq[tid] = kc == tcls;
    }
    for (int s = warp; s < W; s += kBeamWarps) {
      const int last = C.last[s];
      const double base = S.full[s], baseb = C.pb[s];
      const double penx = __dmul_rn((double)S.nwx[s], p.wip), pens = __dmul_rn((double)C.nw[s], p.wip);
      int ngt = 0, neq = 0;
      for (int c0 = 0; c0 < V; c0 += 32) {
        const int c = c0 + lane;
        bool valid = c < V && c != p.blank;
        if (valid) valid = !((S.bitmap[s * VW + (c0 >> 5)] >> lane) & 1u);
        unsigned long long kc = 0ull;
        if (valid) {
          const double v = __dadd_rn(S.lp[c], c == last ? baseb : base);
          kc = beam_key(__dsub_rn(__dadd_rn(v, 0.0), c == p.space ? pens : penx)) >> sh;
        }
        ngt += __popc(__ballot_sync(0xffffffffu, valid && kc > tcls));
        neq += __popc(__ballot_sync(0xffffffffu, valid && kc == tcls));
      }
      if (lane == 0) { S.cgt[W + s] = ngt; S.ceq[W + s] = neq; }
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan over the 2W position groups (members, then member rows)
      int run_gt = 0, run_eq = 0;
      for (int i0 = 0; i0 < 2 * W; i0 += 32) {
        const int i = i0 + lane;
        int g = i < 2 * W ? S.cgt[i] : 0, e = i < 2 * W ? S.ceq[i] : 0;
        int sg = g, se = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int ug = __shfl_up_sync(0xffffffffu, sg, o), ue = __shfl_up_sync(0xffffffffu, se, o);
          if (lane >= o) { sg += ug; se += ue; }
        }
        if (i < 2 * W) { S.cgt[i] = run_gt + sg - g; S.ceq[i] = run_eq + se - e; }
        if (i == W - 1) s_keepm = run_gt + sg + min(run_eq + se, need);   // surviving members
        run_gt += __shfl_sync(0xffffffffu, sg, 31); run_eq += __shfl_sync(0xffffffffu, se, 31);
      }
      if (lane == 0) s_keepall = run_gt + min(run_eq, need);
    }
    __syncthreads();
    const int keepm = s_keepm, keepall = s_keepall, node0 = s_nodes;
    // ---- phase F: write the surviving list (next_step, :397) --------------------------------------------------
    bool dropped = false;
    if (tid < W) {
      const int s = tid;
      const unsigned long long kc = S.mkey[s] >> sh;
      const int gb = S.cgt[s], eb = S.ceq[s];
      const bool keep = kc > tcls || (kc == tcls && eb < need);
      const int n = C.node[s];
      if (keep) {
        const int ns = gb + min(eb, need);
        N.node[ns] = n; N.last[ns] = C.last[s]; N.nw[ns] = C.nw[s]; N.dep[ns] = C.dep[s];
        N.pb[ns] = S.npb[s]; N.pnb[ns] = S.npnb[s];
        nodes[n].slot = ns;
      } else {
        nodes[n].slot = -1;
        dropped = true;
      }
    }
    for (int s = warp; s < W; s += kBeamWarps) {
      const int last = C.last[s];
      const double base = S.full[s], baseb = C.pb[s];
      const double penx = __dmul_rn((double)S.nwx[s], p.wip), pens = __dmul_rn((double)C.nw[s], p.wip);
      int gb = S.cgt[W + s], eb = S.ceq[W + s];
      const int pn = C.node[s], pd = C.dep[s];
      for (int c0 = 0; c0 < V; c0 += 32) {
        const int c = c0 + lane;
        bool valid = c < V && c != p.blank;
        if (valid) valid = !((S.bitmap[s * VW + (c0 >> 5)] >> lane) & 1u);
        unsigned long long kc = 0ull;
        double v = 0.0;
        if (valid) {
          v = __dadd_rn(S.lp[c], c == last ? baseb : base);
          kc = beam_key(__dsub_rn(__dadd_rn(v, 0.0), c == p.space ? pens : penx)) >> sh;
        }
        const bool isgt = valid && kc > tcls, iseq = valid && kc == tcls;
        const unsigned mg = __ballot_sync(0xffffffffu, isgt), me = __ballot_sync(0xffffffffu, iseq);
        const unsigned below = (1u << lane) - 1u;
        const int g = gb + __popc(mg & below), e = eb + __popc(me & below);
        if (isgt || (iseq && e < need)) {
          const int ns = g + min(e, need);
          const int id = node0 + (ns - keepm);
          N.node[ns] = id; N.last[ns] = c; N.dep[ns] = pd + 1;
          N.nw[ns] = c == p.space ? C.nw[s] : S.nwx[s];
          N.pb[ns] = -INFINITY; N.pnb[ns] = beam_lse(-INFINITY, v);
          BeamNode r; r.parent = pn; r.chr = c; r.refs = 1; r.slot = ns; r.depth = pd + 1; r.pad = 0; r.first_child = -1;
          r.next_sib = atomicExch(&nodes[pn].first_child, id);
          nodes[id] = r;
          atomicAdd(&nodes[pn].refs, 1);
        }
        gb += __popc(mg); eb += __popc(me);
      }
    }
    __syncthreads();
    // ---- phase G: members that left the beam release their node; a node nobody holds releases its parent --------
    if (dropped) {
      int n = C.node[tid];
      while (n >= 0) {
        const int r = atomicSub(&nodes[n].refs, 1) - 1;
        if (r > 0) break;
        n = nodes[n].parent;
      }
    }
    if (tid == 0) { s_W = keepall; s_nodes = node0 + (keepall - keepm); }
    __syncthreads();
    { const BeamBuf tmp = C; C = N; N = tmp; }
  }

  // ---- the best prefix (:413-419) and its symbols (get_sentence, :225-239) -----------------------------------------
  const int W = s_W;
  if (tid < W) S.mkey[tid] = beam_key(beam_score(beam_lse(C.pnb[tid], C.pb[tid]), C.nw[tid], p.wip));
  __syncthreads();
  long long* out = p.decoded + (long long)b * p.T;
  if (warp == 0) {
    unsigned long long bk = 0ull; int bi = 0x7fffffff;
    for (int s = lane; s < W; s += 32) { const unsigned long long k = S.mkey[s]; if (bi == 0x7fffffff || k > bk) { bk = k; bi = s; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ok > bk || (ok == bk && oi < bi))) { bk = ok; bi = oi; }
    }
    int same = 0;
    for (int s = lane; s < W; s += 32) same += (S.mkey[s] == bk);
    same = warp_sum(same);
    if (lane == 0) {
      if (same > 1) s_ties++;
      const int depth = C.dep[bi];
      int n = C.node[bi];
      int len;
      if (depth == 0) { out[0] = -1; len = 1; }          // the empty prefix: get_sentence pushes the root's last_char = -1
      else {
        len = depth;
        for (int i = depth - 1; i >= 0; i--) { const BeamNode q = nodes[n]; out[i] = q.chr; n = q.parent; }
      }
      p.decoded_len[b] = len;
      if (p.ties) p.ties[b] = s_ties;
      s_need = len;
    }
  }
  __syncthreads();
  for (int k = s_need + tid; k < p.T; k += kBeamThreads) out[k] = 0;
}

size_t beam_smem_bytes(int V, int WB) {
  auto a16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const int VW = (V + 31) >> 5, Vp = (V + 1) & ~1;
  size_t n = a16(sizeof(double) * Vp);
  n += 4 * a16(8 * (size_t)WB);        // pb, pnb x 2
  n += 4 * a16(8 * (size_t)WB);        // full, npb, npnb, mkey
  n += 8 * a16(4 * (size_t)WB);        // node, last, nw, dep x 2
  n += a16(4 * (size_t)WB);            // nwx
  n += 2 * a16(4 * 2 * (size_t)WB);    // cgt, ceq
  n += a16((size_t)4 * WB * VW);
  n += a16(4 * 256);
  return n;
}

}  // namespace

size_t beam_workspace_bytes(const e2e_ctc_desc& d, int beam_width) {
  const size_t cap = (size_t)d.max_frames * (size_t)beam_width + 1;
  return (size_t)d.batch * cap * sizeof(BeamNode);
}

bool beam_supported(const e2e_ctc_desc& d, int beam_width) {
  return beam_width >= 1 && beam_width <= kBeamThreads && beam_smem_bytes(d.alphabet, beam_width) <= 200 * 1024;
}

int launch_beam(const e2e_ctc_desc& d, int beam_width, int space_idx, double wip, const void* logits, const void* in_len,
                int64_t* decoded, int64_t* decoded_len, int64_t* ties, char* ws, cudaStream_t s) {
  BeamParams p;
  p.logits = logits; p.dtype = d.dtype; p.sb = d.logits_stride_b; p.st = d.logits_stride_t;
  p.in_len = in_len; p.len_is64 = d.lengths_itype == E2E_I64;
  p.decoded = reinterpret_cast<long long*>(decoded); p.decoded_len = reinterpret_cast<long long*>(decoded_len);
  p.ties = reinterpret_cast<long long*>(ties);
  p.nodes = reinterpret_cast<BeamNode*>(ws);
  p.node_cap = (long long)d.max_frames * beam_width + 1;
  p.B = d.batch; p.T = d.max_frames; p.V = d.alphabet; p.blank = d.blank_idx; p.beam = beam_width; p.space = space_idx;
  p.from_logits = d.from_logits; p.wip = wip;
  const size_t smem = beam_smem_bytes(d.alphabet, beam_width);
  KernelTimer timer(kKernelBeam, s);
#define E2E_K8(TYPE)                                                                                         \
  do {                                                                                                       \
    E2E_CUDA_TRY(cudaFuncSetAttribute(ctc_beam_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ctc_beam_kernel<TYPE><<<(unsigned)d.batch, kBeamThreads, smem, s>>>(p);                                  \
  } while (0)
  switch (d.dtype) {
    case E2E_F32: E2E_K8(float); break;
    case E2E_BF16: E2E_K8(__nv_bfloat16); break;
    case E2E_F16: E2E_K8(__half); break;
    case E2E_F64: E2E_K8(double); break;
    default: set_error("beam: unsupported dtype %d", d.dtype); return E2E_ERR_INVALID_ARGUMENT;
  }
#undef E2E_K8
  E2E_CUDA_TRY(cudaGetLastError());
  return E2E_OK;
}

}  // namespace e2e
