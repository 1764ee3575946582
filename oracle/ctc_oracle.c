/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, float64) of the reference's CTC hot path.
 *
 * Nothing in the product path (end2end_b200/, pytorch_end2end/) may include, link, import or
 * execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the reported CPU baseline.
 *
 * Parity status: PINNED.  oracle/test vectors: the five known-answer losses of the reference's
 * tests/test_ctc.py:69-165, the greedy and prefix-beam known answers of tests/test_ctc_decoder.py:44-59,86-166,
 * and differential runs against the compiled, unmodified reference (oracle/_ref, built by
 * oracle/build_ref.py) -- see tests/test_oracle.py and tests/golden/.
 *
 * What is restated (reference file:line):
 *   lse2()                 <- src/utils/math_utils.h:8-16      two-argument log-sum-exp, -inf guards,
 *                                                              max + log(1 + exp(min - max))
 *   ctc_oracle_utterance() <- src/losses/ctc_loss.cpp:15-118   blank-extended labels (:25-31),
 *                                                              alpha (:33-61), loss (:63-70),
 *                                                              beta (:72-100), gradient (:102-117)
 *   ctc_oracle_batch()     <- src/losses/forward_backward.cpp:7-59  per-utterance driver (the
 *                                                              reference's one-thread-per-utterance
 *                                                              pool becomes a plain loop / OpenMP)
 *   ctc_oracle_greedy()    <- src/decoders/ctc_decoder.cpp:443-490  argmax (first maximum, NaN is
 *                                                              maximal -- torch's rule) then
 *                                                              collapse repeats / drop blanks
 *   ctc_oracle_log_softmax_f32() <- pytorch_end2end/modules/ctc_loss.py:40 (F.log_softmax, fp32)
 *   ctc_oracle_noblank()   <- pytorch_end2end/functions/ctc_without_blank.py:13-88 (_ctc_without_blank_loss) and
 *                                                              :91-117 (batch driver); pinned by tests/golden/noblank_*.npz
 *   ctc_oracle_align()     <- pytorch_end2end/utils/alignment.py:50-106 (_get_alignment_ctc_1d), :9-47
 *                                                              (_get_alignment_asg_1d), :109-138 (batch driver,
 *                                                              -100 fill); pinned by tests/golden/align_*.npz,
 *                                                              made by running the reference's numba code
 *   ctc_oracle_beam()      <- src/decoders/ctc_decoder.cpp:153-198 (decode), :353-441 (decode_sentence), :241-309
 *                                                              (get_next_prefix without a language model), :311-340
 *                                                              (scores), :225-239 (get_sentence); the shared_ptr /
 *                                                              weak_ptr ownership is restated with reference counts;
 *                                                              pinned by tests/golden/beam_*.npz, made by running the
 *                                                              compiled reference (make_beam_golden.py), and by a
 *                                                              differential test wherever oracle/_ref travelled
 *
 * Storage is frame-major ([T][S]) here, where the reference keeps [S][T]; the arithmetic and the
 * order of the chained two-argument log-sum-exp calls are the reference's.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)

static double lse2(double a, double b) {
  if (a == NEG_INF) return b;
  if (b == NEG_INF) return a;
  if (a > b) return a + log(1.0 + exp(b - a));
  return b + log(1.0 + exp(a - b));
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/*
 * One utterance.  lp: [T_pad][V] log-probabilities (row stride V); frames >= T belong to padding.
 * grad: [T_pad][V], receives exp(lp) - exp(label-summed alpha+beta - logZ) on ALL T_pad rows
 * (padding rows therefore hold exp(lp); an infeasible utterance gives NaN everywhere), as the
 * reference does (src/losses/ctc_loss.cpp:105-117).  Returns the loss.
 */
double ctc_oracle_utterance(const double* lp, int T_pad, int V, const int64_t* labels, int T, int L,
                            int blank, double* grad) {
  const int S = 2 * L + 1;
  int64_t* ext = (int64_t*)malloc(sizeof(int64_t) * (size_t)S);
  double* alpha = (double*)malloc(sizeof(double) * (size_t)S * (size_t)T);
  double* beta = (double*)malloc(sizeof(double) * (size_t)S * (size_t)T);
  double* psum = (double*)malloc(sizeof(double) * (size_t)T_pad * (size_t)V);
  for (int s = 0; s < S; s++) ext[s] = (s & 1) ? labels[s / 2] : (int64_t)blank;
  for (size_t i = 0; i < (size_t)S * T; i++) alpha[i] = beta[i] = NEG_INF;
  for (size_t i = 0; i < (size_t)T_pad * V; i++) psum[i] = NEG_INF;
#define A(t, s) alpha[(size_t)(t) * S + (s)]
#define Bt(t, s) beta[(size_t)(t) * S + (s)]
#define LP(t, v) lp[(size_t)(t) * V + (v)]

  /* alpha, ctc_loss.cpp:39-61 */
  if (T > 1 || S == 1) A(0, 0) = LP(0, ext[0]);
  if (S > 1) A(0, 1) = LP(0, ext[1]);
  for (int t = 1; t < T; t++) {
    const int lo = imax(0, S - 2 * (T - t)), hi = imin(2 * t + 2, S);
    for (int s = lo; s < hi; s++) {
      double a = A(t - 1, s);
      if (s > 0) {
        a = lse2(a, A(t - 1, s - 1));
        if (ext[s] != blank && s >= 2 && ext[s - 2] != ext[s]) a = lse2(a, A(t - 1, s - 2));
      }
      A(t, s) = a + LP(t, ext[s]);
    }
  }
  /* loss, ctc_loss.cpp:63-70 */
  double loss;
  if (S > 1) loss = -lse2(A(T - 1, S - 1), A(T - 1, S - 2));
  else loss = -A(T - 1, S - 1);
  const double logz = -loss;

  /* beta (excludes the emission at t), ctc_loss.cpp:72-100 */
  if (T > 1 || S == 1) Bt(T - 1, S - 1) = 0.0;
  if (S > 1) Bt(T - 1, S - 2) = 0.0;
  for (int t = T - 2; t >= 0; t--) {
    const int lo = imax(0, S - 2 * (T - t)), hi = imin(2 * t + 2, S);
    for (int s = lo; s < hi; s++) {
      double b = Bt(t + 1, s) + LP(t + 1, ext[s]);
      if (s < S - 1) {
        b = lse2(b, Bt(t + 1, s + 1) + LP(t + 1, ext[s + 1]));
        if (ext[s] != blank && s + 2 < S && ext[s + 2] != ext[s])
          b = lse2(b, Bt(t + 1, s + 2) + LP(t + 1, ext[s + 2]));
      }
      Bt(t, s) = b;
    }
  }
  /* gradient, ctc_loss.cpp:102-117: label-wise lse of alpha+beta, s outer / t inner */
  for (int s = 0; s < S; s++)
    for (int t = 0; t < T; t++) {
      double* cell = &psum[(size_t)t * V + ext[s]];
      *cell = lse2(*cell, A(t, s) + Bt(t, s));
    }
  for (size_t i = 0; i < (size_t)T_pad * V; i++) grad[i] = exp(lp[i]) - exp(psum[i] - logz);
#undef A
#undef Bt
#undef LP
  free(ext); free(alpha); free(beta); free(psum);
  return loss;
}

/*
 * Batch driver (forward_backward.cpp:7-59).  lp [B][T][V] float64 contiguous, targets [B][Lmax],
 * lengths int64.  losses [B], grads [B][T][V].  Returns 0, or -1 on arguments the reference would
 * read out of bounds on (T_i < 1, T_i > T, L_i > Lmax, label outside [0,V)).
 */
int ctc_oracle_batch(const double* lp, int B, int T, int V, const int64_t* targets, int Lmax,
                     const int64_t* in_len, const int64_t* tgt_len, int blank, double* losses,
                     double* grads) {
  for (int b = 0; b < B; b++) {
    if (in_len[b] < 1 || in_len[b] > T || tgt_len[b] < 0 || tgt_len[b] > Lmax) return -1;
    for (int i = 0; i < tgt_len[b]; i++)
      if (targets[(size_t)b * Lmax + i] < 0 || targets[(size_t)b * Lmax + i] >= V) return -1;
  }
  if (blank < 0 || blank >= V) return -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic)
#endif
  for (int b = 0; b < B; b++)
    losses[b] = ctc_oracle_utterance(lp + (size_t)b * T * V, T, V, targets + (size_t)b * Lmax,
                                     (int)in_len[b], (int)tgt_len[b], blank,
                                     grads + (size_t)b * T * V);
  return 0;
}

/* fp32 row log-softmax: x - max - log(sum(exp(x - max))), all in float, as torch's CPU kernel
 * evaluates it (modules/ctc_loss.py:40).  Summation order differs from torch's vectorised loop,
 * so results agree to an ulp or two, not bitwise. */
void ctc_oracle_log_softmax_f32(const float* x, int64_t rows, int V, float* out) {
  for (int64_t r = 0; r < rows; r++) {
    const float* xr = x + r * V;
    float m = xr[0];
    for (int v = 1; v < V; v++) if (xr[v] > m) m = xr[v];
    float sum = 0.f;
    for (int v = 0; v < V; v++) sum += expf(xr[v] - m);
    const float ls = logf(sum);
    for (int v = 0; v < V; v++) out[r * V + v] = xr[v] - m - ls;
  }
}

/*
 * Greedy decode (ctc_decoder.cpp:443-490).  logits [B][T][V] float64 (the caller widens; argmax
 * is order-preserving under widening).  argmax = first index of the maximum, a NaN beats any
 * number and the first NaN wins (torch.argmax).  out [B][T] zero-filled then packed; out_len [B].
 */
void ctc_oracle_greedy(const double* logits, int B, int T, int V, const int64_t* in_len, int blank,
                       int64_t* out, int64_t* out_len) {
  memset(out, 0, sizeof(int64_t) * (size_t)B * T);
  for (int b = 0; b < B; b++) {
    int64_t prev = blank, n = 0;
    for (int t = 0; t < in_len[b] && t < T; t++) {
      const double* row = logits + ((size_t)b * T + t) * V;
      int64_t best = 0;
      double bv = row[0];
      if (!isnan(bv))
        for (int v = 1; v < V; v++) {
          if (isnan(row[v])) { best = v; break; }
          if (row[v] > bv) { bv = row[v]; best = v; }
        }
      if (best != blank && best != prev) out[(size_t)b * T + n++] = best;
      prev = best;
    }
    out_len[b] = n;
  }
}

/*
 * Viterbi forced alignment, one utterance.  lp: [T][V] log-probabilities as doubles (the reference adds the
 * float32 inputs into a float64 matrix: the same values).  out: [T] label ids.
 * alignment.py:50-106 (is_ctc) / :9-47 (ASG); blank is 0 in the reference, a parameter here.
 */
static void align_utterance(const double* lp, int V, const int64_t* targets, int T, int L, int blank, int is_ctc,
                            int64_t* out) {
  for (int k = 0; k < T; k++) out[k] = 0;                               /* np.zeros(prediction_len) */
  if (T == 0) return;
  if (is_ctc) {
    const int S = 2 * L + 1;
    if (S == 1) return;                                                 /* only blank: zeros (:68-70) */
    if (T == 1) { out[0] = targets[0]; return; }                         /* :71-73 */
    int64_t* ext = (int64_t*)malloc(sizeof(int64_t) * (size_t)S);
    for (int i = 0; i < S; i++) ext[i] = (i & 1) ? targets[i / 2] : (int64_t)blank;
    double* alpha = (double*)malloc(sizeof(double) * (size_t)S * (size_t)T);
    int* path = (int*)calloc((size_t)S * (size_t)T, sizeof(int));        /* np.zeros_like: 0 outside the window */
    for (size_t q = 0; q < (size_t)S * (size_t)T; q++) alpha[q] = NEG_INF;
#define A(i, k) alpha[(size_t)(i) * T + (k)]
#define P(i, k) path[(size_t)(i) * T + (k)]
    A(0, 0) = lp[ext[0]];
    A(1, 0) = lp[ext[1]];
    for (int k = 1; k < T; k++) {
      const int start = imax(0, S - 2 * (T - k)), end = imin(k * 2 + 2, S);
      for (int i = start; i < end; i++) { A(i, k) = A(i, k - 1); P(i, k) = i; }
      for (int i = start; i < end; i++) {
        const int64_t cur = ext[i];
        if (i > 0) {
          if (A(i - 1, k - 1) > A(i, k)) { A(i, k) = A(i - 1, k - 1); P(i, k) = i - 1; }
          if (cur != blank && i - 2 > 0 && ext[i - 2] != cur && A(i - 2, k - 1) > A(i, k)) { A(i, k) = A(i - 2, k - 1); P(i, k) = i - 2; }
        }
        A(i, k) += lp[(size_t)k * V + cur];
      }
    }
    int i = S - 1;
    if (A(i - 1, T - 1) > A(i, T - 1)) i = i - 1;
    for (int k = T - 1; k >= 0; k--) { out[k] = ext[i]; i = P(i, k); }
    free(ext); free(alpha); free(path);
  } else {
    if (L == 0) return;
    if (T == 1) { out[0] = targets[0]; return; }                         /* :20-22 */
    double* alpha = (double*)malloc(sizeof(double) * (size_t)L * (size_t)T);
    int* path = (int*)calloc((size_t)L * (size_t)T, sizeof(int));
    for (size_t q = 0; q < (size_t)L * (size_t)T; q++) alpha[q] = NEG_INF;
    A(0, 0) = lp[targets[0]];
    for (int k = 1; k < T; k++) {
      const int start = imax(0, L - (T - k)), end = imin(k + 1, L);
      for (int i = start; i < end; i++) { A(i, k) = A(i, k - 1); P(i, k) = i; }
      for (int i = start; i < end; i++) {
        if (i > 0 && A(i - 1, k - 1) > A(i, k)) { A(i, k) = A(i - 1, k - 1); P(i, k) = i - 1; }
        A(i, k) += lp[(size_t)k * V + targets[i]];
      }
    }
    int i = L - 1;
    for (int k = T - 1; k >= 0; k--) { out[k] = targets[i]; i = P(i, k); }
    free(alpha); free(path);
#undef A
#undef P
  }
}

/* get_alignment_3d (alignment.py:109-138): out [B][T] int64, -100 past every utterance's frames. */
void ctc_oracle_align(const double* lp, int B, int T, int V, const int64_t* targets, int Lmax, const int64_t* in_len,
                      const int64_t* tgt_len, int blank, int is_ctc, int64_t* out) {
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < B; b++) {
    int64_t* o = out + (size_t)b * T;
    for (int k = 0; k < T; k++) o[k] = -100;
    align_utterance(lp + (size_t)b * T * V, V, targets + (size_t)b * Lmax, (int)in_len[b], (int)tgt_len[b], blank, is_ctc, o);
  }
}

/*
 * CTC without blank, one utterance (ctc_without_blank.py:13-88).  lp: [T][V] log-probabilities widened to double,
 * lp32: the same as float32 (the reference's np.exp(logits) runs in float32); grad: [T][V].  Returns -loss_forward.
 */
static double noblank_utterance(const double* lp, const float* lp32, int V, const int64_t* targets, int T, int L, int space,
                                double* grad) {
  int uas = 0, S;
  int64_t* ext;
  if (L == 0 || (L == 1 && targets[0] == space)) { S = 1; ext = (int64_t*)malloc(sizeof(int64_t)); ext[0] = space; }
  else if (space == -1) { S = L; ext = (int64_t*)malloc(sizeof(int64_t) * (size_t)S); memcpy(ext, targets, sizeof(int64_t) * (size_t)S); }
  else {
    uas = 1; S = L + 2; ext = (int64_t*)malloc(sizeof(int64_t) * (size_t)S);
    for (int j = 0; j < S; j++) ext[j] = space;
    memcpy(ext + 1, targets, sizeof(int64_t) * (size_t)L);
  }
  for (int j = 0; j < S; j++) if (ext[j] < 0) ext[j] += V;               /* numpy's negative index */
  double* alpha = (double*)malloc(sizeof(double) * (size_t)S * (size_t)T);
  double* beta = (double*)malloc(sizeof(double) * (size_t)S * (size_t)T);
  for (size_t q = 0; q < (size_t)S * (size_t)T; q++) { alpha[q] = NEG_INF; beta[q] = NEG_INF; }
#define AL(j, t) alpha[(size_t)(j) * T + (t)]
#define BE(j, t) beta[(size_t)(j) * T + (t)]
#define LP(t, v) lp[(size_t)(t) * V + (v)]
  if (T > 1 || S == 1) AL(0, 0) = LP(0, ext[0]);
  if (S > 1 && uas) AL(1, 0) = LP(0, ext[1]);
  for (int t = 1; t < T; t++) {
    const int start = uas ? imax(0, S - T + t - 1) : imax(0, S - T + t);
    const int end = uas ? imin(t + 2, S) : imin(t + 1, S);
    for (int j = start; j < end; j++) AL(j, t) = AL(j, t - 1);
    for (int j = start; j < end; j++) {
      if (j > 0) AL(j, t) = lse2(AL(j, t), AL(j - 1, t - 1));
      AL(j, t) += LP(t, ext[j]);
    }
  }
  const double loss_forward = (S > 1 && uas) ? lse2(AL(S - 1, T - 1), AL(S - 2, T - 1)) : AL(S - 1, T - 1);
  if (T > 1 || S == 1) BE(S - 1, T - 1) = 0.0;
  if (S > 1 && uas) BE(S - 2, T - 1) = 0.0;
  for (int t = T - 2; t >= 0; t--) {
    const int start = uas ? imax(0, S - T + t - 1) : imax(0, S - T + t);
    const int end = uas ? imin(t + 2, S) : imin(t + 1, S);
    for (int j = start; j < end; j++) {
      BE(j, t) = BE(j, t + 1) + LP(t + 1, ext[j]);
      if (j < S - 1) BE(j, t) = lse2(BE(j, t), BE(j + 1, t + 1) + LP(t + 1, ext[j + 1]));
    }
  }
  double* psum = (double*)malloc(sizeof(double) * (size_t)T * (size_t)V);
  for (size_t q = 0; q < (size_t)T * (size_t)V; q++) psum[q] = NEG_INF;
  for (int i = 0; i < S; i++)
    for (int t = 0; t < T; t++) psum[(size_t)t * V + ext[i]] = lse2(psum[(size_t)t * V + ext[i]], AL(i, t) + BE(i, t));
  for (size_t q = 0; q < (size_t)T * (size_t)V; q++) grad[q] = (double)expf(lp32[q]) - exp(psum[q] - loss_forward);
#undef AL
#undef BE
#undef LP
  free(ext); free(alpha); free(beta); free(psum);
  return -loss_forward;
}

/* _ctc_without_blank_3d_loss (:91-117): grads zero past every utterance's frames. */
void ctc_oracle_noblank(const float* lp32, int B, int T, int V, const int64_t* targets, int Lmax, const int64_t* in_len,
                        const int64_t* tgt_len, int space, double* losses, double* grads) {
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < B; b++) {
    const int Ti = (int)in_len[b];
    const float* x32 = lp32 + (size_t)b * T * V;
    double* x = (double*)malloc(sizeof(double) * (size_t)Ti * (size_t)V);
    for (size_t q = 0; q < (size_t)Ti * (size_t)V; q++) x[q] = (double)x32[q];
    double* g = grads + (size_t)b * T * V;
    for (size_t q = 0; q < (size_t)T * (size_t)V; q++) g[q] = 0.0;
    losses[b] = noblank_utterance(x, x32, V, targets + (size_t)b * Lmax, Ti, (int)tgt_len[b], space, g);
    free(x);
  }
}

/*
 * LM-free prefix beam search, one utterance: src/decoders/ctc_decoder.cpp:353-441 (decode_sentence) with
 * get_next_prefix (:241-309, the lm_model == nullptr branch), Prefix::next_step / get_prev_full_prob (:331-340),
 * get_prev_full_prob_with_lmwt (:311-315) and Prefix::get_sentence (:225-239).
 *
 * The reference keeps prefixes as shared_ptr objects; a child is found through a weak_ptr in its parent's
 * `next_data`.  What that ownership model DOES is restated with explicit reference counts (BNode.refs: one for
 * membership of the beam + one per living child, exactly the shared_ptr holders `prefixes` and `Prefix::parent`):
 *   - a child that is still in the beam is found and accumulates (is_new == false, :244-246),
 *   - a child that was pruned but is kept alive by a descendant in the beam is ALSO found, is not in the list,
 *     and therefore swallows the extension: that prefix cannot re-enter the beam while the descendant lives
 *     (a property of the reference, reproduced on purpose),
 *   - a child whose last owner went away has expired: a fresh prefix is made (:248-308).
 * The V-1 fresh prefixes every beam member makes per frame (:377-378) are records in a flat array here; only
 * those that survive the prune become nodes (the others are created and destroyed within the frame in the
 * reference, unobserved).  Scores: get_prev_full_prob() + lm_score*lmwt - num_words*wip + num_oov_words*oov_penalty
 * with lm_score = 0, lmwt = 0 (ctc_decoder.cpp:76), num_oov_words = 0.
 * Prune (:399-410) keeps the beam_width best; std::nth_element / std::sort leave the choice among EQUAL scores to
 * the library -- here ties go to the lower position (beam members in beam order, then fresh prefixes by
 * (member, symbol)), and the surviving list keeps that order.
 * Result (:430-435): the symbols of the best prefix; the EMPTY prefix yields the single symbol -1 with length 1
 * (get_sentence pushes last_char = -1 of the root), as the compiled reference returns.
 */
typedef struct BNode {
  struct BNode *parent, *first_child, *next_sib;
  int chr, refs, slot;
} BNode;

static void bnode_unref(BNode* n) {
  while (n && --n->refs == 0) {
    BNode* p = n->parent;
    if (p) {  /* weak_ptr in the parent's next_data expires */
      BNode** q = &p->first_child;
      while (*q != n) q = &(*q)->next_sib;
      *q = n->next_sib;
    }
    free(n);
    n = p;
  }
}

typedef struct { double score; int idx; } BKey;
static int bkey_cmp(const void* a, const void* b) {
  const BKey *x = (const BKey*)a, *y = (const BKey*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

static double beam_score(double full, int num_words, double wip, double oov_penalty) {
  return full + 0.0 * 0.0 - num_words * wip + 0 * oov_penalty;
}

static int beam_utterance(const double* lp, int V, int T, int blank, int beam_width, int space_id, double wip,
                          double oov_penalty, int64_t* out, int64_t* ties) {
  const size_t cap = (size_t)beam_width + 1;
  int64_t n_ties = 0;   /* prunes (and the final pick) where EQUAL scores straddle the cut: the reference's choice is libstdc++'s */
  BNode** node = (BNode**)malloc(sizeof(BNode*) * cap);
  BNode** node2 = (BNode**)malloc(sizeof(BNode*) * cap);
  double *pb = (double*)malloc(sizeof(double) * cap), *pnb = (double*)malloc(sizeof(double) * cap);
  double *pb2 = (double*)malloc(sizeof(double) * cap), *pnb2 = (double*)malloc(sizeof(double) * cap);
  double *npb = (double*)malloc(sizeof(double) * cap), *npnb = (double*)malloc(sizeof(double) * cap), *full = (double*)malloc(sizeof(double) * cap);
  int *nw = (int*)malloc(sizeof(int) * cap), *nw2 = (int*)malloc(sizeof(int) * cap);
  /* fresh prefixes of this frame, indexed member * V + symbol */
  const size_t fcap = cap * (size_t)V;
  double* f_pnb = (double*)malloc(sizeof(double) * fcap);
  int* f_nw = (int*)malloc(sizeof(int) * fcap);
  char* f_made = (char*)malloc(fcap);
  char* alive = (char*)malloc(cap + fcap);
  BKey* keys = (BKey*)malloc(sizeof(BKey) * (fcap + cap));

  BNode* root = (BNode*)calloc(1, sizeof(BNode));   /* get_initial_prefix (:201-209) */
  root->chr = -1; root->refs = 1; root->slot = 0;
  int W = 1;
  node[0] = root; pb[0] = 0.0; pnb[0] = NEG_INF; nw[0] = 0;

  for (int t = 0; t < T; t++) {
    const double* row = lp + (size_t)t * V;
    for (int s = 0; s < W; s++) { full[s] = lse2(pnb[s], pb[s]); npb[s] = NEG_INF; npnb[s] = NEG_INF; }
    memset(f_made, 0, (size_t)W * V);
    for (int c = 0; c < V; c++) {          /* for every character (:368) */
      const double cur = row[c];
      for (int s = 0; s < W; s++) {        /* for every prefix (:371) */
        if (c == blank) { npb[s] = lse2(npb[s], cur + full[s]); continue; }
        BNode* ch = node[s]->first_child;  /* get_next_prefix (:244-246): a living child is found */
        while (ch && ch->chr != c) ch = ch->next_sib;
        double sink = NEG_INF, *target;
        if (ch) target = ch->slot >= 0 ? &npnb[ch->slot] : &sink;   /* in the beam, or kept alive by a descendant */
        else {
          const size_t f = (size_t)s * V + c;
          f_made[f] = 1; f_pnb[f] = NEG_INF;
          f_nw[f] = nw[s] + ((c != space_id && (nw[s] == 0 || node[s]->chr == space_id)) ? 1 : 0);   /* :252-257 */
          target = &f_pnb[f];
        }
        if (c == node[s]->chr) {   /* repeated character (:381-385) */
          *target = lse2(*target, cur + pb[s]);
          npnb[s] = lse2(npnb[s], cur + pnb[s]);
        } else {
          *target = lse2(*target, cur + full[s]);
        }
      }
    }
    /* next_step for everybody (:397), then the prune (:399-410) */
    int total = 0;
    for (int s = 0; s < W; s++) { keys[total].score = beam_score(lse2(npnb[s], npb[s]), nw[s], wip, oov_penalty); keys[total++].idx = s; }
    for (size_t f = 0; f < (size_t)W * V; f++)
      if (f_made[f]) { keys[total].score = beam_score(lse2(f_pnb[f], NEG_INF), f_nw[f], wip, oov_penalty); keys[total++].idx = W + (int)f; }
    int keep = total;
    if (total > beam_width) {
      qsort(keys, (size_t)total, sizeof(BKey), bkey_cmp);
      keep = beam_width;
      if (keys[keep - 1].score == keys[keep].score) n_ties++;
    }
    memset(alive, 0, (size_t)W + (size_t)W * V);
    for (int i = 0; i < keep; i++) alive[keys[i].idx] = 1;
    int W2 = 0;   /* the surviving list, in position order */
    for (int s = 0; s < W; s++)
      if (alive[s]) { node2[W2] = node[s]; pb2[W2] = npb[s]; pnb2[W2] = npnb[s]; nw2[W2] = nw[s]; W2++; }
    for (int s = 0; s < W; s++)
      for (int c = 0; c < V; c++) {
        const size_t f = (size_t)s * V + c;
        if (!f_made[f] || !alive[W + f]) continue;
        BNode* n = (BNode*)calloc(1, sizeof(BNode));
        n->chr = c; n->refs = 1; n->parent = node[s]; node[s]->refs++;
        n->next_sib = node[s]->first_child; node[s]->first_child = n;
        node2[W2] = n; pb2[W2] = NEG_INF; pnb2[W2] = f_pnb[f]; nw2[W2] = f_nw[f]; W2++;
      }
    for (int s = 0; s < W; s++) if (!alive[s]) { node[s]->slot = -1; bnode_unref(node[s]); }
    W = W2;
    for (int s = 0; s < W; s++) { node[s] = node2[s]; node[s]->slot = s; pb[s] = pb2[s]; pnb[s] = pnb2[s]; nw[s] = nw2[s]; }
  }

  int best = 0;   /* std::sort (:413-419) + prefixes[0] */
  double best_score = beam_score(lse2(pnb[0], pb[0]), nw[0], wip, oov_penalty);
  for (int s = 1; s < W; s++) {
    const double sc = beam_score(lse2(pnb[s], pb[s]), nw[s], wip, oov_penalty);
    if (sc > best_score) { best_score = sc; best = s; }
  }
  for (int s = 0; s < W; s++)
    if (s != best && beam_score(lse2(pnb[s], pb[s]), nw[s], wip, oov_penalty) == best_score) { n_ties++; break; }
  if (ties) *ties = n_ties;
  /* get_sentence (:225-239): last_char of the prefix, then of every ancestor except the root */
  int n = 0;
  for (BNode* p = node[best]; p; p = p->parent) if (p == node[best] || p->parent) n++;
  int k = n;
  for (BNode* p = node[best]; p; p = p->parent) if (p == node[best] || p->parent) out[--k] = p->chr;
  for (int s = 0; s < W; s++) bnode_unref(node[s]);
  free(node); free(node2); free(pb); free(pnb); free(pb2); free(pnb2); free(npb); free(npnb); free(full); free(nw); free(nw2);
  free(f_pnb); free(f_nw); free(f_made); free(alive); free(keys);
  return n;
}

/* CTCDecoder::decode (:153-198): out [B][Tw] zero-filled, Tw = max(T, 1); out_len [B].  lp are log-probabilities
 * (the Python wrapper has applied log_softmax: decoders/ctc_decoder.py:95-97).  ties [B] (optional): how many prunes of
 * the utterance had equal scores on both sides of the cut (0: the result does not depend on the sort library). */
void ctc_oracle_beam(const double* lp, int B, int T, int V, const int64_t* in_len, int blank, int beam_width, int space_id,
                     double wip, double oov_penalty, int64_t* out, int64_t* out_len, int64_t* ties) {
  const int Tw = T > 1 ? T : 1;
  memset(out, 0, sizeof(int64_t) * (size_t)B * Tw);
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < B; b++) {
    int Ti = (int)in_len[b];
    if (Ti > T) Ti = T;
    if (Ti < 0) Ti = 0;
    out_len[b] = beam_utterance(lp + (size_t)b * T * V, V, Ti, blank, beam_width, space_id, wip, oov_penalty, out + (size_t)b * Tw,
                                ties ? ties + b : NULL);
  }
}
