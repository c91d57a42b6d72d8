/* Oracle (test infrastructure): CTC loss + gradient w.r.t. raw activations, plain C.
 *
 * The reference computes this in the third-party warp-ctc (SeanNaren/warp-ctc pytorch_binding, un-pinned, not
 * vendored under /root/reference; call sites src/train_cnn_lstm.py:12,52,138,358).  This file restates the
 * published algorithm (Graves et al. 2006, as implemented by warp-ctc's CPU path): softmax inside, blank = 0,
 * alpha/beta over the blank-extended labelling l' (S = 2L+1), log space, per-utterance cost = -ln p(l|x),
 * grad[t,k] = y[t,k] - (1/p) * sum_{s: l'_s = k} alpha_t(s) beta_t(s) / y[t,l'_s], zero for t >= act_len,
 * infeasible utterance => cost 0 / grad 0.
 * PARITY PIN: warp-ctc's own known-answer test vector (T=2, A=5, labels {1,2}: cost 2.4628584384918, see
 * tests/test_oracle_ctc.py), brute-force enumeration of all paths for T <= 6, and torch.nn.functional.ctc_loss on
 * CPU.  warp-ctc itself is absent, so parity with the reference binary is otherwise UNPINNED.
 *
 * Compile: gcc -O2 -shared -fPIC -o oracle/_build/libctc_ref.so oracle/ctc_ref.c -lm   (oracle/build_oracle.py)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#endif

static REAL lse2(REAL a, REAL b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  REAL m = a > b ? a : b;
  return m + log(exp(a - m) + exp(b - m));
}

/* acts: [T,B,A] row-major (float32 input, accumulated in REAL); grads may be NULL.
 * labels: concatenated; returns 0 on success. */
int ctc_ref(const float* acts, float* grads, const int* labels, const int* label_lens, const int* act_lens, int T,
            int B, int A, double* costs) {
  int off = 0;
  if (grads) memset(grads, 0, sizeof(float) * (size_t)T * B * A);
  for (int b = 0; b < B; ++b) {
    const int L = label_lens[b];
    const int S = 2 * L + 1;
    int Tb = act_lens[b];
    if (Tb > T) Tb = T;
    const int* lab = labels + off;
    off += L;
    costs[b] = 0.0;
    if (Tb <= 0) continue;
    REAL* lp = (REAL*)malloc(sizeof(REAL) * (size_t)Tb * A); /* log-softmax */
    REAL* al = (REAL*)malloc(sizeof(REAL) * (size_t)Tb * S);
    REAL* be = (REAL*)malloc(sizeof(REAL) * (size_t)Tb * S);
    for (int t = 0; t < Tb; ++t) {
      const float* row = acts + ((size_t)t * B + b) * A;
      REAL m = -INFINITY;
      for (int a = 0; a < A; ++a)
        if (row[a] > m) m = row[a];
      REAL s = 0;
      for (int a = 0; a < A; ++a) s += exp((REAL)row[a] - m);
      const REAL lse = m + log(s);
      for (int a = 0; a < A; ++a) lp[(size_t)t * A + a] = (REAL)row[a] - lse;
    }
#define SYM(s) (((s)&1) ? lab[(s) >> 1] : 0)
    for (int s = 0; s < S; ++s) al[s] = be[(size_t)(Tb - 1) * S + s] = -INFINITY;
    al[0] = lp[0];
    if (S > 1) al[1] = lp[SYM(1)];
    for (int t = 1; t < Tb; ++t)
      for (int s = 0; s < S; ++s) {
        REAL v = al[(size_t)(t - 1) * S + s];
        if (s >= 1) v = lse2(v, al[(size_t)(t - 1) * S + s - 1]);
        if ((s & 1) && s >= 3 && SYM(s) != SYM(s - 2)) v = lse2(v, al[(size_t)(t - 1) * S + s - 2]);
        al[(size_t)t * S + s] = (v == -INFINITY) ? v : v + lp[(size_t)t * A + SYM(s)];
      }
    be[(size_t)(Tb - 1) * S + S - 1] = lp[(size_t)(Tb - 1) * A];
    if (S > 1) be[(size_t)(Tb - 1) * S + S - 2] = lp[(size_t)(Tb - 1) * A + SYM(S - 2)];
    for (int t = Tb - 2; t >= 0; --t)
      for (int s = 0; s < S; ++s) {
        REAL v = be[(size_t)(t + 1) * S + s];
        if (s + 1 < S) v = lse2(v, be[(size_t)(t + 1) * S + s + 1]);
        if ((s & 1) && s + 2 < S && SYM(s) != SYM(s + 2)) v = lse2(v, be[(size_t)(t + 1) * S + s + 2]);
        be[(size_t)t * S + s] = (v == -INFINITY) ? v : v + lp[(size_t)t * A + SYM(s)];
      }
    REAL ll = al[(size_t)(Tb - 1) * S + S - 1];
    if (S > 1) ll = lse2(ll, al[(size_t)(Tb - 1) * S + S - 2]);
    if (ll != -INFINITY) {
      costs[b] = (double)(-ll);
      if (grads) {
        REAL* occ = (REAL*)malloc(sizeof(REAL) * A);
        for (int t = 0; t < Tb; ++t) {
          for (int a = 0; a < A; ++a) occ[a] = 0;
          for (int s = 0; s < S; ++s) {
            const REAL ab = al[(size_t)t * S + s] + be[(size_t)t * S + s];
            if (ab == -INFINITY) continue;
            occ[SYM(s)] += exp(ab - lp[(size_t)t * A + SYM(s)] - ll);
          }
          float* g = grads + ((size_t)t * B + b) * A;
          for (int a = 0; a < A; ++a) g[a] = (float)(exp(lp[(size_t)t * A + a]) - occ[a]);
        }
        free(occ);
      }
    }
#undef SYM
    free(lp);
    free(al);
    free(be);
  }
  return 0;
}
