/* TEST INFRASTRUCTURE ONLY (oracle/): the value of KADAL's concentrated ln-likelihood formula
 * (kadal/surrogate_models/supports/likelihood_func.py:168-213, ordinary Kriging, trainvar=False)
 * evaluated in IEEE binary128 (__float128, libquadmath) from the SAME float64 Psi the reference
 * builds.  It is the yard-stick of tests/golden/c2_truth_full.npz: how far the reference's own
 * float64 result and the CUDA result are from the exact value of the formula.
 *
 *   Psi += eye * eps                    (:171, the addition itself in float64 as the reference does)
 *   U = chol(Psi)^T                     (:175-176)
 *   LnDetPsi = 2 sum ln |U_ii|          (:180)
 *   mu  = (1' Psi^-1 y) / (1' Psi^-1 1) (:183-193, F = ones)
 *   s2  = (y - mu)' Psi^-1 (y - mu) / n (:202)
 *   NLL = (n/2) ln s2 + LnDetPsi / 2    (:203-204)
 *
 * Build: oracle/truth/Makefile (gcc -O2 -shared -fPIC nll_quad.c -lquadmath).
 */
#include <quadmath.h>
#include <stdlib.h>

typedef __float128 q128;

/* A: n x n row-major float64 (Psi + eps I already formed in float64); y: n.  Returns the NLL rounded to double,
 * or NaN when a pivot is not positive.  work: n*(n+1)/2 + 2n q128 values allocated here. */
double kb_truth_nll_quad(const double* A, const double* y, int n) {
  q128* L = (q128*)malloc(((size_t)n * (n + 1) / 2 + 2 * (size_t)n) * sizeof(q128));
  if (!L) return 0.0 / 0.0;
  q128* z1 = L + (size_t)n * (n + 1) / 2;
  q128* zy = z1 + n;
  q128 logdet = 0;
  for (int i = 0; i < n; ++i) {
    q128* Li = L + (size_t)i * (i + 1) / 2;
    for (int j = 0; j <= i; ++j) {
      const q128* Lj = L + (size_t)j * (j + 1) / 2;
      q128 s = (q128)A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= Li[k] * Lj[k];
      if (j < i) {
        Li[j] = s / Lj[j];
      } else {
        if (!(s > 0)) { free(L); return 0.0 / 0.0; }
        Li[i] = sqrtq(s);
        logdet += logq(Li[i]);
      }
    }
    /* forward substitution of [1, y] rides along */
    q128 a = 1, b = (q128)y[i];
    for (int k = 0; k < i; ++k) { a -= Li[k] * z1[k]; b -= Li[k] * zy[k]; }
    z1[i] = a / Li[i];
    zy[i] = b / Li[i];
  }
  q128 s11 = 0, s1y = 0;
  for (int i = 0; i < n; ++i) { s11 += z1[i] * z1[i]; s1y += z1[i] * zy[i]; }
  const q128 mu = s1y / s11;
  q128 rr = 0;
  for (int i = 0; i < n; ++i) { const q128 r = zy[i] - mu * z1[i]; rr += r * r; }
  const q128 s2 = rr / n;
  const q128 nll = ((q128)n / 2) * logq(s2) + logdet;   /* LnDetPsi / 2 = sum ln L_ii */
  free(L);
  return (double)nll;
}

/* The same with a general regression matrix F (n x p, row-major): TrendOrder >= 1, likelihood_func.py:183-193.
 *   beta = (F' Psi^-1 F)^-1 F' Psi^-1 y ;  r = y - F beta ;  s2 = r' Psi^-1 r / n
 * In factor form: B = L^-1 F, a = L^-1 y, G = B'B (p x p, Cholesky in binary128), beta = G^-1 B'a, rr = |a - B beta|^2.
 * beta_out (p doubles) may be NULL. */
double kb_truth_nll_quad_F(const double* A, const double* y, const double* F, int n, int p, double* beta_out) {
  const size_t tri = (size_t)n * (n + 1) / 2;
  q128* L = (q128*)malloc((tri + (size_t)n * (p + 1) + (size_t)p * p + 2 * (size_t)p) * sizeof(q128));
  if (!L) return 0.0 / 0.0;
  q128* Z = L + tri;               /* [n][p + 1]: columns 0..p-1 = B, column p = a */
  q128* G = Z + (size_t)n * (p + 1);
  q128* g = G + (size_t)p * p;
  q128* beta = g + p;
  q128 logdet = 0;
  for (int i = 0; i < n; ++i) {
    q128* Li = L + (size_t)i * (i + 1) / 2;
    for (int j = 0; j <= i; ++j) {
      const q128* Lj = L + (size_t)j * (j + 1) / 2;
      q128 s = (q128)A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= Li[k] * Lj[k];
      if (j < i) Li[j] = s / Lj[j];
      else {
        if (!(s > 0)) { free(L); return 0.0 / 0.0; }
        Li[i] = sqrtq(s);
        logdet += logq(Li[i]);
      }
    }
    for (int c = 0; c <= p; ++c) {
      q128 v = c < p ? (q128)F[(size_t)i * p + c] : (q128)y[i];
      for (int k = 0; k < i; ++k) v -= Li[k] * Z[(size_t)k * (p + 1) + c];
      Z[(size_t)i * (p + 1) + c] = v / Li[i];
    }
  }
  for (int a = 0; a < p; ++a) {
    for (int b = 0; b < p; ++b) {
      q128 s = 0;
      for (int i = 0; i < n; ++i) s += Z[(size_t)i * (p + 1) + a] * Z[(size_t)i * (p + 1) + b];
      G[a * p + b] = s;
    }
    q128 s = 0;
    for (int i = 0; i < n; ++i) s += Z[(size_t)i * (p + 1) + a] * Z[(size_t)i * (p + 1) + p];
    g[a] = s;
  }
  /* G = C C' ; solve C C' beta = g */
  for (int i = 0; i < p; ++i)
    for (int j = 0; j <= i; ++j) {
      q128 s = G[i * p + j];
      for (int k = 0; k < j; ++k) s -= G[i * p + k] * G[j * p + k];
      if (j < i) G[i * p + j] = s / G[j * p + j];
      else {
        if (!(s > 0)) { free(L); return 0.0 / 0.0; }
        G[i * p + i] = sqrtq(s);
      }
    }
  for (int i = 0; i < p; ++i) {
    q128 s = g[i];
    for (int k = 0; k < i; ++k) s -= G[i * p + k] * beta[k];
    beta[i] = s / G[i * p + i];
  }
  for (int i = p - 1; i >= 0; --i) {
    q128 s = beta[i];
    for (int k = i + 1; k < p; ++k) s -= G[k * p + i] * beta[k];
    beta[i] = s / G[i * p + i];
  }
  q128 rr = 0;
  for (int i = 0; i < n; ++i) {
    q128 r = Z[(size_t)i * (p + 1) + p];
    for (int c = 0; c < p; ++c) r -= Z[(size_t)i * (p + 1) + c] * beta[c];
    rr += r * r;
  }
  if (beta_out) for (int c = 0; c < p; ++c) beta_out[c] = (double)beta[c];
  const q128 nll = ((q128)n / 2) * logq(rr / n) + logdet;
  free(L);
  return (double)nll;
}
