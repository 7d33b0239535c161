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
