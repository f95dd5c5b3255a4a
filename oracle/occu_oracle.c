/*
 * CPU restatement (plain C + OpenMP) of the occupancy log-density + gradient.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Never part of the product path.
 * PARITY STATUS: unpinned against numpyro (see oracle/occupancy.py header); this file is pinned
 * to oracle/occupancy.py (tests/test_oracle_c.py) which is pinned to the golden fixtures.
 *
 * Five entry points, one per model (occu below; occu_cop, occu_rn, nmixture, occu_cs further down, each
 * with its own header citing the reference lines it follows).
 *
 * Restates, for the `occu` model without false positives (the BASELINE.json metric config):
 *   biolith/models/occu.py:136-142   NaN mask + nan_to_num            (done on the fly per visit)
 *   biolith/regression/linear.py:59-66  eta = b0 + X.b, nu = a0 + W.a
 *   biolith/models/occu.py:207-242   psi, p, masked Bernoulli log-prob, enumeration over z
 *   numpyro clamp_probs / BernoulliProbs.log_prob semantics (tiny / 1-eps clamps, zero slope outside)
 *   reverse-mode gradient (hand-derived; SURVEY.md section 8a closed forms)
 * Arithmetic type = `real` (float or double, chosen per call), reductions in double.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXK 16

#define DEFINE_ORACLE(NAME, real, EXP, LOG1P, FABS, LOG_TINY, LOG_EPS, LOG1M_EPS, NEG_TINY, RMAX)                   \
  static inline real NAME##_n2n(real v) { return isnan(v) ? (real)0 : (isinf(v) ? (v > 0 ? RMAX : -RMAX) : v); }   \
  static inline void NAME##_lsp(real x, real* p, real* q, real* lp, real* l1, int* inr) {                          \
    real t = EXP(-FABS(x)), l = LOG1P(t), inv = (real)1 / ((real)1 + t), ti = t * inv;                              \
    *p = x >= 0 ? inv : ti;                                                                                         \
    *q = x >= 0 ? ti : inv;                                                                                         \
    real a = (x < 0 ? x : 0) - l, b = -(x > 0 ? x : 0) - l;                                                         \
    int lo = a <= (real)LOG_TINY, hi = b <= (real)LOG_EPS;                                                          \
    *inr = !(lo || hi);                                                                                             \
    *lp = lo ? (real)LOG_TINY : (hi ? (real)LOG1M_EPS : a);                                                         \
    *l1 = lo ? (real)NEG_TINY : (hi ? (real)LOG_EPS : b);                                                           \
  }                                                                                                                 \
  static void NAME(long S, int P, int J, int Ks, int Ko, const real* y, const real* X, const real* W,              \
                   const double* theta, int C, int prior, double* logp, double* grad) {                             \
    const int D = Ks + Ko + 2, NQ = D + 1;                                                                          \
    const long U = S * (long)P;                                                                                     \
    for (int c = 0; c < C; ++c) {                                                                                   \
      real b[MAXK + 1], a[MAXK + 1];                                                                                \
      for (int k = 0; k <= Ks; ++k) b[k] = (real)theta[(size_t)c * D + k];                                          \
      for (int k = 0; k <= Ko; ++k) a[k] = (real)theta[(size_t)c * D + Ks + 1 + k];                                 \
      double acc[2 * MAXK + 3];                                                                                     \
      memset(acc, 0, sizeof(acc));                                                                                  \
      _Pragma("omp parallel")                                                                                       \
      {                                                                                                             \
        double loc[2 * MAXK + 3];                                                                                   \
        memset(loc, 0, sizeof(loc));                                                                                \
        _Pragma("omp for schedule(static)")                                                                         \
        for (long u = 0; u < U; ++u) {                                                                              \
          const long s = u / P;                                                                                     \
          int site_nan = 0;                                                                                         \
          real x[MAXK], eta = b[0];                                                                                 \
          for (int k = 0; k < Ks; ++k) {                                                                            \
            real v = X[s * Ks + k];                                                                                 \
            site_nan |= isnan(v);                                                                                   \
            x[k] = NAME##_n2n(v);                                                                                   \
            eta += x[k] * b[k + 1];                                                                                 \
          }                                                                                                         \
          real L1 = 0, ga[MAXK + 1];                                                                                \
          for (int k = 0; k <= Ko; ++k) ga[k] = 0;                                                                  \
          int n1 = 0, n0 = 0;                                                                                       \
          for (int j = 0; j < J; ++j) {                                                                             \
            const real* w = W + ((size_t)u * J + j) * Ko;                                                           \
            real wv[MAXK], nu = a[0];                                                                               \
            int cov_nan = site_nan;                                                                                 \
            for (int k = 0; k < Ko; ++k) {                                                                          \
              cov_nan |= isnan(w[k]);                                                                               \
              wv[k] = NAME##_n2n(w[k]);                                                                             \
              nu += wv[k] * a[k + 1];                                                                               \
            }                                                                                                       \
            const real yv = y[(size_t)u * J + j];                                                                   \
            if (cov_nan || !isfinite(yv)) continue; /* mask_missing_obs */                                          \
            real p, q, lp, l1;                                                                                      \
            int inr;                                                                                                \
            NAME##_lsp(nu, &p, &q, &lp, &l1, &inr);                                                                 \
            const int yb = yv != 0;                                                                                 \
            n1 += yb;                                                                                               \
            n0 += !yb;                                                                                              \
            L1 += yb ? lp : l1;                                                                                     \
            const real g = inr ? (yb ? q : -p) : 0;                                                                 \
            ga[0] += g;                                                                                             \
            for (int k = 0; k < Ko; ++k) ga[k + 1] += g * wv[k];                                                    \
          }                                                                                                         \
          const real L0 = (real)n1 * (real)LOG_TINY + (real)n0 * (real)NEG_TINY;                                    \
          real psi, qpsi, lpsi, l1psi;                                                                              \
          int in_psi;                                                                                               \
          NAME##_lsp(eta, &psi, &qpsi, &lpsi, &l1psi, &in_psi);                                                     \
          const real av = lpsi + L1, bv = l1psi + L0, dd = av - bv;                                                 \
          const real td = EXP(-FABS(dd)), inv = (real)1 / ((real)1 + td);                                           \
          const real r = dd >= 0 ? inv : td * inv;                                                                  \
          const real ell = (av > bv ? av : bv) + LOG1P(td);                                                         \
          const real geta = in_psi ? r - psi : 0;                                                                   \
          loc[0] += ell;                                                                                            \
          loc[1] += geta;                                                                                           \
          for (int k = 0; k < Ks; ++k) loc[2 + k] += geta * x[k];                                                   \
          for (int k = 0; k <= Ko; ++k) loc[2 + Ks + k] += r * ga[k];                                               \
        }                                                                                                           \
        _Pragma("omp critical")                                                                                     \
        for (int i = 0; i < NQ; ++i) acc[i] += loc[i];                                                              \
      }                                                                                                             \
      double lp = acc[0];                                                                                           \
      for (int i = 0; i < D; ++i) {                                                                                 \
        double g = acc[1 + i], t = theta[(size_t)c * D + i];                                                        \
        if (prior) {                                                                                                \
          lp += -0.5 * t * t - 0.91893853320467274178;                                                              \
          g -= t;                                                                                                   \
        }                                                                                                           \
        grad[(size_t)c * D + i] = g;                                                                                \
      }                                                                                                             \
      logp[c] = lp;                                                                                                 \
    }                                                                                                               \
  }

DEFINE_ORACLE(occu_f32, float, expf, log1pf, fabsf, -87.33654475f, -15.9423851f, -1.19209297e-07f, -FLT_MIN, FLT_MAX)
DEFINE_ORACLE(occu_f64, double, exp, log1p, fabs, -708.3964185322641, -36.04365338911715, -2.2204460492503136e-16,
              -DBL_MIN, DBL_MAX)

/* dtype: 0 = float arithmetic on float arrays, 1 = double on double arrays.  theta/logp/grad double.
 * Arrays in the reference layout: y (1,S,P,J), X (S,Ks), W (S,P,J,Ko), C-contiguous. */
int oracle_occu_logp_grad(int dtype, long S, int P, int J, int Ks, int Ko, const void* y, const void* X,
                          const void* W, const double* theta, int C, int prior, int nthreads, double* logp,
                          double* grad) {
  if (Ks > MAXK || Ko > MAXK) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  if (dtype == 0) occu_f32(S, P, J, Ks, Ko, (const float*)y, (const float*)X, (const float*)W, theta, C, prior, logp, grad);
  else occu_f64(S, P, J, Ks, Ko, (const double*)y, (const double*)X, (const double*)W, theta, C, prior, logp, grad);
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * occu_cop (count detections, BASELINE config 4) -- restates
 *   biolith/models/occu_cop.py:151-157   NaN mask + nan_to_num
 *   biolith/models/occu_cop.py:160-171   rate_fp_constant / rate_fp_unoccupied ~ Exponential(1), sampled as
 *                                        x = log(rate) (ExpTransform, log|J| = x)
 *   biolith/models/occu_cop.py:222-255   psi, mu = exp(nu), rate = T (z mu + (1 - z) u + c), masked Poisson
 *                                        log-prob (xlogy(y, rate) - lgamma(y + 1) - rate), enumeration over z
 * in double arithmetic; `f32_clamps` selects numpyro's clamp_probs constants of the dtype the reference
 * would run in (they only touch psi here).  Closed form and gradient: oracle/occupancy.py:occu_cop_logp_grad.
 * theta = [beta | alpha | log c (if fpc) | log u (if fpu)].  Arrays: y (1,S,P,J), X (S,Ks), W (S,P,J,Ko),
 * T (S,P,J) or NULL (ones), all double, C-contiguous.
 * ---------------------------------------------------------------------------------------------- */
static inline double n2n_clamped(double v, double rmax) { return isnan(v) ? 0.0 : (isinf(v) ? (v > 0 ? rmax : -rmax) : v); }

int oracle_occu_cop_logp_grad(int f32_clamps, long S, int P, int J, int Ks, int Ko, const double* y, const double* X,
                              const double* W, const double* T, const double* theta, int C, int fpc, int fpu,
                              int prior, int nthreads, double* logp, double* grad) {
  if (Ks > MAXK || Ko > MAXK) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  /* an fp32 run holds the logs of the clamp constants in fp32 (as the numpy oracle and the kernels do) */
  const double log_tiny = f32_clamps ? (double)logf(FLT_MIN) : log(DBL_MIN);
  const double log_eps = f32_clamps ? (double)logf(FLT_EPSILON) : log(DBL_EPSILON);
  const double log1m_eps = f32_clamps ? (double)log1pf(-FLT_EPSILON) : log1p(-DBL_EPSILON);
  const double neg_tiny = f32_clamps ? (double)log1pf(-FLT_MIN) : log1p(-DBL_MIN);
  const double rmax = f32_clamps ? (double)FLT_MAX : DBL_MAX;
  const int D = Ks + Ko + 2 + (fpc != 0) + (fpu != 0);
  const long U = S * (long)P;
  for (int ci = 0; ci < C; ++ci) {
    const double* th = theta + (size_t)ci * D;
    const double* b = th;
    const double* a = th + Ks + 1;
    int ie = Ks + Ko + 2;
    const double xc = fpc ? th[ie++] : 0.0, xu = fpu ? th[ie++] : 0.0;
    const double c = fpc ? exp(xc) : 0.0, u = fpu ? exp(xu) : 0.0;
    const double rho0 = u + c;
    double acc[2 * MAXK + 5];
    memset(acc, 0, sizeof(acc));
    const int NQ = D + 1;
#pragma omp parallel
    {
      double loc[2 * MAXK + 5];
      memset(loc, 0, sizeof(loc));
#pragma omp for schedule(static)
      for (long un = 0; un < U; ++un) {
        const long s = un / P;
        int site_nan = 0;
        double x[MAXK], eta = b[0];
        for (int k = 0; k < Ks; ++k) {
          const double v = X[s * Ks + k];
          site_nan |= isnan(v);
          x[k] = n2n_clamped(v, rmax);
          eta += x[k] * b[k + 1];
        }
        double t1 = 0, t0 = 0, s1 = 0, s0 = 0, ga[MAXK + 1];
        int b_is_neginf = 0;
        for (int k = 0; k <= Ko; ++k) ga[k] = 0;
        for (int j = 0; j < J; ++j) {
          const double* w = W + ((size_t)un * J + j) * Ko;
          double wv[MAXK], nu = a[0];
          int cov_nan = site_nan;
          for (int k = 0; k < Ko; ++k) {
            cov_nan |= isnan(w[k]);
            wv[k] = n2n_clamped(w[k], rmax);
            nu += wv[k] * a[k + 1];
          }
          const double yv = y[(size_t)un * J + j];
          if (cov_nan || !isfinite(yv)) continue; /* mask_missing_obs */
          const double tv = T ? T[(size_t)un * J + j] : 1.0;
          const double mu = exp(nu), rho1 = mu + c, lg = lgamma(yv + 1.0);
          t1 += (yv > 0 ? yv * log(tv * rho1) : 0.0) - lg - tv * rho1;
          if (yv > 0 && !(tv * rho0 > 0)) b_is_neginf = 1; /* xlogy(y, 0) = -inf */
          else t0 += (yv > 0 ? yv * log(tv * rho0) : 0.0) - lg - tv * rho0;
          const double d1 = (yv > 0 ? yv / rho1 : 0.0) - tv;
          const double d0 = ((yv > 0 && rho0 > 0) ? yv / rho0 : 0.0) - tv;
          s1 += d1;
          s0 += d0;
          const double g = d1 * mu;
          ga[0] += g;
          for (int k = 0; k < Ko; ++k) ga[k + 1] += g * wv[k];
        }
        /* clamped log psi / log(1 - psi), decisions in log space */
        const double te = exp(-fabs(eta)), le = log1p(te), inv = 1.0 / (1.0 + te);
        const double psi = eta >= 0 ? inv : te * inv;
        const double lp0 = (eta < 0 ? eta : 0) - le, l10 = -(eta > 0 ? eta : 0) - le;
        const int lo = lp0 <= log_tiny, hi = l10 <= log_eps, in_psi = !(lo || hi);
        const double lpsi = lo ? log_tiny : (hi ? log1m_eps : lp0);
        const double l1psi = lo ? neg_tiny : (hi ? log_eps : l10);
        const double av = lpsi + t1;
        double ell, r;
        if (b_is_neginf) {
          ell = av;
          r = 1.0;
        } else {
          const double bv = l1psi + t0, dd = av - bv, td = exp(-fabs(dd)), iv = 1.0 / (1.0 + td);
          r = dd >= 0 ? iv : td * iv;
          ell = (av > bv ? av : bv) + log1p(td);
        }
        const double geta = in_psi ? r - psi : 0.0;
        const double w0 = r < 1.0 ? (1.0 - r) * s0 : 0.0;
        loc[0] += ell;
        loc[1] += geta;
        for (int k = 0; k < Ks; ++k) loc[2 + k] += geta * x[k];
        for (int k = 0; k <= Ko; ++k) loc[2 + Ks + k] += r * ga[k];
        int q = 3 + Ks + Ko;
        if (fpc) loc[q++] += r * s1 + w0;
        if (fpu) loc[q++] += w0;
      }
#pragma omp critical
      for (int i = 0; i < NQ; ++i) acc[i] += loc[i];
    }
    double lp = acc[0];
    for (int i = 0; i < Ks + Ko + 2; ++i) {
      double g = acc[1 + i];
      if (prior) {
        lp += -0.5 * th[i] * th[i] - 0.91893853320467274178;
        g -= th[i];
      }
      grad[(size_t)ci * D + i] = g;
    }
    int q = Ks + Ko + 2;
    if (fpc) {
      if (prior) lp += -c + xc; /* Exponential(1) on c, + log|dc/dx| */
      grad[(size_t)ci * D + q] = acc[1 + q] * c + (prior ? 1.0 - c : 0.0);
      ++q;
    }
    if (fpu) {
      if (prior) lp += -u + xu;
      grad[(size_t)ci * D + q] = acc[1 + q] * u + (prior ? 1.0 - u : 0.0);
      ++q;
    }
    logp[ci] = lp;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * occu_rn (Royle-Nichols, BASELINE config 3) -- restates
 *   biolith/models/occu_rn.py:124-130    NaN mask + nan_to_num
 *   biolith/models/occu_rn.py:139-143    prob_fp_constant ~ Beta(2, 5), sampled as x = logit(c)
 *   biolith/models/occu_rn.py:188-194 + utils/distributions.py:31-40
 *                                        lambda = exp(eta); N ~ Categorical(logits_k = k log lambda - lambda -
 *                                        lgamma(k + 1), k = 0..K), normalised
 *   biolith/models/occu_rn.py:205-222    r = sigmoid(nu); p_k = 1 - (1 - r)^k; Bernoulli(1 - (1 - p_k)(1 - c)),
 *                                        masked, clamp_probs(tiny, 1 - eps); enumeration over N
 * in double arithmetic with log(1 - P_kj) = k log(1 - r_j) + log(1 - c) carried in log space (the closed
 * form of oracle/occupancy.py:occu_rn_logp_grad; see DESIGN.md section 2 for why not 1 - (1 - r)**N).
 * theta = [beta | alpha | logit c (if fpc)].  Arrays double, reference layout.
 * ---------------------------------------------------------------------------------------------- */
int oracle_occu_rn_logp_grad(int f32_clamps, long S, int P, int J, int Ks, int Ko, int K, const double* y,
                             const double* X, const double* W, const double* theta, int C, int fpc, int prior,
                             int nthreads, double* logp, double* grad) {
  if (Ks > MAXK || Ko > MAXK || K < 1 || K > 4096 || J > 4096) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const double tiny = f32_clamps ? (double)FLT_MIN : DBL_MIN;
  const double eps = f32_clamps ? (double)FLT_EPSILON : DBL_EPSILON;
  /* an fp32 run holds the logs of the clamp constants in fp32 (as the numpy oracle and the kernels do) */
  const double log_tiny = f32_clamps ? (double)logf(FLT_MIN) : log(tiny);
  const double log_eps = f32_clamps ? (double)logf(FLT_EPSILON) : log(eps);
  const double log1m_eps = f32_clamps ? (double)log1pf(-FLT_EPSILON) : log1p(-eps);
  const double neg_tiny = f32_clamps ? (double)log1pf(-FLT_MIN) : log1p(-tiny);
  const double rmax = f32_clamps ? (double)FLT_MAX : DBL_MAX;
  const int D = Ks + Ko + 2 + (fpc != 0);
  const long U = S * (long)P;
  double* lgk = (double*)malloc((size_t)(K + 1) * sizeof(double));
  if (!lgk) return -2;
  for (int k = 0; k <= K; ++k) lgk[k] = lgamma((double)k + 1.0);
  int status = 0;
  for (int ci = 0; ci < C; ++ci) {
    const double* th = theta + (size_t)ci * D;
    const double* b = th;
    const double* a = th + Ks + 1;
    const double xc = fpc ? th[Ks + Ko + 2] : 0.0;
    const double c = fpc ? 1.0 / (1.0 + exp(-xc)) : 0.0;
    const double l1mc = fpc ? log1p(-c) : 0.0;
    double acc[2 * MAXK + 5];
    memset(acc, 0, sizeof(acc));
    const int NQ = D + 1;
#pragma omp parallel
    {
      double loc[2 * MAXK + 5];
      memset(loc, 0, sizeof(loc));
      double* A = (double*)malloc((size_t)(K + 1) * 2 * sizeof(double));   /* A_k, then w_k; log pi_k */
      double* vis = (double*)malloc((size_t)J * (4 + MAXK) * sizeof(double)); /* per visit: m, y, r, u, W */
      if (!A || !vis) {
#pragma omp atomic write
        status = -2;
      }
      double* lpi = A ? A + (K + 1) : NULL;
#pragma omp for schedule(static)
      for (long un = 0; un < U; ++un) {
        if (!A || !vis) continue;
        const long s = un / P;
        int site_nan = 0;
        double x[MAXK], eta = b[0];
        for (int k = 0; k < Ks; ++k) {
          const double v = X[s * Ks + k];
          site_nan |= isnan(v);
          x[k] = n2n_clamped(v, rmax);
          eta += x[k] * b[k + 1];
        }
        const double lam = exp(eta);
        /* normalised truncated-Poisson log-weights */
        double mx = -INFINITY;
        for (int k = 0; k <= K; ++k) {
          lpi[k] = (k > 0 ? (double)k * eta : 0.0) - lgk[k] - lam; /* xlogy(k, lam) = k log lam = k eta */
          if (lpi[k] > mx) mx = lpi[k];
        }
        double z = 0;
        for (int k = 0; k <= K; ++k) z += exp(lpi[k] - mx);
        const double lse = mx + log(z);
        double Epi = 0;
        for (int k = 0; k <= K; ++k) {
          lpi[k] -= lse;
          Epi += (double)k * exp(lpi[k]);
          A[k] = lpi[k];
        }
        for (int j = 0; j < J; ++j) {
          double* v = vis + (size_t)j * (4 + MAXK);
          const double* w = W + ((size_t)un * J + j) * Ko;
          double nu = a[0];
          int cov_nan = site_nan;
          for (int k = 0; k < Ko; ++k) {
            cov_nan |= isnan(w[k]);
            v[4 + k] = n2n_clamped(w[k], rmax);
            nu += v[4 + k] * a[k + 1];
          }
          const double yv = y[(size_t)un * J + j];
          v[0] = (cov_nan || !isfinite(yv)) ? 0.0 : 1.0;
          v[1] = yv;
          const double t = exp(-fabs(nu)), inv = 1.0 / (1.0 + t);
          v[2] = nu >= 0 ? inv : t * inv;                 /* r */
          v[3] = -(nu > 0 ? nu : 0) - log1p(t);           /* log(1 - r) */
          if (v[0] == 0.0) continue;
          for (int k = 0; k <= K; ++k) {
            const double lq = (double)k * v[3] + l1mc, Pk = -expm1(lq);
            const int inr = (lq > log_eps) && (Pk > tiny);
            if (yv > 0.5) A[k] += inr ? log(Pk) : (Pk <= tiny ? log_tiny : log1m_eps);
            else A[k] += inr ? lq : (Pk <= tiny ? neg_tiny : log_eps);
          }
        }
        mx = -INFINITY;
        for (int k = 0; k <= K; ++k) if (A[k] > mx) mx = A[k];
        z = 0;
        for (int k = 0; k <= K; ++k) z += exp(A[k] - mx);
        const double ell = mx + log(z);
        double Ew = 0;
        for (int k = 0; k <= K; ++k) {
          A[k] = exp(A[k] - ell); /* posterior weight of N = k */
          Ew += (double)k * A[k];
        }
        const double geta = Ew - Epi;
        double ga[MAXK + 1], gc = 0;
        for (int k = 0; k <= Ko; ++k) ga[k] = 0;
        for (int j = 0; j < J; ++j) {
          const double* v = vis + (size_t)j * (4 + MAXK);
          if (v[0] == 0.0) continue;
          double dnu = 0;
          for (int k = 0; k <= K; ++k) {
            const double lq = (double)k * v[3] + l1mc, Pk = -expm1(lq);
            const int inr = (lq > log_eps) && (Pk > tiny);
            const double dt = inr ? (v[1] > 0.5 ? -exp(lq) / Pk : 1.0) : 0.0; /* dt / dlq */
            dnu += A[k] * dt * (-(double)k * v[2]);
            gc += A[k] * dt;
          }
          ga[0] += dnu;
          for (int k = 0; k < Ko; ++k) ga[k + 1] += dnu * v[4 + k];
        }
        loc[0] += ell;
        loc[1] += geta;
        for (int k = 0; k < Ks; ++k) loc[2 + k] += geta * x[k];
        for (int k = 0; k <= Ko; ++k) loc[2 + Ks + k] += ga[k];
        if (fpc) loc[3 + Ks + Ko] += gc * (-1.0 / (1.0 - c));
      }
      free(A);
      free(vis);
#pragma omp critical
      for (int i = 0; i < NQ; ++i) acc[i] += loc[i];
    }
    double lp = acc[0];
    for (int i = 0; i < Ks + Ko + 2; ++i) {
      double g = acc[1 + i];
      if (prior) {
        lp += -0.5 * th[i] * th[i] - 0.91893853320467274178;
        g -= th[i];
      }
      grad[(size_t)ci * D + i] = g;
    }
    if (fpc) {
      const int q = Ks + Ko + 2;
      const double pa = 2.0, pb = 5.0;
      const double lsc = -log1p(exp(-xc)) , ls1 = -log1p(exp(xc)); /* log c, log(1 - c) */
      if (prior) lp += pa * lsc + pb * ls1 + lgamma(pa + pb) - lgamma(pa) - lgamma(pb);
      grad[(size_t)ci * D + q] = acc[1 + q] * c * (1.0 - c) + (prior ? pa * (1.0 - c) - pb * c : 0.0);
    }
    logp[ci] = lp;
  }
  free(lgk);
  return status;
}

/* ------------------------------------------------------------------------------------------------
 * nmixture (Royle 2004) -- restates biolith/models/nmixture.py:124-220: lambda = exp(eta);
 * N in {max_j y_j .. K} with truncated, un-renormalised Poisson weights (the Categorical normaliser and the
 * numpyro.factor cancel); y_j ~ Binomial(N, sigmoid(nu_j)), NaN-masked.  Double arithmetic, no clamps in this
 * model.  Closed form: oracle/occupancy.py:nmixture_logp_grad.  theta = [beta | alpha].
 * ---------------------------------------------------------------------------------------------- */
int oracle_nmixture_logp_grad(int f32_data, long S, int P, int J, int Ks, int Ko, int K, const double* y,
                              const double* X, const double* W, const double* theta, int C, int prior,
                              int nthreads, double* logp, double* grad) {
  if (Ks > MAXK || Ko > MAXK || K < 1 || K > 4096 || J > 4096) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const double rmax = f32_data ? (double)FLT_MAX : DBL_MAX;
  const int D = Ks + Ko + 2, NQ = D + 1;
  const long U = S * (long)P;
  double* lgk = (double*)malloc((size_t)(K + 1) * sizeof(double));
  if (!lgk) return -2;
  for (int k = 0; k <= K; ++k) lgk[k] = lgamma((double)k + 1.0);
  int status = 0;
  for (int ci = 0; ci < C; ++ci) {
    const double* th = theta + (size_t)ci * D;
    const double* b = th;
    const double* a = th + Ks + 1;
    double acc[2 * MAXK + 5];
    memset(acc, 0, sizeof(acc));
#pragma omp parallel
    {
      double loc[2 * MAXK + 5];
      memset(loc, 0, sizeof(loc));
      double* A = (double*)malloc((size_t)(K + 1) * sizeof(double));
      double* vis = (double*)malloc((size_t)J * (3 + MAXK) * sizeof(double)); /* per visit: m, y, p, W */
      if (!A || !vis) {
#pragma omp atomic write
        status = -2;
      }
#pragma omp for schedule(static)
      for (long un = 0; un < U; ++un) {
        if (!A || !vis) continue;
        const long s = un / P;
        int site_nan = 0;
        double x[MAXK], eta = b[0];
        for (int k = 0; k < Ks; ++k) {
          const double v = X[s * Ks + k];
          site_nan |= isnan(v);
          x[k] = n2n_clamped(v, rmax);
          eta += x[k] * b[k + 1];
        }
        const double lam = exp(eta);
        double Usum = 0, V = 0;
        int kmin = 0;
        for (int j = 0; j < J; ++j) {
          double* v = vis + (size_t)j * (3 + MAXK);
          const double* w = W + ((size_t)un * J + j) * Ko;
          double nu = a[0];
          int cov_nan = site_nan;
          for (int k = 0; k < Ko; ++k) {
            cov_nan |= isnan(w[k]);
            v[3 + k] = n2n_clamped(w[k], rmax);
            nu += v[3 + k] * a[k + 1];
          }
          const double yv = y[(size_t)un * J + j];
          v[0] = (cov_nan || !isfinite(yv)) ? 0.0 : 1.0;
          v[1] = yv;
          const double t = exp(-fabs(nu)), inv = 1.0 / (1.0 + t);
          v[2] = nu >= 0 ? inv : t * inv; /* p */
          if (v[0] == 0.0) continue;
          Usum += -(nu > 0 ? nu : 0) - log1p(t); /* log(1 - p) */
          V += yv * nu;
          if ((int)yv > kmin) kmin = (int)yv;
        }
        double mx = -INFINITY;
        for (int k = kmin; k <= K; ++k) {
          double ck = -lgk[k];
          for (int j = 0; j < J; ++j) {
            const double* v = vis + (size_t)j * (3 + MAXK);
            if (v[0] != 0.0) ck += lgk[k] - lgk[(int)v[1]] - lgk[k - (int)v[1]];
          }
          A[k] = (double)k * (eta + Usum) - lam + ck;
          if (A[k] > mx) mx = A[k];
        }
        double z = 0, Ek = 0;
        for (int k = kmin; k <= K; ++k) {
          const double e = exp(A[k] - mx);
          z += e;
          Ek += (double)k * e;
        }
        Ek /= z;
        const double ell = V + mx + log(z), geta = Ek - lam;
        loc[0] += ell;
        loc[1] += geta;
        for (int k = 0; k < Ks; ++k) loc[2 + k] += geta * x[k];
        for (int j = 0; j < J; ++j) {
          const double* v = vis + (size_t)j * (3 + MAXK);
          if (v[0] == 0.0) continue;
          const double dnu = v[1] - v[2] * Ek;
          loc[2 + Ks] += dnu;
          for (int k = 0; k < Ko; ++k) loc[3 + Ks + k] += dnu * v[3 + k];
        }
      }
      free(A);
      free(vis);
#pragma omp critical
      for (int i = 0; i < NQ; ++i) acc[i] += loc[i];
    }
    double lp = acc[0];
    for (int i = 0; i < D; ++i) {
      double g = acc[1 + i];
      if (prior) {
        lp += -0.5 * th[i] * th[i] - 0.91893853320467274178;
        g -= th[i];
      }
      grad[(size_t)ci * D + i] = g;
    }
    logp[ci] = lp;
  }
  free(lgk);
  return status;
}

/* ------------------------------------------------------------------------------------------------
 * occu_cs (continuous-score occupancy) -- restates biolith/models/occu_cs.py:146-223:
 * z ~ Bernoulli(psi~) and f_j ~ Bernoulli(clamp(z sigmoid(nu_j))) enumerated, s_j ~ Normal((1-f) mu0 + f mu1,
 * (1-f) sigma0 + f sigma1) NaN-masked; extras in unconstrained space [mu0, log(mu1 - mu0), log sigma0,
 * log sigma1] with Normal(0, s) / left-truncated Normal(0, s) / Gamma(a, b) priors and their log-Jacobians.
 * Double arithmetic, clamp constants of either dtype.  Closed form: oracle/occupancy.py:occu_cs_logp_grad.
 * ---------------------------------------------------------------------------------------------- */
static inline double softplus_d(double x) { return (x > 0 ? x : 0) + log1p(exp(-fabs(x))); }
static inline double sigmoid_d(double x) { const double t = exp(-fabs(x)), inv = 1.0 / (1.0 + t); return x >= 0 ? inv : t * inv; }

int oracle_occu_cs_logp_grad(int f32_clamps, long S, int P, int J, int Ks, int Ko, const double* y, const double* X,
                             const double* W, const double* theta, int C, int prior, double prior_mu_scale,
                             double prior_sigma_a, double prior_sigma_b, int nthreads, double* logp, double* grad) {
  if (Ks > MAXK || Ko > MAXK) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const double log_tiny = f32_clamps ? (double)logf(FLT_MIN) : log(DBL_MIN);
  const double log_eps = f32_clamps ? (double)logf(FLT_EPSILON) : log(DBL_EPSILON);
  const double log1m_eps = f32_clamps ? (double)log1pf(-FLT_EPSILON) : log1p(-DBL_EPSILON);
  const double neg_tiny = f32_clamps ? (double)log1pf(-FLT_MIN) : log1p(-DBL_MIN);
  const double rmax = f32_clamps ? (double)FLT_MAX : DBL_MAX;
  const double h2pi = 0.91893853320467274178;
  const int D = Ks + Ko + 6, NQ = D + 1, i0 = Ks + Ko + 2;
  const long U = S * (long)P;
  for (int ci = 0; ci < C; ++ci) {
    const double* th = theta + (size_t)ci * D;
    const double* b = th;
    const double* a = th + Ks + 1;
    const double mu0 = th[i0], e1x = exp(th[i0 + 1]), mu1 = mu0 + e1x, xs0 = th[i0 + 2], xs1 = th[i0 + 3];
    const double sg0 = exp(xs0), sg1 = exp(xs1);
    double acc[2 * MAXK + 8];
    memset(acc, 0, sizeof(acc));
#pragma omp parallel
    {
      double loc[2 * MAXK + 8];
      memset(loc, 0, sizeof(loc));
#pragma omp for schedule(static)
      for (long un = 0; un < U; ++un) {
        const long s = un / P;
        int site_nan = 0;
        double x[MAXK], eta = b[0];
        for (int k = 0; k < Ks; ++k) {
          const double v = X[s * Ks + k];
          site_nan |= isnan(v);
          x[k] = n2n_clamped(v, rmax);
          eta += x[k] * b[k + 1];
        }
        double L1 = 0, L0 = 0, ga[MAXK + 1];
        double Ae0 = 0, Aq0 = 0, Ae1 = 0, Aq1 = 0, Be0 = 0, Bq0 = 0, Be1 = 0, Bq1 = 0;
        for (int k = 0; k <= Ko; ++k) ga[k] = 0;
        for (int j = 0; j < J; ++j) {
          const double* w = W + ((size_t)un * J + j) * Ko;
          double wv[MAXK], nu = a[0];
          int cov_nan = site_nan;
          for (int k = 0; k < Ko; ++k) {
            cov_nan |= isnan(w[k]);
            wv[k] = n2n_clamped(w[k], rmax);
            nu += wv[k] * a[k + 1];
          }
          const double sc = y[(size_t)un * J + j];
          if (cov_nan || !isfinite(sc)) continue; /* mask_missing_obs */
          const double e0 = (sc - mu0) / sg0, e1 = (sc - mu1) / sg1;
          const double n0 = -0.5 * e0 * e0 - xs0 - h2pi, n1 = -0.5 * e1 * e1 - xs1 - h2pi;
          /* clamped log q~, log(1 - q~) of q = sigmoid(nu) */
          const double pj = sigmoid_d(nu), lq0 = -softplus_d(-nu), l1q0 = -softplus_d(nu);
          const int lo = lq0 <= log_tiny, hi = l1q0 <= log_eps, inr = !(lo || hi);
          const double lq = lo ? log_tiny : (hi ? log1m_eps : lq0), l1q = lo ? neg_tiny : (hi ? log_eps : l1q0);
          const double a0 = l1q + n0, a1 = lq + n1;           /* z = 1 */
          const double c0 = neg_tiny + n0, c1 = log_tiny + n1; /* z = 0: q~ = tiny */
          L1 += a0 + softplus_d(a1 - a0);
          L0 += c0 + softplus_d(c1 - c0);
          const double w1 = sigmoid_d(a1 - a0), v1 = sigmoid_d(c1 - c0);
          const double g = inr ? w1 - pj : 0.0;
          ga[0] += g;
          for (int k = 0; k < Ko; ++k) ga[k + 1] += g * wv[k];
          const double q0 = e0 * e0 - 1.0, q1 = e1 * e1 - 1.0;
          Ae0 += (1 - w1) * e0; Aq0 += (1 - w1) * q0; Ae1 += w1 * e1; Aq1 += w1 * q1;
          Be0 += (1 - v1) * e0; Bq0 += (1 - v1) * q0; Be1 += v1 * e1; Bq1 += v1 * q1;
        }
        const double psi = sigmoid_d(eta), lp0 = -softplus_d(-eta), l10 = -softplus_d(eta);
        const int lo = lp0 <= log_tiny, hi = l10 <= log_eps, in_psi = !(lo || hi);
        const double lpsi = lo ? log_tiny : (hi ? log1m_eps : lp0), l1psi = lo ? neg_tiny : (hi ? log_eps : l10);
        const double av = lpsi + L1, bv = l1psi + L0;
        const double ell = bv + softplus_d(av - bv), r = sigmoid_d(av - bv);
        const double geta = in_psi ? r - psi : 0.0;
        loc[0] += ell;
        loc[1] += geta;
        for (int k = 0; k < Ks; ++k) loc[2 + k] += geta * x[k];
        for (int k = 0; k <= Ko; ++k) loc[2 + Ks + k] += r * ga[k];
        const double g_mu0 = (r * Ae0 + (1 - r) * Be0) / sg0, g_mu1 = (r * Ae1 + (1 - r) * Be1) / sg1;
        loc[1 + i0] += g_mu0 + g_mu1; /* mu1 = mu0 + exp(x1) */
        loc[2 + i0] += g_mu1 * e1x;
        loc[3 + i0] += r * Aq0 + (1 - r) * Bq0;
        loc[4 + i0] += r * Aq1 + (1 - r) * Bq1;
      }
#pragma omp critical
      for (int i = 0; i < NQ; ++i) acc[i] += loc[i];
    }
    double lp = acc[0];
    for (int i = 0; i < i0; ++i) {
      double g = acc[1 + i];
      if (prior) {
        lp += -0.5 * th[i] * th[i] - h2pi;
        g -= th[i];
      }
      grad[(size_t)ci * D + i] = g;
    }
    double ge[4] = {acc[1 + i0], acc[2 + i0], acc[3 + i0], acc[4 + i0]};
    if (prior) {
      const double sm = prior_mu_scale, pa = prior_sigma_a, pb = prior_sigma_b, t = mu0 / sm;
      const double sf = 0.5 * erfc(t * 0.70710678118654752440); /* 1 - Phi(t) */
      const double hazard = exp(-0.5 * t * t - h2pi) / sf;
      lp += -0.5 * t * t - log(sm) - h2pi;
      lp += -0.5 * (mu1 / sm) * (mu1 / sm) - log(sm) - h2pi - log(sf) + th[i0 + 1];
      ge[0] += -mu0 / (sm * sm) - mu1 / (sm * sm) + hazard / sm;
      ge[1] += -mu1 / (sm * sm) * e1x + 1.0;
      for (int i = 0; i < 2; ++i) {
        const double xs = th[i0 + 2 + i], sg = i ? sg1 : sg0;
        lp += pa * log(pb) + (pa - 1.0) * xs - pb * sg - lgamma(pa) + xs;
        ge[2 + i] += pa - pb * sg;
      }
    }
    for (int i = 0; i < 4; ++i) grad[(size_t)ci * D + i0 + i] = ge[i];
    logp[ci] = lp;
  }
  return 0;
}
