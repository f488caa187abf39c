/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, per-site restatement of one iteration of the reference's SPGG loop
 * (reference paths relative to /root/reference):
 *   src/model/spgg.py:368-592   (SPGG.run loop body)
 *   src/model/algorithms.py:96-133 (QLearning.select_action / update_q_table)
 * It is the checker for the CUDA path; the product never links or loads it.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
 *
 * Parity status: PINNED.  oracle_step_f64 is compared bit-for-bit (S, R, Q and
 * the integer statistics) against the executed reference through
 * oracle/spgg_numpy.py, the golden fixtures in tests/golden/ and, inside the
 * build container, the reference itself (tests/test_oracle_vs_reference.py).
 * The reference has no tests/golden vectors of its own.
 *
 * oracle_step_f32 restates the *throughput-mode* arithmetic of the CUDA
 * kernel (fp32 Q, exact-count reward table, explicit fmaf, Philox draws); it
 * has no bit-level counterpart in the reference and is pinned statistically
 * (cooperation-rate curves inside the reference's seed-to-seed band) and
 * against oracle_step_f64 on dyadic parameter sets where both are exact.
 *
 * Build with -ffp-contract=off (see oracle/Makefile): the reference is NumPy,
 * every operation is individually rounded.
 */
#include "spgg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* neighbour offsets (dx,dy) in the reference's enumeration order;
 * neighbour k of (i,j) is (i-dx, j-dy) because np.roll(X,(dx,dy))[i,j]==X[i-dx,j-dy].
 * spgg.py:479-485 */
static const int OFF[12][2] = {{1, 0},  {-1, 0}, {0, 1},  {0, -1}, {2, 0},  {-2, 0},
                               {0, 2},  {0, -2}, {1, 1},  {1, -1}, {-1, 1}, {-1, -1}};
/* groups a site belongs to, in the summation order of spgg.py:373-377:
 * centres (i,j), (i-1,j), (i+1,j), (i,j-1), (i,j+1) */
static const int GRP[5][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}};

static inline int wrapi(int a, int L) {
  a %= L;
  return a < 0 ? a + L : a;
}
#define IDX(i, j) ((size_t)wrapi((i), L) * (size_t)L + (size_t)wrapi((j), L))

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- Philox */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}

uint32_t oracle_thr24(double eps) {
  double t = ceil(eps * 16777216.0);
  if (t < 0) t = 0;
  if (t > 16777216.0) t = 16777216.0;
  return (uint32_t)t;
}

/* draw word of site (i,j) at iteration `step`: counter = (j>>2, row, step, 0), word index
 * j&3 - one Philox call serves four consecutive sites of a row. */
static inline uint32_t site_word(uint64_t seed, uint32_t step, int i, int j) {
  uint32_t ctr[4] = {(uint32_t)(j >> 2), (uint32_t)i, step, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t w[4];
  oracle_philox4x32_10(ctr, key, w);
  return w[j & 3];
}

/* ------------------------------------------------------------ shared bits */
static void count_groups(int L, const uint8_t *S, uint8_t *N) {
  /* spgg.py:23-36 on the cooperator mask */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int n = (S[IDX(i, j)] == 0) + (S[IDX(i + 1, j)] == 0) + (S[IDX(i - 1, j)] == 0) +
              (S[IDX(i, j + 1)] == 0) + (S[IDX(i, j - 1)] == 0);
      N[(size_t)i * L + j] = (uint8_t)n;
    }
}

static void zero_stats(double *st) {
  if (st) memset(st, 0, sizeof(double) * OST_NSTAT);
}

/* =================================================================== fp64 */
int oracle_step_f64(const oracle_params_t *p, uint8_t *S, double *R, double *Q, double eps,
                    const double *u, const uint8_t *b, uint64_t seed, uint32_t step,
                    uint32_t thr24, double *stats) {
  const int L = p->L;
  const size_t n_sites = (size_t)L * L;
  const int nk = (p->M == 2) ? 12 : 4;
  const int nst = nk + 1;
  const double rc = p->r * p->c;
  double g[6];
  for (int n = 0; n < 6; ++n) g[n] = rc * (double)n / 5.0; /* spgg.py:256 */
  const double lo = p->r - 5.0;                            /* spgg.py:149 */
  const double span = 4.0 * p->r - lo;                     /* spgg.py:148,377 */
  const double wP = p->wP, wR = 1.0 - p->wP;               /* spgg.py:108 */

  uint8_t *N = (uint8_t *)malloc(n_sites);
  uint8_t *a = (uint8_t *)malloc(n_sites);
  uint8_t *s_old = (uint8_t *)malloc(n_sites);
  uint8_t *s_new = (uint8_t *)malloc(n_sites);
  double *Rn = (double *)malloc(n_sites * sizeof(double));
  double *P = (double *)malloc(n_sites * sizeof(double));
  double *rew = (double *)malloc(n_sites * sizeof(double));
  zero_stats(stats);

  count_groups(L, S, N);

  /* payoff of the configuration before the action: spgg.py:373-377 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      const double C = (S[x] == 0) ? 1.0 : 0.0, D = 1.0 - C;
      double tot = 0.0;
      for (int q = 0; q < 5; ++q) {
        const double share = g[N[IDX(i + GRP[q][0], j + GRP[q][1])]];
        const double term = (share - p->cost) * C + share * D;
        tot = (q == 0) ? term : tot + term;
      }
      P[x] = (tot - lo) / span;
    }

  /* state, action, reputation: spgg.py:409-415, algorithms.py:102-110 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      int s;
      if (p->state_mode == 1) {
        s = (S[x] == 0);
      } else {
        double acc = 0.0;
        acc += R[x];
        for (int k = 0; k < nk; ++k) acc += R[IDX(i - OFF[k][0], j - OFF[k][1])];
        s = (acc / (double)nst > 0.0);
      }
      s_old[x] = (uint8_t)s;
      int explore, rnd;
      if (u) {
        explore = (u[x] < eps);
        rnd = b[x];
      } else {
        const uint32_t w = site_word(seed, step, i, j);
        explore = ((w >> 8) < thr24);
        rnd = (int)(w & 1u);
      }
      const double *q = Q + 4 * x + 2 * s;
      const int greedy = (q[1] > q[0]) ? 1 : 0; /* np.argmax: first max */
      a[x] = (uint8_t)(explore ? rnd : greedy);
      double rn = R[x] + (a[x] == 0 ? p->rep_gain_C : -p->delta_R_D); /* spgg.py:321 */
      rn = fmax(rn, p->R_min);
      rn = fmin(rn, p->R_max);
      Rn[x] = rn;
    }

  /* new state + rewards: spgg.py:423-427 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      int s;
      if (p->state_mode == 1) {
        s = (a[x] == 0);
      } else {
        double acc = 0.0;
        acc += Rn[x];
        for (int k = 0; k < nk; ++k) acc += Rn[IDX(i - OFF[k][0], j - OFF[k][1])];
        s = (acc / (double)nst > 0.0);
      }
      s_new[x] = (uint8_t)s;
      const double rr = (a[x] == 0) ? 0.5 : 0.0;
      rew[x] = wP * P[x] + wR * rr;
    }

  /* lattice-global max |diff|: spgg.py:486-488 */
  double gmax = 0.0;
#pragma omp parallel for schedule(static) reduction(max : gmax)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      for (int k = 0; k < nk; ++k) {
        const double d = fabs(rew[IDX(i - OFF[k][0], j - OFF[k][1])] - rew[x]);
        if (d > gmax) gmax = d;
      }
    }

  /* updates + statistics */
  double acc_st[OST_NSTAT];
  memset(acc_st, 0, sizeof(acc_st));
  /* plain sequential accumulation: statistics are compared with a tolerance */
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      const int s = s_old[x], sn = s_new[x], act = a[x];
      double *q = Q + 4 * x;
      const double q0 = q[2 * s + act];
      const double mx = fmax(q[2 * sn], q[2 * sn + 1]);
      const double td = rew[x] + p->gamma * mx - q0; /* algorithms.py:128 */
      const double qtd = q0 + p->alpha * td;         /* algorithms.py:131 */
      double best = 0.0;
      int kstar = 0;
      for (int k = 0; k < nk; ++k) {
        const double d = rew[IDX(i - OFF[k][0], j - OFF[k][1])] - rew[x];
        if (k == 0 || d > best) {
          best = d;
          kstar = k;
        }
      }
      const double lam = p->kappa * fmax(0.0, best) / (gmax + p->lambda_eps); /* spgg.py:489 */
      const int a_star = a[IDX(i - OFF[kstar][0], j - OFF[kstar][1])];
      const double nu = lam * ((a_star == act) ? 1.0 : -1.0);
      /* NI statistic uses the table after the TD write (spgg.py:446-473,512) */
      q[2 * s + act] = qtd;
      const double mx2 = fmax(q[2 * sn], q[2 * sn + 1]);
      const double td2 = rew[x] + p->gamma * mx2 - qtd;
      q[2 * s + act] = qtd + nu; /* spgg.py:509 */
      const double pct = fabs(nu) / (fabs(p->alpha * td2) + fabs(nu) + 1e-8) * 100.0;

      const int wasC = (S[x] == 0);
      acc_st[OST_NC_OLD] += wasC;
      acc_st[OST_N_CD] += (wasC && act == 1);
      acc_st[OST_N_DC] += (!wasC && act == 0);
      acc_st[OST_NC_NEW] += (act == 0);
      acc_st[OST_SUM_P] += P[x];
      acc_st[wasC ? OST_SUM_P_C : OST_SUM_P_D] += P[x];
      acc_st[OST_SUM_WP_P] += wP * P[x];
      acc_st[act == 0 ? OST_SUM_REW_C : OST_SUM_REW_D] += rew[x];
      if (act == 0) acc_st[OST_SUM_RATIO] += fabs(wR * 0.5) / (fabs(rew[x]) + 1e-9) * 100.0;
      acc_st[OST_SUM_R] += R[x];
      for (int e = 0; e < 4; ++e) {
        acc_st[OST_SUM_Q + e] += q[e];
        acc_st[(wasC ? OST_SUM_Q_C : OST_SUM_Q_D) + e] += q[e];
      }
      acc_st[OST_SUM_NI] += pct;
      if (best > 0.0) {
        acc_st[OST_N_BEST_POS] += 1;
        acc_st[OST_N_BEST_2ND] += (kstar >= 4);
      }
      const int nd = (a[IDX(i, j)] == 1) + (a[IDX(i + 1, j)] == 1) + (a[IDX(i - 1, j)] == 1) +
                     (a[IDX(i, j + 1)] == 1) + (a[IDX(i, j - 1)] == 1);
      acc_st[OST_GROUP0 + nd] += 1;
    }
  acc_st[OST_GMAX] = gmax;
  if (stats) memcpy(stats, acc_st, sizeof(acc_st));

  memcpy(S, a, n_sites);
  memcpy(R, Rn, n_sites * sizeof(double));
  free(N); free(a); free(s_old); free(s_new); free(Rn); free(P); free(rew);
  return 0;
}

/* =================================================================== fp32 */
void oracle_reward_table(const oracle_params_t *p, float *tab) {
  const double rc = p->r * p->c;
  const double lo = p->r - 5.0, span = 4.0 * p->r - lo;
  const double wP = p->wP, wR = 1.0 - p->wP;
  for (int code = 0; code < 128; ++code) {
    const int sn = code >> 2, C = (code >> 1) & 1, coop = code & 1;
    /* exact-count payoff: sum over the 5 groups of (rc*N/5 - cost*C) */
    const double tot = rc * (double)sn / 5.0 - (C ? 5.0 * p->cost : 0.0);
    const double P = (tot - lo) / span;
    tab[code] = (float)(wP * P + wR * (coop ? 0.5 : 0.0));
  }
}

static double payoff_from_count(const oracle_params_t *p, int sn, int C) {
  const double rc = p->r * p->c;
  const double lo = p->r - 5.0, span = 4.0 * p->r - lo;
  const double tot = rc * (double)sn / 5.0 - (C ? 5.0 * p->cost : 0.0);
  return (tot - lo) / span;
}

int oracle_step_f32(const oracle_params_t *p, uint8_t *S, float *R, float *Q, double eps,
                    const double *u, const uint8_t *b, uint64_t seed, uint32_t step,
                    uint32_t thr24, double *stats) {
  const int L = p->L;
  const size_t n_sites = (size_t)L * L;
  const int nk = (p->M == 2) ? 12 : 4;
  const int nst = nk + 1;
  const float alpha = (float)p->alpha, gamma = (float)p->gamma, kappa = (float)p->kappa;
  const float leps = (float)p->lambda_eps;
  const float gainC = (float)p->rep_gain_C, lossD = (float)p->delta_R_D;
  const float rmin = (float)p->R_min, rmax = (float)p->R_max;
  const double wP = p->wP, wR = 1.0 - p->wP;
  float tab[128];
  oracle_reward_table(p, tab);

  uint8_t *N = (uint8_t *)malloc(n_sites);
  uint8_t *SN = (uint8_t *)malloc(n_sites);
  uint8_t *a = (uint8_t *)malloc(n_sites);
  uint8_t *s_old = (uint8_t *)malloc(n_sites);
  uint8_t *s_new = (uint8_t *)malloc(n_sites);
  float *Rn = (float *)malloc(n_sites * sizeof(float));
  float *rew = (float *)malloc(n_sites * sizeof(float));
  zero_stats(stats);

  count_groups(L, S, N);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int sn = 0;
      for (int q = 0; q < 5; ++q) sn += N[IDX(i + GRP[q][0], j + GRP[q][1])];
      SN[(size_t)i * L + j] = (uint8_t)sn;
    }

#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      int s;
      if (p->state_mode == 1) {
        s = (S[x] == 0);
      } else {
        float acc = R[x];
        for (int k = 0; k < nk; ++k) acc += R[IDX(i - OFF[k][0], j - OFF[k][1])];
        s = (acc / (float)nst > 0.0f);
      }
      s_old[x] = (uint8_t)s;
      int explore, rnd;
      if (u) {
        explore = (u[x] < eps);
        rnd = b[x];
      } else {
        const uint32_t w = site_word(seed, step, i, j);
        explore = ((w >> 8) < thr24);
        rnd = (int)(w & 1u);
      }
      const float *q = Q + 4 * x + 2 * s;
      const int greedy = (q[1] > q[0]) ? 1 : 0;
      a[x] = (uint8_t)(explore ? rnd : greedy);
      float rn = R[x] + (a[x] == 0 ? gainC : -lossD);
      rn = fmaxf(rn, rmin);
      rn = fminf(rn, rmax);
      Rn[x] = rn;
    }

#pragma omp parallel for schedule(static)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      int s;
      if (p->state_mode == 1) {
        s = (a[x] == 0);
      } else {
        float acc = Rn[x];
        for (int k = 0; k < nk; ++k) acc += Rn[IDX(i - OFF[k][0], j - OFF[k][1])];
        s = (acc / (float)nst > 0.0f);
      }
      s_new[x] = (uint8_t)s;
      rew[x] = tab[(SN[x] << 2) | ((S[x] == 0) << 1) | (a[x] == 0)];
    }

  float gmax = 0.0f;
#pragma omp parallel for schedule(static) reduction(max : gmax)
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      for (int k = 0; k < nk; ++k) {
        const float d = fabsf(rew[IDX(i - OFF[k][0], j - OFF[k][1])] - rew[x]);
        if (d > gmax) gmax = d;
      }
    }
  const float inv_den = 1.0f / (gmax + leps);

  double acc_st[OST_NSTAT];
  memset(acc_st, 0, sizeof(acc_st));
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const size_t x = (size_t)i * L + j;
      const int s = s_old[x], sn = s_new[x], act = a[x];
      float *q = Q + 4 * x;
      const float q0 = q[2 * s + act];
      const float mx = fmaxf(q[2 * sn], q[2 * sn + 1]);
      const float td = fmaf(gamma, mx, rew[x]) - q0;
      const float qtd = fmaf(alpha, td, q0);
      float best = 0.0f;
      int kstar = 0;
      for (int k = 0; k < nk; ++k) {
        const float d = rew[IDX(i - OFF[k][0], j - OFF[k][1])] - rew[x];
        if (k == 0 || d > best) {
          best = d;
          kstar = k;
        }
      }
      const float lam = (kappa * fmaxf(0.0f, best)) * inv_den;
      const int a_star = a[IDX(i - OFF[kstar][0], j - OFF[kstar][1])];
      const float nu = (a_star == act) ? lam : -lam;
      q[2 * s + act] = qtd;
      const float mx2 = fmaxf(q[2 * sn], q[2 * sn + 1]);
      const float td2 = fmaf(gamma, mx2, rew[x]) - qtd;
      q[2 * s + act] = qtd + nu;
      const double pct =
          (double)fabsf(nu) / ((double)fabsf(alpha * td2) + (double)fabsf(nu) + 1e-8) * 100.0;

      const int wasC = (S[x] == 0);
      const double Px = payoff_from_count(p, SN[x], wasC);
      acc_st[OST_NC_OLD] += wasC;
      acc_st[OST_N_CD] += (wasC && act == 1);
      acc_st[OST_N_DC] += (!wasC && act == 0);
      acc_st[OST_NC_NEW] += (act == 0);
      acc_st[OST_SUM_P] += Px;
      acc_st[wasC ? OST_SUM_P_C : OST_SUM_P_D] += Px;
      acc_st[OST_SUM_WP_P] += wP * Px;
      const double rewd = wP * Px + wR * (act == 0 ? 0.5 : 0.0);
      acc_st[act == 0 ? OST_SUM_REW_C : OST_SUM_REW_D] += rewd;
      if (act == 0) acc_st[OST_SUM_RATIO] += fabs(wR * 0.5) / (fabs(rewd) + 1e-9) * 100.0;
      acc_st[OST_SUM_R] += R[x];
      for (int e = 0; e < 4; ++e) {
        acc_st[OST_SUM_Q + e] += q[e];
        acc_st[(wasC ? OST_SUM_Q_C : OST_SUM_Q_D) + e] += q[e];
      }
      acc_st[OST_SUM_NI] += pct;
      if (best > 0.0f) {
        acc_st[OST_N_BEST_POS] += 1;
        acc_st[OST_N_BEST_2ND] += (kstar >= 4);
      }
      const int nd = (a[IDX(i, j)] == 1) + (a[IDX(i + 1, j)] == 1) + (a[IDX(i - 1, j)] == 1) +
                     (a[IDX(i, j + 1)] == 1) + (a[IDX(i, j - 1)] == 1);
      acc_st[OST_GROUP0 + nd] += 1;
    }
  acc_st[OST_GMAX] = gmax;
  if (stats) memcpy(stats, acc_st, sizeof(acc_st));

  memcpy(S, a, n_sites);
  memcpy(R, Rn, n_sites * sizeof(float));
  free(N); free(SN); free(a); free(s_old); free(s_new); free(Rn); free(rew);
  return 0;
}
