/* TEST INFRASTRUCTURE ONLY - C restatement of the reference SPGG step.
 * See spgg_oracle.c for the parity status and the reference citations.
 * Nothing under the product package may include, link or load this. */
#ifndef SPGG_ORACLE_H
#define SPGG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* stat-row layout; identical to the device row layout (include/spgg.h) so the
 * parity tests can compare rows column by column. */
enum {
  OST_NC_OLD = 0,      /* #cooperators before the action (S_{t-1})            spgg.py:383 */
  OST_N_CD = 1,        /* C -> D switches                                      spgg.py:419 */
  OST_N_DC = 2,        /* D -> C switches                                      spgg.py:420 */
  OST_NC_NEW = 3,      /* #sites whose action is cooperate                               */
  OST_SUM_P = 4,       /* sum of normalised payoff P                           spgg.py:388 */
  OST_SUM_P_C = 5,     /* ... over sites with S_{t-1}==C                       spgg.py:389 */
  OST_SUM_P_D = 6,     /* ... over sites with S_{t-1}==D                       spgg.py:390 */
  OST_SUM_WP_P = 7,    /* sum of w_P*P                                         spgg.py:425 */
  OST_SUM_REW_C = 8,   /* sum of reward over a==C                              spgg.py:542 */
  OST_SUM_REW_D = 9,   /* sum of reward over a==D                              spgg.py:543 */
  OST_SUM_RATIO = 10,  /* sum over a==C of |wR*0.5|/(|rew|+1e-9)*100           spgg.py:529-536 */
  OST_GROUP0 = 11,     /* 11..16: #groups with k defectors (k=0..5) in S_t     spgg.py:586-592 */
  OST_SUM_R = 17,      /* sum of R before the action (R_{t-1})                 spgg.py:394 */
  OST_SUM_Q = 18,      /* 18..21: sum of Q[:,:,s,a] after both updates         spgg.py:562-565 */
  OST_SUM_Q_C = 22,    /* 22..25: same over S_{t-1}==C                         spgg.py:568-583 */
  OST_SUM_Q_D = 26,    /* 26..29: same over S_{t-1}==D                                    */
  OST_SUM_NI = 30,     /* sum of neighbour-influence percent                   spgg.py:512 */
  OST_N_BEST_POS = 31, /* #sites with max_diff > 0                             spgg.py:521 */
  OST_N_BEST_2ND = 32, /* ... whose arg-max neighbour is second order          spgg.py:520 */
  OST_GMAX = 33,       /* lattice-global max |diff|                            spgg.py:488 */
  OST_NSTAT = 40
};

typedef struct {
  int32_t L;
  int32_t M;          /* 1 or 2 (use_second_order)           */
  int32_t state_mode; /* 0 = reputation, 1 = action          */
  int32_t reserved;
  double r, c, cost;
  double alpha, gamma;
  double kappa;       /* influence_factor                    */
  double lambda_eps;
  double rep_gain_C, delta_R_D, R_min, R_max;
  double wP;          /* reward_weight_payoff                */
} oracle_params_t;

/* One iteration, fp64, reference operation order, replayed draws.
 * S: L*L bytes (0=C, 1=D); R: L*L doubles; Q: L*L*4 doubles laid out (s,a).
 * u,b: that iteration's draw arrays (algorithms.py:105,108).
 * If u == NULL the draws come from Philox4x32-10 (seed, step) with the
 * integer threshold thr24 (explore <=> (word>>8) < thr24; random action = word&1).
 * stats: OST_NSTAT doubles (may be NULL). Returns 0. */
int oracle_step_f64(const oracle_params_t *p, uint8_t *S, double *R, double *Q,
                    double eps, const double *u, const uint8_t *b,
                    uint64_t seed, uint32_t step, uint32_t thr24, double *stats);

/* One iteration with the throughput-mode arithmetic: fp32 Q, fp32 R, rewards
 * from the integer table rew[SigmaN][C_old][coop_new] (exact-count payoff),
 * explicit fmaf, Philox draws (or replayed u,b when u != NULL; u compared in
 * double against eps). */
int oracle_step_f32(const oracle_params_t *p, uint8_t *S, float *R, float *Q,
                    double eps, const double *u, const uint8_t *b,
                    uint64_t seed, uint32_t step, uint32_t thr24, double *stats);

/* Philox4x32-10, exposed for known-answer tests. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* explore threshold used by both implementations: ceil(eps * 2^24) clamped to [0, 2^24] */
uint32_t oracle_thr24(double eps);

/* reward table of the throughput mode: 128 floats indexed (SigmaN<<2 | C_old<<1 | coop_new) */
void oracle_reward_table(const oracle_params_t *p, float *tab128);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
