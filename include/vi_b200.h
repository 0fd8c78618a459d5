/*
 * vi_b200.h -- C ABI of the value-iteration kernels (secondary path; same libpomcp_b200.so).
 *
 * Replaces ValueIteration.value_iteration / prune of the reference (pomdpy/solvers/value_iteration.py:24-142).
 * Plain pointers and sizes; host arrays unless a parameter is named *_dev.
 */
#ifndef VI_B200_H
#define VI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* value_iteration(t, o, r, horizon) (value_iteration.py:24-73) with the reference's projection formula
 *   candidate(g,u,z)[i] = discount * (R[u][i] + sum_j (v_g[i] * O[u][i][z]) * T[u][j][i])        (:51-67)
 * and the pointwise-dominance prune its pairwise LP computes (:88-142; delta = 1e-10): vector i is dropped iff some j has
 * max_s(v_i - v_j) < delta and (i does not beat j the same way, or j comes first).  All fp64, no FMA contraction.
 * T [A][S][S], O [A][S][Z], R [A][S]; outputs alpha [capacity][S], action [capacity], *n_out vectors.
 * Returns 0, or < 0 on error (capacity exceeded, CUDA failure; text in err). */
int vi_solve_reference(int device, int S, int A, int Z, const double *T, const double *O, const double *R, double discount,
                       int horizon, int capacity, double *alpha_out, int32_t *action_out, int32_t *n_out,
                       char *err, int errlen);

/* Textbook projection for large |S| (BASELINE config 5):
 *   proj[a][s][o*G + g] = sum_s' T[a][s][s'] * O[a][s'][o] * gamma[g][s']
 * as |A| GEMMs [S x S] . [S x Z*G] on the fp64 tensor cores (DMMA), all pointers on the device, row-major.
 * scratch_dev holds S*Z*G doubles; S % 128 == 0 and (Z*G) % 64 == 0.  Asynchronous on `stream`. */
int vi_project_textbook_dev(int device, int S, int A, int Z, int G, const double *T_dev, const double *O_dev,
                            const double *gamma_dev, double *scratch_dev, double *proj_dev, void *stream,
                            char *err, int errlen);

/* The same projection on the tcgen05 tensor cores, FAST mode: fp64 operands split into TF32 pairs (hi, lo), three
 * kind::tf32 MMAs per k-step (hi.hi + hi.lo + lo.hi) into fp32 accumulators in tensor memory, TMA operand loads, fp64
 * output -- fp32-grade accuracy (the fp64 DMMA path above stays the default and the parity path).
 *   vi_split_tf32_dev:  hi[i] = tf32(x[i]), lo[i] = tf32(x[i] - hi[i])  (T is split once per model, not per backup)
 *   vi_gemm_tf32x3_dev: C[M x N] (fp64, row-major) = A[M x K] . Bt[N x K]^T from the split operands (fp32 arrays,
 *                       row-major = K-major); M % 128 == 0, N % 128 == 0, K % 32 == 0.
 *                       variant: VI_TC_TILE128    one CTA per 128 x 128 tile, K in 4 partial accumulators (most accurate)
 *                                VI_TC_PERSISTENT one CTA per SM walking the tiles, 2 partial accumulators, two accumulator
 *                                                 sets in tensor memory: a tile's epilogue runs under the next tile's MMAs
 *   vi_project_textbook_tf32x3_dev: proj as above from Thi / Tlo [A][S][S]; scratch_dev holds 2*S*Z*G floats. */
#define VI_TC_TILE128 0
#define VI_TC_PERSISTENT 1
int vi_split_tf32_dev(int device, long long n, const double *x_dev, float *hi_dev, float *lo_dev, void *stream,
                      char *err, int errlen);
int vi_gemm_tf32x3_dev(int device, int M, int N, int K, const float *Ahi_dev, const float *Alo_dev,
                       const float *Bthi_dev, const float *Btlo_dev, double *C_dev, int variant, void *stream,
                       char *err, int errlen);
int vi_project_textbook_tf32x3_dev(int device, int S, int A, int Z, int G, const float *Thi_dev, const float *Tlo_dev,
                                   const double *O_dev, const double *gamma_dev, float *scratch_dev, double *proj_dev,
                                   int variant, void *stream, char *err, int errlen);

/* Point-based backup for large |S| (the cross-sum of the exact backup taken at P belief points), all pointers on the
 * device, row-major fp64, asynchronous on `stream`:
 *   g*(a,p,o) = argmax_g b_p . proj[a][:, o, g]      (proj as above; scores by DMMA GEMMs [P x S] . [S x Z*G])
 *   a*(p)     = argmax_a b_p . R[a] + discount * sum_o max_g b_p . proj[a][:, o, g]
 *   alpha_p   = R[a*] + discount * sum_o proj[a*][:, o, g*(a*,p,o)]         -> alpha_dev [P][S], action_dev [P]
 * followed, if `prune`, by the pointwise-dominance prune (delta 1e-10) into pruned_dev / pruned_action_dev / *n_out_dev.
 * Work buffers: scratch [S*Z*G], proj [A*S*Z*G], score [A*P*Z*G] doubles, best [A*P*Z], keep [P] int32, value [A*P].
 * S % 128 == 0, P % 128 == 0, (Z*G) % 64 == 0. */
int vi_point_based_backup_dev(int device, int S, int A, int Z, int G, int P, const double *T_dev, const double *O_dev,
                              const double *R_dev, const double *gamma_dev, const double *beliefs_dev, double discount,
                              double *scratch_dev, double *proj_dev, double *score_dev, int32_t *best_dev,
                              double *value_dev, double *alpha_dev, int32_t *action_dev, int prune, double *pruned_dev,
                              int32_t *pruned_action_dev, int32_t *keep_dev, int32_t *n_out_dev, void *stream,
                              char *err, int errlen);

/* The part of the point-based backup after the projection, for a proj_dev computed by either projection above
 * (vi_point_based_backup_dev = vi_project_textbook_dev + this). */
int vi_point_based_select_dev(int device, int S, int A, int Z, int G, int P, const double *R_dev,
                              const double *beliefs_dev, double discount, const double *proj_dev, double *score_dev,
                              int32_t *best_dev, double *value_dev, double *alpha_dev, int32_t *action_dev, int prune,
                              double *pruned_dev, int32_t *pruned_action_dev, int32_t *keep_dev, int32_t *n_out_dev,
                              void *stream, char *err, int errlen);

/* Policy execution -- Agent.run_value_iteration (pomdpy/agent.py:215-264): per step the arg-max alpha vector gives the
 * action (value_iteration.py:161-181), the environment steps, the belief is updated (tiger_model.py:109-131 generalised
 * to any tabular model), until a terminal action or max_steps.  One block per epoch, the whole episode in one launch:
 * no host round trip per step; `n_epochs` episodes at once.  The environment is the tabular model itself (hidden state
 * s ~ b0, s' ~ T[a][s], z ~ O[a][s'], r = R[a][s]); actions in term_action_mask end the episode (opening a door in Tiger).
 * Random draws: Philox4x32-10, key = seed, counter = (step, epoch, 0, 5).  out_dev [n_epochs][4] =
 * {discounted return, undiscounted return, steps, last action}.  All pointers on the device. */
int vi_policy_rollout_dev(int device, int S, int A, int Z, int G, const double *alpha_dev, const int32_t *action_dev,
                          const double *T_dev, const double *O_dev, const double *R_dev, const double *b0_dev,
                          uint32_t term_action_mask, double discount, int max_steps, int n_epochs, uint64_t seed,
                          double *out_dev, void *stream, char *err, int errlen);

/* Belief-point expansion of point-based value iteration (stochastic simulation with a random action): for p < P,
 * beliefs[P + p] = update(beliefs[p], a, z) with a uniform, z ~ P(z | b_p, a) (counter = (round, p, 0, 6)).
 * beliefs_dev holds 2 * P rows of S doubles. */
int vi_expand_beliefs_dev(int device, int S, int A, int Z, int P, const double *T_dev, const double *O_dev,
                          double *beliefs_dev, uint32_t round, uint64_t seed, void *stream, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif
