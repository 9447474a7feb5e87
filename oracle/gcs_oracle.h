/*
 * gcs_oracle.h — CPU oracle for the batched Newton-Raphson path.  TEST INFRASTRUCTURE ONLY
 * (see the header of gcs_oracle.c).  It consumes the same batch descriptor as the product's
 * C ABI (include/gcs_b200.h) with mem == GCS_MEM_HOST, so a test can hand one batch to both.
 */
#ifndef GCS_ORACLE_H
#define GCS_ORACLE_H

#include "../include/gcs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* whole batch on the host; threads < 1 = all OpenMP threads, 1 = scalar */
int gcs_oracle_solve(const gcs_b200_batch* batch, int threads);

/* gcs_oracle_solve, plus per run (slack[2][n_seeds][n]) how far the literal trajectory's
 * convergence decisions stayed from the threshold in excess of half the margins the contracted
 * kernels claim: plane 0 against min(band[i], sdr[i] / |det J| + 2^-40 tol), plane 1 against the
 * second term alone (see newton2d_decision_slack).  Not re-entrant. */
int gcs_oracle_decision_slack(const gcs_b200_batch* batch, const double* sdr, const double* band, double* slack, int threads);

/* one Newton run (newton_raphson.hpp:53-99) of kind `kind` with the kind's 12 evaluation
 * constants (see `system2` in gcs_oracle.c) from guess (gx, gy) */
int gcs_oracle_newton2d(int kind, const double* consts, double gx, double gy, double* x, double* y,
    int* iters, int* converged);

/* the restated Eigen colPivHouseholderQr().solve for J = [a b; c d] given as {a,b,c,d} */
void gcs_oracle_qr_solve_2x2(const double J[4], const double rhs[2], double step[2]);

int gcs_oracle_kind_in_cols(int kind);
int gcs_oracle_kind_out_cols(int kind);
int gcs_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif

#endif
