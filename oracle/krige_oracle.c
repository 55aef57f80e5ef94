/*
 * krige_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, never on the product path).
 *
 * Plain-C restatement of the reference's detrended-kriging weight solver, written from the
 * algorithm description, not copied.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this.
 *
 * Follows, in reference file:line terms (all under /root/reference/smrf/spatial/dk/):
 *   orc_krige_grid  <- krige.c:28-62    (loop over cells, weights row-major [ngrid][nsta])
 *   orc_krige_cell  <- krige.c:66-216   (active-set loop: build augmented system, solve, drop the
 *                                        highest-elevation station that has a negative weight)
 *   crout_factor    <- lusolv.c:57-157  (Crout LU, implicit row scaling, partial pivoting with >=,
 *                                        row swap limited to the first n columns)
 *   crout_solve     <- lusolv.c:167-202 (permute RHS through indx while forward substituting,
 *                                        "first non-zero" shortcut keyed on sum > 0.0, back-subst.)
 *
 * Must be compiled with -ffp-contract=off: the reference .so holds no FMA instructions, and the
 * parity target for the active set is bit-exact sign decisions.
 *
 * Parity pinning: tests/test_oracle.py checks this file bit-for-bit against oracle/_ref
 * (the reference's own krige.c/lusolv.c/array.c compiled in place) and against
 * tests/golden/dk_*.npz produced by the reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* LU factorisation of the leading n x n block of the row-major matrix a (row stride lda).
 * Returns 0, or 1 when a row is all zero / a pivot is exactly zero. */
static int crout_factor(double *a, int n, int lda, int *indx, double *vv)
{
    int i, j, k, imax = 0;
    for (i = 0; i < n; i++) {
        double big = 0.0;
        for (j = 0; j < n; j++) {
            double t = fabs(a[i * lda + j]);
            if (t > big) big = t;
        }
        if (big == 0.0) return 1;
        vv[i] = 1.0 / big;
    }
    for (j = 0; j < n; j++) {
        for (i = 0; i < j; i++) {
            double sum = a[i * lda + j];
            for (k = 0; k < i; k++) sum -= a[i * lda + k] * a[k * lda + j];
            a[i * lda + j] = sum;
        }
        double big = 0.0;
        for (i = j; i < n; i++) {
            double sum = a[i * lda + j];
            for (k = 0; k < j; k++) sum -= a[i * lda + k] * a[k * lda + j];
            a[i * lda + j] = sum;
            double dum = vv[i] * fabs(sum);
            if (dum >= big) { big = dum; imax = i; }   /* >= : the LAST maximal row wins */
        }
        if (j != imax) {
            for (k = 0; k < n; k++) {                   /* only the n matrix columns move */
                double t = a[imax * lda + k];
                a[imax * lda + k] = a[j * lda + k];
                a[j * lda + k] = t;
            }
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (a[j * lda + j] == 0.0) return 1;
        if (j != n - 1) {
            double inv = 1.0 / a[j * lda + j];
            for (i = j + 1; i < n; i++) a[i * lda + j] *= inv;
        }
    }
    return 0;
}

/* Solve using the factorisation; the right-hand side lives in column n of a. */
static void crout_solve(double *a, int n, int lda, const int *indx)
{
    int i, j, ii = -1;
    for (i = 0; i < n; i++) {
        int ip = indx[i];
        double sum = a[ip * lda + n];
        a[ip * lda + n] = a[i * lda + n];
        if (ii >= 0) {
            for (j = ii; j < i; j++) sum -= a[i * lda + j] * a[j * lda + n];
        } else if (sum > 0.0) {
            ii = i;
        }
        a[i * lda + n] = sum;
    }
    for (i = n - 1; i >= 0; i--) {
        double sum = a[i * lda + n];
        for (j = i + 1; j < n; j++) sum -= a[i * lda + j] * a[j * lda + n];
        a[i * lda + n] = sum / a[i * lda + i];
    }
}

/* One cell.  work: (nsta+1)*(nsta+2) + (nsta+1) doubles; iwork: 2*(nsta+1) ints.
 * active_out (may be NULL) receives the final 0/1 station flags.
 * Returns 0 ok, 1 singular system. */
int orc_krige_cell(int nsta, const double *ad, const double *dg, const double *elev,
                   double *w, int *active_out, double *work, int *iwork)
{
    const int lda = nsta + 2;
    double *a = work;
    double *vv = work + (size_t)(nsta + 1) * lda;
    int *flag = iwork;
    int *indx = iwork + (nsta + 1);
    int ns = nsta, m, n, mm, nn;

    for (m = 0; m < nsta; m++) flag[m] = 1;

    for (;;) {
        mm = -1;
        for (m = 0; m < nsta; m++) {
            if (!flag[m]) continue;
            mm++;
            nn = -1;
            for (n = 0; n < nsta; n++) {
                if (!flag[n]) continue;
                nn++;
                a[mm * lda + nn] = ad[m * nsta + n];
            }
            a[mm * lda + ns] = 1.0;
            a[ns * lda + mm] = 1.0;
            a[mm * lda + ns + 1] = dg[m];
        }
        a[ns * lda + ns] = 0.0;
        a[ns * lda + ns + 1] = 1.0;

        if (crout_factor(a, ns + 1, lda, indx, vv)) return 1;
        crout_solve(a, ns + 1, lda, indx);

        /* negative weight with the largest elevation (> 0, strict >, lowest index on ties) */
        double esave = 0.0;
        int msave = -1;
        mm = -1;
        for (m = 0; m < nsta; m++) {
            if (!flag[m]) continue;
            mm++;
            if (a[mm * lda + ns + 1] < 0.0 && elev[m] > esave) {
                msave = m;
                esave = elev[m];
            }
        }
        if (msave >= 0) {
            flag[msave] = 0;
            ns--;
            continue;
        }
        mm = -1;
        for (m = 0; m < nsta; m++) {
            if (flag[m]) { mm++; w[m] = a[mm * lda + ns + 1]; }
            else w[m] = 0.0;
        }
        break;
    }
    if (active_out) for (m = 0; m < nsta; m++) active_out[m] = flag[m];
    return 0;
}

/* Whole grid; same argument meaning/order as the reference's krige_grid (dk_header.h:23),
 * plus an optional per-cell active-flag output [ngrid][nsta] (may be NULL).
 * Returns the number of cells whose system was singular (the reference exit()s instead). */
int orc_krige_grid(int nsta, int ngrid, const double *ad, const double *dgrid,
                   const double *elev, int nthreads, double *weights, int *active)
{
    int nbad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : nbad)
    {
        double *work = (double *)malloc(sizeof(double) * ((size_t)(nsta + 1) * (nsta + 2) + nsta + 1));
        int *iwork = (int *)malloc(sizeof(int) * 2 * (nsta + 1));
#pragma omp for schedule(dynamic, 16)
        for (int c = 0; c < ngrid; c++) {
            nbad += orc_krige_cell(nsta, ad, dgrid + (size_t)c * nsta, elev,
                                   weights + (size_t)c * nsta,
                                   active ? active + (size_t)c * nsta : NULL, work, iwork);
        }
        free(work);
        free(iwork);
    }
    return nbad;
}

/* Solve one (n x n | rhs) augmented system exactly as the reference's lusolv() would:
 * a is row-major n x (n+1); x receives the n solutions.  Used to pin crout_* on their own. */
int orc_lusolv(int n, double *a, double *x)
{
    double *vv = (double *)malloc(sizeof(double) * n);
    int *indx = (int *)malloc(sizeof(int) * n);
    int rc = crout_factor(a, n, n + 1, indx, vv);
    if (!rc) {
        crout_solve(a, n, n + 1, indx);
        for (int i = 0; i < n; i++) x[i] = a[i * (n + 1) + n];
    }
    free(vv);
    free(indx);
    return rc;
}
