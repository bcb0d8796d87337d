/*
 * fftw3.h -- minimal FFTW3-API stand-in for the ORACLE build only (test infrastructure).
 *
 * FFTW is an un-vendored third-party dependency of the reference (Makefile:3-5, "requires FFTW
 * version 3") and is not installed in this image.  This header + fftw_stub.c restate the six FFTW
 * entry points the reference calls, from FFTW's published definitions:
 *   REDFT10:  Y[k] = 2 * sum_j X[j] cos(pi (j+1/2) k / n)
 *   REDFT01:  Y[k] = X[0] + 2 * sum_{j>=1} X[j] cos(pi j (k+1/2) / n)
 *   split DFT: Y[k] = sum_j X[j] exp(-2 pi i j k / n), unnormalised, fftw_iodim strides
 * Call sites in the reference: src/FST_semi_memo.c:81,350,461-495, src/FST_semi_fly.c:81,368,470-504,
 * src/legendre_transform/seminaive.c:107,170, src/legendre_polynomials/cospml.c:203,204,224,241.
 *
 * Nothing in the product (s2kit_b200/, include/) includes or links this file.
 */
#ifndef ORACLE_FFTW3_STUB_H
#define ORACLE_FFTW3_STUB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_fftw_plan_s* fftw_plan;

typedef struct {
    int n;
    int is;
    int os;
} fftw_iodim;

typedef enum {
    FFTW_R2HC = 0,
    FFTW_HC2R = 1,
    FFTW_DHT = 2,
    FFTW_REDFT00 = 3,
    FFTW_REDFT01 = 4,
    FFTW_REDFT10 = 5,
    FFTW_REDFT11 = 6
} fftw_r2r_kind;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags);
fftw_plan fftw_plan_guru_split_dft(int rank, const fftw_iodim* dims, int howmany_rank,
                                   const fftw_iodim* howmany_dims, double* ri, double* ii, double* ro,
                                   double* io, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_execute_r2r(const fftw_plan p, double* in, double* out);
void fftw_execute_split_dft(const fftw_plan p, double* ri, double* ii, double* ro, double* io);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
