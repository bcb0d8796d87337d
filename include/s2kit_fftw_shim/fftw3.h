/*
 * fftw3.h (shim) -- for S2kit callers that are relinked against libs2kit_cuda.so on a machine without FFTW.
 *
 * The reference's public headers include <fftw3.h> and its callers create FFTW plans that they pass, by address,
 * into FSTSemiMemo & co. (include/s2kit/FST_semi_memo.h:4, test/test_s2_semi_memo.c:100-134).  The GPU engine never
 * reads those plans, so a caller without FFTW only needs the TYPES and the plan create/destroy entry points to
 * compile and link.  libs2kit_fftw_shim.so provides them as descriptor stubs; executing a shim plan aborts.
 * Callers that do have FFTW keep using the real header and library.
 */
#ifndef S2KIT_FFTW3_SHIM_H
#define S2KIT_FFTW3_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct s2kit_fftw_shim_plan* fftw_plan;
typedef struct { int n, is, os; } fftw_iodim;
typedef enum { FFTW_R2HC = 0, FFTW_HC2R, FFTW_DHT, FFTW_REDFT00, FFTW_REDFT01, FFTW_REDFT10, FFTW_REDFT11 } fftw_r2r_kind;
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags);
fftw_plan fftw_plan_guru_split_dft(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims,
                                   double* ri, double* ii, double* ro, double* io, unsigned flags);
void fftw_destroy_plan(fftw_plan p);
void fftw_execute(const fftw_plan p);
void fftw_execute_r2r(const fftw_plan p, double* in, double* out);
void fftw_execute_split_dft(const fftw_plan p, double* ri, double* ii, double* ro, double* io);

#ifdef __cplusplus
}
#endif
#endif
