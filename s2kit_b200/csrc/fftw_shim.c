/* fftw_shim.c -- descriptor stubs of the FFTW entry points S2kit callers use to CREATE plans (see
 * include/s2kit_fftw_shim/fftw3.h).  Built into libs2kit_fftw_shim.so, separate from libs2kit_cuda.so so that callers
 * with a real FFTW never see these symbols.  No transform is implemented here: the GPU engine ignores the plans. */
#include <stdio.h>
#include <stdlib.h>

#include "../../include/s2kit_fftw_shim/fftw3.h"

struct s2kit_fftw_shim_plan {
    int kind, n;
};

fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags) {
    (void)in; (void)out; (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->kind = (int)kind;
    p->n = n;
    return p;
}

fftw_plan fftw_plan_guru_split_dft(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims,
                                   double* ri, double* ii, double* ro, double* io, unsigned flags) {
    (void)rank; (void)howmany_rank; (void)howmany_dims; (void)ri; (void)ii; (void)ro; (void)io; (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(*p));
    p->kind = -1;
    p->n = dims ? dims[0].n : 0;
    return p;
}

void fftw_destroy_plan(fftw_plan p) { free(p); }

static void no_execute(void) {
    fprintf(stderr, "s2kit fftw shim: plans are descriptors only and cannot be executed (link a real FFTW)\n");
    abort();
}
void fftw_execute(const fftw_plan p) { (void)p; no_execute(); }
void fftw_execute_r2r(const fftw_plan p, double* in, double* out) { (void)p; (void)in; (void)out; no_execute(); }
void fftw_execute_split_dft(const fftw_plan p, double* ri, double* ii, double* ro, double* io) {
    (void)p; (void)ri; (void)ii; (void)ro; (void)io; no_execute();
}
