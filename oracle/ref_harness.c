/*
 * ref_harness.c -- thin driver around the UNMODIFIED reference sources (test infrastructure).
 *
 * Built by oracle/Makefile into oracle/_ref/libs2kit_ref.so together with the reference's own
 * src/ files, which are compiled where they lie under /root/reference (never copied here) against the
 * FFTW-API stub in oracle/fftw_stub/.  The harness only packages what a caller of the reference has
 * to do anyway -- allocate workspaces, create the FFTW plans with the strides the reference expects
 * (test/test_s2_semi_memo.c:100-134), build the tables -- behind a few flat entry points that Python
 * (ctypes) can call, plus a multi-threaded timing loop for the CPU baseline.
 *
 * Known reference defects handled here: the table space gets 2*bw doubles of
 * slack because SemiNaive_Naive_Pml_Table with cutoff == bw writes past the end
 * (cospml.c:464-467 -> pml.c:66).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <fftw3.h>

#include "s2kit/FST_semi_fly.h"
#include "s2kit/FST_semi_memo.h"
#include "s2kit/cospml.h"
#include "s2kit/seminaive.h"
#include "s2kit/util.h"
#include "s2kit/weights.h"

typedef struct {
    fftw_plan dct, idct, fft, ifft;
    double* dummy;
} RefPlans;

typedef struct RefCtx {
    int bw, cutoff;
    long table_doubles;
    double *tablespace, *trans_tablespace;
    double **table, **trans_table;
    double* weights;
    double* workspace; /* big enough for Memo and Fly calls */
    RefPlans plans;
} RefCtx;

static void make_plans(int bw, RefPlans* p) {
    int n = 2 * bw;
    p->dummy = (double*)malloc(sizeof(double) * 4 * n);
    double* a = p->dummy;
    p->dct = fftw_plan_r2r_1d(n, a, a + n, FFTW_REDFT10, FFTW_ESTIMATE);
    p->idct = fftw_plan_r2r_1d(n, a, a + n, FFTW_REDFT01, FFTW_ESTIMATE);
    fftw_iodim d, h;
    /* forward: unit-stride rows in, transposed out */
    d.n = n; d.is = 1; d.os = n;
    h.n = n; h.is = n; h.os = 1;
    p->fft = fftw_plan_guru_split_dft(1, &d, 1, &h, a, a, a, a, FFTW_ESTIMATE);
    /* inverse: transposed in, unit-stride rows out */
    d.n = n; d.is = n; d.os = 1;
    h.n = n; h.is = 1; h.os = n;
    p->ifft = fftw_plan_guru_split_dft(1, &d, 1, &h, a, a, a, a, FFTW_ESTIMATE);
}

static void free_plans(RefPlans* p) {
    fftw_destroy_plan(p->dct);
    fftw_destroy_plan(p->idct);
    fftw_destroy_plan(p->fft);
    fftw_destroy_plan(p->ifft);
    free(p->dummy);
}

static size_t workspace_doubles(int bw) {
    /* max over FSTSemiMemo (8B^2+7B), InvFSTSemiMemo (8B^2+10B), Fly twins (10B^2+24B), FZT (13B+9B) */
    return (size_t)10 * bw * bw + (size_t)64 * bw + 64;
}

/* with_tables = 0 builds a context for the Fly entry points only */
RefCtx* ref_ctx_create(int bw, int cutoff, int with_tables) {
    RefCtx* c = (RefCtx*)calloc(1, sizeof(RefCtx));
    c->bw = bw;
    c->cutoff = cutoff;
    c->weights = (double*)malloc(sizeof(double) * 4 * bw);
    GenerateWeightsForDLT(bw, c->weights);
    c->workspace = (double*)malloc(sizeof(double) * workspace_doubles(bw));
    make_plans(bw, &c->plans);
    if (with_tables) {
        long sz = (long)Reduced_Naive_TableSize(bw, cutoff) + (long)Reduced_SpharmonicTableSize(bw, cutoff);
        c->table_doubles = sz;
        c->tablespace = (double*)calloc((size_t)sz + 2 * bw + 16, sizeof(double));
        c->trans_tablespace = (double*)calloc((size_t)sz + 2 * bw + 16, sizeof(double));
        c->table = SemiNaive_Naive_Pml_Table(bw, cutoff, c->tablespace, c->workspace);
        c->trans_table =
            Transpose_SemiNaive_Naive_Pml_Table(c->table, bw, cutoff, c->trans_tablespace, c->workspace);
    }
    return c;
}

void ref_ctx_destroy(RefCtx* c) {
    if (!c) return;
    free_plans(&c->plans);
    free(c->weights);
    free(c->workspace);
    free(c->table);
    free(c->trans_table);
    free(c->tablespace);
    free(c->trans_tablespace);
    free(c);
}

long ref_ctx_table_doubles(RefCtx* c) { return c->table_doubles; }
double* ref_ctx_table(RefCtx* c, int m) { return c->table[m]; }
double* ref_ctx_trans_table(RefCtx* c, int m) { return c->trans_table[m]; }
double* ref_ctx_weights(RefCtx* c) { return c->weights; }

void ref_fst_memo(RefCtx* c, double* rdata, double* idata, double* rcoeffs, double* icoeffs, int data_format) {
    FSTSemiMemo(rdata, idata, rcoeffs, icoeffs, c->bw, c->table, c->workspace, (DataFormat)data_format,
                c->cutoff, &c->plans.dct, &c->plans.fft, c->weights);
}

void ref_inv_fst_memo(RefCtx* c, double* rcoeffs, double* icoeffs, double* rdata, double* idata,
                      int data_format) {
    InvFSTSemiMemo(rcoeffs, icoeffs, rdata, idata, c->bw, c->trans_table, c->workspace,
                   (DataFormat)data_format, c->cutoff, &c->plans.idct, &c->plans.ifft);
}

void ref_fzt_memo(RefCtx* c, double* rdata, double* idata, double* rres, double* ires, int data_format) {
    /* FZTSemiMemo REAL zeroes 2*bw entries of ires (FST_semi_memo.c:406): callers pass 2*bw room */
    FZTSemiMemo(rdata, idata, rres, ires, c->bw, c->table[0], c->workspace, (DataFormat)data_format,
                &c->plans.dct, c->weights);
}

void ref_fst_fly(RefCtx* c, double* rdata, double* idata, double* rcoeffs, double* icoeffs, int data_format) {
    FSTSemiFly(rdata, idata, rcoeffs, icoeffs, c->bw, c->workspace, (DataFormat)data_format, c->cutoff,
               &c->plans.dct, &c->plans.fft, c->weights);
}

void ref_inv_fst_fly(RefCtx* c, double* rcoeffs, double* icoeffs, double* rdata, double* idata,
                     int data_format) {
    InvFSTSemiFly(rcoeffs, icoeffs, rdata, idata, c->bw, c->workspace, (DataFormat)data_format, c->cutoff,
                  &c->plans.idct, &c->plans.ifft);
}

void ref_fzt_fly(RefCtx* c, double* rdata, double* idata, double* rres, double* ires, int data_format) {
    FZTSemiFly(rdata, idata, rres, ires, c->bw, c->workspace, (DataFormat)data_format, &c->plans.dct,
               c->weights);
}

void ref_trans_mult(int bw, double* rd, double* id, double* rf, double* ifl, double* rres, double* ires) {
    TransMult(rd, id, rf, ifl, rres, ires, bw);
}

/* Convolution entry points own their workspace: 2*S(B) + 12B^2 + 12B (Memo, FST_semi_memo.c:446-455),
   14B^2 + 26B (Fly, FST_semi_fly.c:458-464); both get slack. */
void ref_conv_memo(int bw, double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
                   double* ires) {
    size_t legendre = (size_t)Reduced_Naive_TableSize(bw, bw) + (size_t)Reduced_SpharmonicTableSize(bw, bw);
    size_t need = 2 * legendre + (size_t)12 * bw * bw + (size_t)64 * bw + 64;
    double* ws = (double*)calloc(need, sizeof(double));
    ConvOn2SphereSemiMemo(rdata, idata, rfilter, ifilter, rres, ires, bw, ws);
    free(ws);
}

void ref_conv_fly(int bw, double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
                  double* ires) {
    size_t need = (size_t)16 * bw * bw + (size_t)64 * bw + 64;
    double* ws = (double*)calloc(need, sizeof(double));
    ConvOn2SphereSemiFly(rdata, idata, rfilter, ifilter, rres, ires, bw, ws);
    free(ws);
}

/* 1-D transforms for one order (seminaive.c:56,153) */
void ref_dlt_semi(RefCtx* c, double* data, int m, double* result) {
    DLTSemi(data, c->bw, m, result, c->workspace, c->table[m], c->weights, &c->plans.dct);
}

void ref_inv_dlt_semi(RefCtx* c, double* coeffs, int m, double* result) {
    int n = 2 * c->bw;
    double* sinv = (double*)malloc(sizeof(double) * n);
    for (int j = 0; j < n; ++j) sinv[j] = sin((2. * j + 1.) * M_PI / (2. * n));
    InvDLTSemi(coeffs, c->bw, m, result, c->trans_table[m], sinv, c->workspace, &c->plans.idct);
    free(sinv);
}

/* one order's packed cosine table straight from the reference generator (cospml.c:161) */
void ref_gen_cos_pml_table(int bw, int m, double* out) {
    double* ws = (double*)malloc(sizeof(double) * 16 * bw);
    GenerateCosPmlTable(bw, m, out, ws);
    free(ws);
}

int ref_table_size(int m, int bw) { return TableSize(m, bw); }
int ref_table_offset(int m, int l) { return TableOffset(m, l); }
int ref_index_of_coeff(int m, int l, int bw) { return IndexOfHarmonicCoeff(m, l, bw); }

/* Seeded coefficients of a real-valued band-limited field: the draw order and the symmetry of
   test/test_s2_semi_memo.c:156-172, with a fixed seed instead of time(). */
void ref_gen_coeffs(int bw, long seed, double* rc, double* ic) {
    srand48(seed);
    for (int m = 0; m < bw; ++m)
        for (int l = m; l < bw; ++l) {
            double x = 2.0 * (drand48() - 0.5);
            double y = 2.0 * (drand48() - 0.5);
            int ip = IndexOfHarmonicCoeff(m, l, bw);
            int in = IndexOfHarmonicCoeff(-m, l, bw);
            rc[ip] = x;
            ic[ip] = y;
            double sg = (m & 1) ? -1.0 : 1.0;
            rc[in] = sg * x;
            ic[in] = -sg * y;
        }
    for (int l = 0; l < bw; ++l) ic[l] = 0.0;
}

/* ---- CPU baseline timing: nthreads workers, each with private workspace/plans, shared tables ---- */
typedef struct {
    RefCtx* shared;
    int first, count, data_format, variant; /* variant 0 = Memo, 1 = Fly */
    long seed0;
    double seconds;
} Worker;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void* worker_main(void* arg) {
    Worker* w = (Worker*)arg;
    RefCtx* s = w->shared;
    int bw = s->bw, n = 2 * bw;
    double* ws = (double*)malloc(sizeof(double) * workspace_doubles(bw));
    double* rc = (double*)malloc(sizeof(double) * 4 * bw * bw);
    double* ic = rc + bw * bw;
    double* rr = ic + bw * bw;
    double* ir = rr + bw * bw;
    double* rd = (double*)malloc(sizeof(double) * 2 * n * n);
    double* id = rd + n * n;
    RefPlans pl;
    make_plans(bw, &pl);
    double acc = 0.0;
    for (int f = w->first; f < w->first + w->count; ++f) {
        ref_gen_coeffs(bw, w->seed0 + f, rc, ic);
        double t0 = now_s();
        if (w->variant == 0) {
            InvFSTSemiMemo(rc, ic, rd, id, bw, s->trans_table, ws, (DataFormat)w->data_format, s->cutoff,
                           &pl.idct, &pl.ifft);
            FSTSemiMemo(rd, id, rr, ir, bw, s->table, ws, (DataFormat)w->data_format, s->cutoff, &pl.dct,
                        &pl.fft, s->weights);
        } else {
            InvFSTSemiFly(rc, ic, rd, id, bw, ws, (DataFormat)w->data_format, s->cutoff, &pl.idct, &pl.ifft);
            FSTSemiFly(rd, id, rr, ir, bw, ws, (DataFormat)w->data_format, s->cutoff, &pl.dct, &pl.fft,
                       s->weights);
        }
        acc += now_s() - t0;
    }
    w->seconds = acc;
    free_plans(&pl);
    free(rd);
    free(rc);
    free(ws);
    return NULL;
}

/* Runs nfun inverse+forward pairs split over nthreads; returns wall seconds of the threaded region
   (tables excluded, as test_s2_semi_memo.c:174-190 times it). per_thread_busy (may be NULL) gets the
   summed per-thread busy time. */
double ref_bench_pairs(RefCtx* c, int nfun, int nthreads, long seed0, int data_format, int variant,
                       double* per_thread_busy) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nfun) nthreads = nfun;
    Worker* w = (Worker*)calloc(nthreads, sizeof(Worker));
    pthread_t* th = (pthread_t*)calloc(nthreads, sizeof(pthread_t));
    int base = nfun / nthreads, extra = nfun % nthreads, first = 0;
    for (int t = 0; t < nthreads; ++t) {
        w[t].shared = c;
        w[t].first = first;
        w[t].count = base + (t < extra ? 1 : 0);
        first += w[t].count;
        w[t].data_format = data_format;
        w[t].variant = variant;
        w[t].seed0 = seed0;
    }
    double t0 = now_s();
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker_main, &w[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    double wall = now_s() - t0;
    if (per_thread_busy) {
        double s = 0.0;
        for (int t = 0; t < nthreads; ++t) s += w[t].seconds;
        *per_thread_busy = s;
    }
    free(th);
    free(w);
    return wall;
}
