/*
 * fftw_stub.c -- FFTW3-API stand-in used ONLY to build the CPU oracle (test infrastructure).
 *
 * See fftw3.h in this directory for the definitions being restated and the reference call sites.
 * Power-of-two lengths use an iterative radix-2 FFT whose twiddles are evaluated one by one with
 * libm (no recurrences, so the error stays at the few-ulp level); other lengths fall back to the
 * O(n^2) definition with an exact-index cosine/sine table.  DCT-II / DCT-III of a power-of-two
 * length go through one complex FFT of the same length (even/odd reordering).
 *
 * Not part of the product: nothing under s2kit_b200/ or include/ uses this file.
 */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { PLAN_R2R = 1, PLAN_SPLIT = 2 };

struct oracle_fftw_plan_s {
    int type;
    int n;
    fftw_r2r_kind kind;
    /* planned arrays (fftw_execute uses them) */
    double *in, *out, *ri, *ii, *ro, *io;
    /* split-dft geometry */
    int is, os, hn, his, hos;
    /* tables */
    int pow2;
    double* wr; /* cos(2 pi k / n), k < n */
    double* wi; /* sin(2 pi k / n), k < n */
    double* qr; /* cos(pi k / (2n)), k < 4n  (quarter-sample tables for the DCTs) */
    double* qi; /* sin(pi k / (2n)), k < 4n */
    int* rev;   /* bit reversal for pow2 */
    double *tr, *ti; /* scratch, length n */
};

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static void make_tables(fftw_plan p, int need_quarter) {
    int n = p->n;
    p->pow2 = is_pow2(n);
    p->wr = (double*)malloc(sizeof(double) * n);
    p->wi = (double*)malloc(sizeof(double) * n);
    for (int k = 0; k < n; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        p->wr[k] = cos(a);
        p->wi[k] = sin(a);
    }
    p->qr = p->qi = NULL;
    if (need_quarter) {
        p->qr = (double*)malloc(sizeof(double) * 4 * n);
        p->qi = (double*)malloc(sizeof(double) * 4 * n);
        for (int k = 0; k < 4 * n; ++k) {
            double a = M_PI * (double)k / (2.0 * (double)n);
            p->qr[k] = cos(a);
            p->qi[k] = sin(a);
        }
    }
    p->rev = NULL;
    if (p->pow2) {
        int bits = 0;
        while ((1 << bits) < n) ++bits;
        p->rev = (int*)malloc(sizeof(int) * n);
        for (int i = 0; i < n; ++i) {
            int r = 0;
            for (int b = 0; b < bits; ++b)
                if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            p->rev[i] = r;
        }
    }
    p->tr = (double*)malloc(sizeof(double) * n);
    p->ti = (double*)malloc(sizeof(double) * n);
}

/* in-place forward DFT (sign -1) of (xr,xi), length p->n, unit stride */
static void cfft_forward(fftw_plan p, double* xr, double* xi) {
    int n = p->n;
    if (n == 1) return;
    if (p->pow2) {
        for (int i = 0; i < n; ++i) {
            int r = p->rev[i];
            if (r > i) {
                double t = xr[i]; xr[i] = xr[r]; xr[r] = t;
                t = xi[i]; xi[i] = xi[r]; xi[r] = t;
            }
        }
        for (int len = 2; len <= n; len <<= 1) {
            int half = len >> 1, step = n / len;
            for (int base = 0; base < n; base += len) {
                for (int j = 0; j < half; ++j) {
                    double c = p->wr[j * step], s = -p->wi[j * step];
                    int a = base + j, b = a + half;
                    double ur = xr[b] * c - xi[b] * s;
                    double ui = xr[b] * s + xi[b] * c;
                    xr[b] = xr[a] - ur; xi[b] = xi[a] - ui;
                    xr[a] = xr[a] + ur; xi[a] = xi[a] + ui;
                }
            }
        }
        return;
    }
    /* definition, exact-index twiddles */
    double* yr = (double*)malloc(sizeof(double) * 2 * n);
    double* yi = yr + n;
    for (int k = 0; k < n; ++k) {
        double sr = 0.0, si = 0.0;
        for (int j = 0; j < n; ++j) {
            int idx = (int)(((long long)j * k) % n);
            double c = p->wr[idx], s = -p->wi[idx];
            sr += xr[j] * c - xi[j] * s;
            si += xr[j] * s + xi[j] * c;
        }
        yr[k] = sr; yi[k] = si;
    }
    memcpy(xr, yr, sizeof(double) * n);
    memcpy(xi, yi, sizeof(double) * n);
    free(yr);
}

static void dct2_unnormalised(fftw_plan p, const double* in, double* out) {
    int n = p->n;
    if (!p->pow2) {
        for (int k = 0; k < n; ++k) {
            double s = 0.0;
            for (int j = 0; j < n; ++j)
                s += in[j] * p->qr[(int)(((long long)(2 * j + 1) * k) % (4 * n))];
            out[k] = 2.0 * s;
        }
        return;
    }
    double *vr = p->tr, *vi = p->ti;
    for (int j = 0; 2 * j < n; ++j) vr[j] = in[2 * j];
    for (int j = 0; 2 * j + 1 < n; ++j) vr[n - 1 - j] = in[2 * j + 1];
    memset(vi, 0, sizeof(double) * n);
    cfft_forward(p, vr, vi);
    for (int k = 0; k < n; ++k) {
        /* 2 Re( exp(-i pi k / 2n) V[k] ) */
        out[k] = 2.0 * (p->qr[k] * vr[k] + p->qi[k] * vi[k]);
    }
}

static void dct3_unnormalised(fftw_plan p, const double* in, double* out) {
    int n = p->n;
    if (!p->pow2) {
        for (int k = 0; k < n; ++k) {
            double s = 0.0;
            for (int j = 1; j < n; ++j)
                s += in[j] * p->qr[(int)(((long long)(2 * k + 1) * j) % (4 * n))];
            out[k] = in[0] + 2.0 * s;
        }
        return;
    }
    /* W[j] = exp(i pi j / 2n) (X[j] - i X[n-j]), X[n] = 0; v = sum_j W[j] exp(+2 pi i j m / n) (real) */
    double *wr_ = p->tr, *wi_ = p->ti;
    for (int j = 0; j < n; ++j) {
        double a = in[j], b = (j == 0) ? 0.0 : in[n - j];
        double c = p->qr[j], s = p->qi[j];
        /* (c + i s)(a - i b) = (c a + s b) + i (s a - c b) */
        wr_[j] = c * a + s * b;
        wi_[j] = s * a - c * b;
    }
    /* inverse DFT through conjugation: conj(FFT(conj(W))) ; only the real part is needed */
    for (int j = 0; j < n; ++j) wi_[j] = -wi_[j];
    cfft_forward(p, wr_, wi_);
    for (int m = 0; 2 * m < n; ++m) out[2 * m] = wr_[m];
    for (int m = 0; 2 * m + 1 < n; ++m) out[2 * m + 1] = wr_[n - 1 - m];
}

fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags) {
    (void)flags;
    if (kind != FFTW_REDFT10 && kind != FFTW_REDFT01) return NULL;
    fftw_plan p = (fftw_plan)calloc(1, sizeof(*p));
    p->type = PLAN_R2R;
    p->n = n;
    p->kind = kind;
    p->in = in;
    p->out = out;
    make_tables(p, 1);
    return p;
}

fftw_plan fftw_plan_guru_split_dft(int rank, const fftw_iodim* dims, int howmany_rank,
                                   const fftw_iodim* howmany_dims, double* ri, double* ii, double* ro,
                                   double* io, unsigned flags) {
    (void)flags;
    if (rank != 1 || howmany_rank > 1) return NULL;
    fftw_plan p = (fftw_plan)calloc(1, sizeof(*p));
    p->type = PLAN_SPLIT;
    p->n = dims[0].n;
    p->is = dims[0].is;
    p->os = dims[0].os;
    if (howmany_rank == 1) {
        p->hn = howmany_dims[0].n;
        p->his = howmany_dims[0].is;
        p->hos = howmany_dims[0].os;
    } else {
        p->hn = 1;
        p->his = p->hos = 0;
    }
    p->ri = ri; p->ii = ii; p->ro = ro; p->io = io;
    make_tables(p, 0);
    return p;
}

void fftw_execute_r2r(const fftw_plan p, double* in, double* out) {
    if (p->kind == FFTW_REDFT10)
        dct2_unnormalised(p, in, out);
    else
        dct3_unnormalised(p, in, out);
}

void fftw_execute_split_dft(const fftw_plan p, double* ri, double* ii, double* ro, double* io) {
    int n = p->n;
    double* xr = (double*)malloc(sizeof(double) * 2 * n);
    double* xi = xr + n;
    for (int h = 0; h < p->hn; ++h) {
        const double* sr = ri + (long long)h * p->his;
        const double* si = ii + (long long)h * p->his;
        for (int j = 0; j < n; ++j) {
            xr[j] = sr[(long long)j * p->is];
            xi[j] = si[(long long)j * p->is];
        }
        cfft_forward(p, xr, xi);
        double* dr = ro + (long long)h * p->hos;
        double* di = io + (long long)h * p->hos;
        for (int k = 0; k < n; ++k) {
            dr[(long long)k * p->os] = xr[k];
            di[(long long)k * p->os] = xi[k];
        }
    }
    free(xr);
}

void fftw_execute(const fftw_plan p) {
    if (p->type == PLAN_R2R)
        fftw_execute_r2r(p, p->in, p->out);
    else
        fftw_execute_split_dft(p, p->ri, p->ii, p->ro, p->io);
}

void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    free(p->wr); free(p->wi); free(p->qr); free(p->qi); free(p->rev); free(p->tr); free(p->ti);
    free(p);
}
