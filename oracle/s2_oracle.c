/*
 * s2_oracle.c -- independent CPU restatement of the S2kit seminaive spherical harmonic transform.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product path (s2kit_b200/, include/) never does.
 *
 * What it restates (reference = Bychin/S2kit 1.1, paths relative to /root/reference):
 *   quadrature weights                src/legendre_transform/weights.c:32-47
 *   Chebyshev nodes / angles          src/util/chebyshev_nodes.c:16-34
 *   P_m^m seed                        src/legendre_polynomials/pmm.c:21-33
 *   recurrence coefficients           src/legendre_polynomials/util/l2_norms.c:16-38
 *   packed cosine-series table        src/legendre_polynomials/cospml.c:39-59,123-134,161-258
 *   forward / inverse Legendre        src/legendre_transform/seminaive.c:56-115,153-198
 *   forward / inverse / zonal / conv  src/FST_semi_memo.c:68-202,228-351,374-407,439-518
 *   coefficient order, spectral mult  src/util/util.c:23-103
 * The longitude FFT and the DCTs come from FFTW in the reference (third-party, un-vendored, "version 3",
 * not installed here); their published definitions are restated in fftw_stub/ and used through the
 * same API.
 *
 * It is written from the matrix formulation, not transliterated: one loop over
 * signed orders serves both halves of the spectrum, the inverse accumulates T^T c by scattering table
 * rows instead of reading a transposed copy, and pure seminaive (cutoff == bw) is the only mode.  The
 * floating-point operation ORDER of everything that feeds the tables (nodes, seed, recurrence, DCT
 * scaling) follows the reference exactly, because outputs at bw >= 1024 are sensitive to it.
 *
 * Pinning: checked in tests/test_oracle.py against the reference's four golden convolution files,
 * its four Y_l^m known-answer grids, and against oracle/_ref (the reference itself compiled here).
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fftw3.h"

/* ------------------------------------------------------------------ layout arithmetic */

/* number of stored cosine coefficients of degree l in order m's table (cospml.c:250-258) */
int orc_row_len(int m, int l) {
    if (l < m) return 0;
    return (m & 1) ? (l - 1) / 2 + 1 : l / 2 + 1;
}

/* start of degree l inside order m's packed table (cospml.c:123-134): sum of the earlier rows */
int orc_row_start(int m, int l) {
    int s = 0;
    for (int d = m; d < l; ++d) s += orc_row_len(m, d);
    return s;
}

/* doubles in order m's packed table (cospml.c:39-59) */
int orc_order_len(int m, int bw) { return orc_row_start(m, bw); }

/* all orders 0..bw-1 (cospml.c:107-115) */
long orc_total_len(int bw) {
    long s = 0;
    for (int m = 0; m < bw; ++m) s += orc_order_len(m, bw);
    return s;
}

/* position of f^(m,l) in the coefficient arrays (util.c:42-49) */
int orc_coef_index(int m, int l, int bw) {
    if (m >= 0) return m * bw - (m * (m - 1)) / 2 + (l - m);
    int a = -m;
    /* orders -(bw-1) .. -1 follow the bw(bw+1)/2 non-negative-order entries */
    int before = bw * (bw + 1) / 2;
    for (int o = bw - 1; o > a; --o) before += bw - o;
    return before + (l - a);
}

/* ------------------------------------------------------------------ setup quantities */

/* weights.c:32-47: w[j], j < 2bw for even orders, w[2bw + j] = w[j] sin(theta_j) for odd orders */
void orc_weights(int bw, double* w) {
    double step = M_PI / (4. * bw);
    for (int j = 0; j < 2 * bw; ++j) {
        double odd = 2. * j + 1.;
        double acc = 0.;
        for (int k = 0; k < bw; ++k) acc += 1. / (2. * k + 1.) * sin(odd * (2. * k + 1.) * step);
        acc *= 2. * sin(odd * step) / bw;
        w[j] = acc;
        w[j + 2 * bw] = acc * sin(odd * step);
    }
}

/* l2_norms.c:16-24 */
static double rec_a(int m, int l) {
    return sqrt(((2. * l + 3.) / (2. * l + 1.)) * ((l - m + 1.) / (l + m + 1.))) * ((2. * l + 1.) / (l - m + 1.));
}

/* l2_norms.c:28-38 */
static double rec_c(int m, int l) {
    if (l == 0) return 0.;
    return -1.0 *
           sqrt(((2. * l + 3.) / (2. * l - 1.)) * ((l - m + 1.) / (l + m + 1.)) *
                (((double)l - m) / ((double)l + m))) *
           ((l + m) / (l - m + 1.));
}

/* pmm.c:21-33 (normalisation constant only) */
static double seed_norm(int m) {
    double c = sqrt(m + 0.5);
    for (int i = 0; i < m; ++i) c *= sqrt((m - (i / 2.)) / ((double)m - i));
    if (m) c *= pow(2., -m / 2.);
    if (m & 1) c *= -1.;
    return c;
}

/* cospml.c:161-242: packed cosine series of P~_l^m (m even) or P~_l^m / sin (m odd), l = m..bw-1,
   sampled at the bw-point Chebyshev grid. */
void orc_cos_table(int bw, int m, double* out) {
    double* buf = (double*)malloc(sizeof(double) * 6 * bw);
    double *x = buf, *th = x + bw, *older = th + bw, *cur = older + bw, *next = cur + bw, *cs = next + bw;
    double den = 2. * bw;
    for (int i = 0; i < bw; ++i) {
        th[i] = (2. * i + 1.) * M_PI / den;       /* chebyshev_nodes.c:16-21 */
        x[i] = cos((2. * i + 1.) * M_PI / den);   /* chebyshev_nodes.c:29-34 */
        older[i] = 0.;
    }
    if (m == 0) {
        for (int i = 0; i < bw; ++i) cur[i] = M_SQRT1_2;
    } else {
        double c = seed_norm(m);
        for (int i = 0; i < bw; ++i) cur[i] = c * pow(sin(th[i]), m);
    }
    if (m & 1)
        for (int i = 0; i < bw; ++i) cur[i] /= sin(th[i]);

    fftw_plan dct = fftw_plan_r2r_1d(bw, cur, cs, FFTW_REDFT10, FFTW_ESTIMATE);
    double inv_root = 1. / sqrt(bw);
    long pos = 0;
    for (int l = m; l < bw; ++l) {
        fftw_execute_r2r(dct, cur, cs);
        cs[0] *= M_SQRT1_2;
        for (int k = 0; k < bw; ++k) cs[k] *= inv_root;
        int par = (l - m) & 1, len = orc_row_len(m, l);
        for (int q = 0; q < len; ++q) out[pos++] = cs[2 * q + par];
        if (l + 1 == bw) break;
        double a = rec_a(m, l), c = rec_c(m, l);
        for (int i = 0; i < bw; ++i) {
            /* cospml.c:218-221 order: c*older ; cur*x ; a*(cur*x) ; sum */
            double t1 = c * older[i];
            double t2 = cur[i] * x[i];
            double t3 = a * t2;
            next[i] = t3 + t1;
        }
        memcpy(older, cur, sizeof(double) * bw);
        memcpy(cur, next, sizeof(double) * bw);
    }
    fftw_destroy_plan(dct);
    free(buf);
}

/* ------------------------------------------------------------------ context */

typedef struct Oracle {
    int bw;
    double* weights;  /* 4 bw */
    double* sines;    /* 2 bw: sin((2j+1) pi / 4bw), FST_semi_memo.c:244-246 */
    double** table;   /* bw packed tables */
    double* tablespace;
    fftw_plan dct2, dct3, rows_to_orders, orders_to_rows;
    double *fr, *fi;  /* 2bw x 2bw spectral planes, order-major */
    double *col, *cosv;
} Oracle;

Oracle* orc_create(int bw) {
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    int n = 2 * bw;
    o->bw = bw;
    o->weights = (double*)malloc(sizeof(double) * 4 * bw);
    orc_weights(bw, o->weights);
    o->sines = (double*)malloc(sizeof(double) * n);
    for (int j = 0; j < n; ++j) o->sines[j] = sin((2. * j + 1.) * M_PI / (2. * n));
    o->tablespace = (double*)malloc(sizeof(double) * (size_t)orc_total_len(bw));
    o->table = (double**)malloc(sizeof(double*) * bw);
    double* p = o->tablespace;
    for (int m = 0; m < bw; ++m) {
        o->table[m] = p;
        orc_cos_table(bw, m, p);
        p += orc_order_len(m, bw);
    }
    o->fr = (double*)malloc(sizeof(double) * 2 * n * n);
    o->fi = o->fr + (size_t)n * n;
    o->col = (double*)malloc(sizeof(double) * 2 * n);
    o->cosv = o->col + n;
    o->dct2 = fftw_plan_r2r_1d(n, o->col, o->cosv, FFTW_REDFT10, FFTW_ESTIMATE);
    o->dct3 = fftw_plan_r2r_1d(n, o->col, o->cosv, FFTW_REDFT01, FFTW_ESTIMATE);
    fftw_iodim d, h;
    d.n = n; d.is = 1; d.os = n; h.n = n; h.is = n; h.os = 1;
    o->rows_to_orders = fftw_plan_guru_split_dft(1, &d, 1, &h, o->fr, o->fi, o->fr, o->fi, FFTW_ESTIMATE);
    d.n = n; d.is = n; d.os = 1; h.n = n; h.is = 1; h.os = n;
    o->orders_to_rows = fftw_plan_guru_split_dft(1, &d, 1, &h, o->fr, o->fi, o->fr, o->fi, FFTW_ESTIMATE);
    return o;
}

void orc_destroy(Oracle* o) {
    if (!o) return;
    fftw_destroy_plan(o->dct2); fftw_destroy_plan(o->dct3);
    fftw_destroy_plan(o->rows_to_orders); fftw_destroy_plan(o->orders_to_rows);
    free(o->col); free(o->fr); free(o->table); free(o->tablespace); free(o->sines); free(o->weights);
    free(o);
}

double* orc_table(Oracle* o, int m) { return o->table[m]; }
double* orc_weights_ptr(Oracle* o) { return o->weights; }

/* ------------------------------------------------------------------ one order, one real column */

/* seminaive.c:153-198: weights, orthonormal DCT-II(2bw), triangular product with the packed table */
static void legendre_forward(Oracle* o, const double* samples, int m, double sign, double* out) {
    int bw = o->bw, n = 2 * bw;
    const double* w = o->weights + ((m & 1) ? n : 0);
    for (int j = 0; j < n; ++j) o->col[j] = samples[j] * w[j];
    fftw_execute_r2r(o->dct2, o->col, o->cosv);
    o->cosv[0] *= M_SQRT1_2;
    double s = 1. / sqrt(2. * n);
    for (int k = 0; k < n; ++k) o->cosv[k] *= s;
    const double* row = o->table[m];
    for (int l = m; l < bw; ++l) {
        int par = (l - m) & 1, len = orc_row_len(m, l);
        double acc = 0.;
        for (int q = 0; q < len; ++q) acc += o->cosv[2 * q + par] * row[q];
        out[l - m] = sign * acc;
        row += len;
    }
}

/* seminaive.c:56-115: cosine series v = T^T c, orthonormal DCT-III(2bw), times sin(theta) if m odd */
static void legendre_inverse(Oracle* o, const double* coeffs, int m, double scale, double* out) {
    int bw = o->bw, n = 2 * bw;
    double* v = o->col;
    memset(v, 0, sizeof(double) * n);
    const double* row = o->table[m];
    for (int l = m; l < bw; ++l) {
        int par = (l - m) & 1, len = orc_row_len(m, l);
        double c = coeffs[l - m];
        for (int q = 0; q < len; ++q) v[2 * q + par] += row[q] * c;
        row += len;
    }
    double half = 0.5 / sqrt(bw);
    double v0 = v[0];
    for (int k = 0; k < bw; ++k) v[k] *= half;
    v[0] = v0 / sqrt((double)n);
    fftw_execute_r2r(o->dct3, v, o->cosv);
    if (m & 1)
        for (int j = 0; j < n; ++j) out[j] = scale * (o->cosv[j] * o->sines[j]);
    else
        for (int j = 0; j < n; ++j) out[j] = scale * o->cosv[j];
}

/* ------------------------------------------------------------------ 2-D transforms */

/* FST_semi_memo.c:68-202.  data_format: 0 = COMPLEX, 1 = REAL (include/s2kit/util.h:10-13) */
void orc_forward(Oracle* o, double* rdata, double* idata, double* rco, double* ico, int data_format) {
    int bw = o->bw, n = 2 * bw;
    fftw_execute_split_dft(o->rows_to_orders, rdata, idata, o->fr, o->fi);
    double norm = sqrt(2. * M_PI) / n;
    for (long i = 0; i < (long)n * n; ++i) {
        o->fr[i] *= norm;
        o->fi[i] *= norm;
    }
    for (int m = 0; m < bw; ++m) {
        int at = orc_coef_index(m, m, bw);
        legendre_forward(o, o->fr + (long)m * n, m, 1., rco + at);
        legendre_forward(o, o->fi + (long)m * n, m, 1., ico + at);
    }
    for (int m = 1; m < bw; ++m) {
        int at = orc_coef_index(-m, m, bw);
        double sg = (m & 1) ? -1. : 1.;
        if (data_format == 1) {
            /* f^(-m,l) = (-1)^m conj f^(m,l)  (FST_semi_memo.c:131-145) */
            int src = orc_coef_index(m, m, bw);
            for (int l = m; l < bw; ++l) {
                rco[at + l - m] = sg * rco[src + l - m];
                ico[at + l - m] = -sg * ico[src + l - m];
            }
        } else {
            /* spectral row 2bw - m carries order -m (FST_semi_memo.c:153-201) */
            legendre_forward(o, o->fr + (long)(n - m) * n, m, sg, rco + at);
            legendre_forward(o, o->fi + (long)(n - m) * n, m, sg, ico + at);
        }
    }
}

/* FST_semi_memo.c:228-351 */
void orc_inverse(Oracle* o, double* rco, double* ico, double* rdata, double* idata, int data_format) {
    int bw = o->bw, n = 2 * bw;
    for (int m = 0; m < bw; ++m) {
        int at = orc_coef_index(m, m, bw);
        legendre_inverse(o, rco + at, m, 1., o->fr + (long)m * n);
        legendre_inverse(o, ico + at, m, 1., o->fi + (long)m * n);
    }
    memset(o->fr + (long)bw * n, 0, sizeof(double) * n);
    memset(o->fi + (long)bw * n, 0, sizeof(double) * n);
    for (int m = 1; m < bw; ++m) {
        double* dr = o->fr + (long)(n - m) * n;
        double* di = o->fi + (long)(n - m) * n;
        if (data_format == 1) {
            /* conjugate mirror; negative-order inputs are ignored (FST_semi_memo.c:333-341) */
            for (int j = 0; j < n; ++j) {
                dr[j] = o->fr[(long)m * n + j];
                di[j] = -o->fi[(long)m * n + j];
            }
        } else {
            int at = orc_coef_index(-m, m, bw);
            double sg = (m & 1) ? -1. : 1.;
            legendre_inverse(o, rco + at, m, sg, dr);
            legendre_inverse(o, ico + at, m, sg, di);
        }
    }
    double norm = 1. / sqrt(2. * M_PI);
    for (long i = 0; i < (long)n * n; ++i) {
        o->fr[i] *= norm;
        o->fi[i] *= norm;
    }
    /* swapping re/im on both sides turns the forward-sign DFT into the inverse (FST_semi_memo.c:350) */
    fftw_execute_split_dft(o->orders_to_rows, o->fi, o->fr, idata, rdata);
}

/* FST_semi_memo.c:374-407: order-0 transform from row sums.  ires gets bw zeros in REAL format (the
   reference clears 2bw, a documented overrun). */
void orc_zonal(Oracle* o, double* rdata, double* idata, double* rres, double* ires, int data_format) {
    int bw = o->bw, n = 2 * bw;
    double norm = sqrt(2. * M_PI) / n;
    double* r0 = (double*)malloc(sizeof(double) * 2 * n);
    double* i0 = r0 + n;
    for (int j = 0; j < n; ++j) {
        double sr = 0., si = 0.;
        for (int k = 0; k < n; ++k) {
            sr += rdata[(long)j * n + k];
            si += idata[(long)j * n + k];
        }
        r0[j] = sr * norm;
        i0[j] = si * norm;
    }
    legendre_forward(o, r0, 0, 1., rres);
    if (data_format == 0)
        legendre_forward(o, i0, 0, 1., ires);
    else
        memset(ires, 0, sizeof(double) * bw);
    free(r0);
}

/* util.c:68-103 with ComplexMult's signs as written (util.c:23-27): im = x*v - y*u */
void orc_spectral_multiply(int bw, const double* rd, const double* id, const double* rf, const double* ifl,
                           double* rres, double* ires) {
    for (int m = -(bw - 1); m < bw; ++m) {
        int a = m < 0 ? -m : m;
        int at = orc_coef_index(m, a, bw);
        for (int l = a; l < bw; ++l) {
            double x = rf[l], y = ifl[l], u = rd[at + l - a], v = id[at + l - a];
            double re = x * u - y * v;
            double im = x * v - y * u;
            double s = sqrt(4. * M_PI / (2. * l + 1.));
            rres[at + l - a] = re * s;
            ires[at + l - a] = im * s;
        }
    }
}

/* FST_semi_memo.c:439-518: forward(REAL) -> zonal(REAL) -> multiply -> inverse(REAL) */
void orc_conv(Oracle* o, double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
              double* ires) {
    int bw = o->bw;
    double* c = (double*)malloc(sizeof(double) * ((size_t)4 * bw * bw + 2 * bw));
    double *fr = c, *fi = fr + bw * bw, *tr = fi + bw * bw, *ti = tr + bw * bw, *hr = ti + bw * bw, *hi = hr + bw;
    orc_forward(o, rdata, idata, fr, fi, 1);
    orc_zonal(o, rfilter, ifilter, hr, hi, 1);
    orc_spectral_multiply(bw, fr, fi, hr, hi, tr, ti);
    orc_inverse(o, tr, ti, rres, ires, 1);
    free(c);
}

/* ------------------------------------------------------------------ seeded test input */

/* drand48 restated: X <- (0x5DEECE66D X + 0xB) mod 2^48, value X / 2^48,
   srand48(s): X = (s mod 2^32) << 16 | 0x330E.  Draw order and symmetry of
   test/test_s2_semi_memo.c:156-172. */
void orc_gen_coeffs(int bw, long seed, double* rc, double* ic) {
    unsigned long long X = (((unsigned long long)seed & 0xFFFFFFFFULL) << 16) | 0x330EULL;
    const unsigned long long A = 0x5DEECE66DULL, C = 0xBULL, MASK = (1ULL << 48) - 1;
    for (int m = 0; m < bw; ++m)
        for (int l = m; l < bw; ++l) {
            X = (A * X + C) & MASK;
            double x = 2.0 * ((double)X / 281474976710656.0 - 0.5);
            X = (A * X + C) & MASK;
            double y = 2.0 * ((double)X / 281474976710656.0 - 0.5);
            int ip = orc_coef_index(m, l, bw), in = orc_coef_index(-m, l, bw);
            double sg = (m & 1) ? -1. : 1.;
            rc[ip] = x;
            ic[ip] = y;
            rc[in] = sg * x;
            ic[in] = -sg * y;
        }
    for (int l = 0; l < bw; ++l) ic[l] = 0.;
}
