/*
 * s2kit_compat.c -- the C host layer: S2kit's public C API on top of the s2kit_cuda_* C-ABI.
 *
 * Every function here has the name, argument order and meaning of the reference's declaration so that a C
 * caller of S2kit can relink against libs2kit_cuda.so:
 *   include/s2kit/FST_semi_memo.h:8-16   FSTSemiMemo InvFSTSemiMemo FZTSemiMemo ConvOn2SphereSemiMemo
 *   include/s2kit/FST_semi_fly.h:8-16    FSTSemiFly  InvFSTSemiFly  FZTSemiFly  ConvOn2SphereSemiFly
 *   include/s2kit/seminaive.h:6-8        DLTSemi InvDLTSemi
 *   include/s2kit/naive.h:4-6            DLTNaive InvDLTNaive
 *   include/s2kit/pmm.h:4                Pmm_L2
 *   include/s2kit/cospml.h:6-30          TableSize ... Transpose_SemiNaive_Naive_Pml_Table
 *   include/s2kit/weights.h:4            GenerateWeightsForDLT
 *   include/s2kit/util.h:15-17           IndexOfHarmonicCoeff TransMult
 * The transforms run on the GPU through a per-bandwidth plan cache; the `workspace`, FFTW-plan and host
 * table arguments are accepted and ignored (device tables are keyed by bandwidth), and `cutoff` no longer
 * switches algorithms: every order uses the seminaive algorithm (the reference's hybrid differs from its
 * own pure-seminaive result by <= 1.3e-13 at bw = 512).  The reference's functions
 * return void and never report errors; here an unrecoverable CUDA failure prints the reason and aborts --
 * there is no CPU fallback.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/s2kit.h"
#include "../../include/s2kit_cuda.h"
#include "host_setup.h"

/* ------------------------------------------------------------------------------------ plan cache */
/* The reference's functions are re-entrant when every thread brings its own workspace (that is how a threaded caller
   drives them, e.g. oracle/ref_harness.c:225-262).  Here a cache entry per (bandwidth, variant) holds ONE set of device
   tables and up to MAX_CTX contexts -- the plan that owns the tables plus clones that share them
   (s2kit_cuda_plan_clone), each with its own stream and workspaces.  A call checks a free context out, runs, and
   checks it back in; concurrent callers therefore never share a workspace, and an entry is only evicted when nobody
   is inside it. */
#define MAX_CACHED 8
#define MAX_CTX 8
typedef struct {
    int bw, variant;
    s2kit_cuda_plan* ctx[MAX_CTX]; /* ctx[0] owns the tables */
    int busy[MAX_CTX];
    int nctx, users;
    unsigned long stamp;
} CacheEntry;
static CacheEntry g_cache[MAX_CACHED]; /* nctx == 0: free slot */
static unsigned long g_clock = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;

static void die(const char* where) {
    fprintf(stderr, "s2kit_cuda: %s failed: %s\n", where, s2kit_cuda_last_error());
    abort();
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

static void entry_destroy(CacheEntry* e) {
    for (int i = e->nctx - 1; i >= 0; --i) s2kit_cuda_plan_destroy(e->ctx[i]); /* clones first, the owner last */
    e->nctx = 0;
}

typedef struct {
    CacheEntry* entry;
    int slot;
    s2kit_cuda_plan* plan;
} Checkout;

static Checkout ctx_acquire(int bw, int variant) {
    Checkout c = {NULL, 0, NULL};
    int max_ctx = env_int("S2KIT_CUDA_CONTEXTS", 4);
    if (max_ctx < 1) max_ctx = 1;
    if (max_ctx > MAX_CTX) max_ctx = MAX_CTX;
    pthread_mutex_lock(&g_lock);
    for (;;) {
        CacheEntry* e = NULL;
        for (int i = 0; i < MAX_CACHED; ++i) /* entries never move: checked-out callers hold pointers to them */
            if (g_cache[i].nctx && g_cache[i].bw == bw && g_cache[i].variant == variant) e = &g_cache[i];
        if (!e) {
            int at = -1;
            for (int i = 0; i < MAX_CACHED && at < 0; ++i)
                if (!g_cache[i].nctx) at = i;
            if (at < 0) {
                for (int i = 0; i < MAX_CACHED; ++i)
                    if (g_cache[i].users == 0 && (at < 0 || g_cache[i].stamp < g_cache[at].stamp)) at = i;
                if (at < 0) { /* every cached plan is in use: wait for one to come back */
                    pthread_cond_wait(&g_cv, &g_lock);
                    continue;
                }
                entry_destroy(&g_cache[at]);
            }
            e = &g_cache[at];
            memset(e, 0, sizeof(*e));
            e->bw = bw;
            e->variant = variant;
            if (s2kit_cuda_plan_create(&e->ctx[0], bw, variant, 1, env_int("S2KIT_CUDA_DEVICE", 0))) {
                pthread_mutex_unlock(&g_lock);
                die("plan_create");
            }
            e->nctx = 1;
        }
        int slot = -1;
        for (int i = 0; i < e->nctx; ++i)
            if (!e->busy[i]) {
                slot = i;
                break;
            }
        if (slot < 0 && e->nctx < max_ctx) {
            if (s2kit_cuda_plan_clone(&e->ctx[e->nctx], e->ctx[0], 1)) {
                pthread_mutex_unlock(&g_lock);
                die("plan_clone");
            }
            slot = e->nctx++;
        }
        if (slot < 0) {
            pthread_cond_wait(&g_cv, &g_lock);
            continue;
        }
        e->busy[slot] = 1;
        e->users++;
        e->stamp = ++g_clock;
        c.entry = e;
        c.slot = slot;
        c.plan = e->ctx[slot];
        break;
    }
    pthread_mutex_unlock(&g_lock);
    return c;
}

static void ctx_release(Checkout c) {
    pthread_mutex_lock(&g_lock);
    c.entry->busy[c.slot] = 0;
    c.entry->users--;
    pthread_cond_broadcast(&g_cv);
    pthread_mutex_unlock(&g_lock);
}

/* ---- single large field on several GPUs: S2KIT_CUDA_NGPU > 1 routes COMPLEX-format FSTSemiMemo / InvFSTSemiMemo calls
   at bw >= S2KIT_CUDA_MULTI_MIN_BW (default 512) through s2kit_cuda_multi_* (multi.cu) */
static s2kit_cuda_multi* g_multi = NULL;
static int g_multi_bw = 0;
static pthread_mutex_t g_multi_lock = PTHREAD_MUTEX_INITIALIZER;

static s2kit_cuda_multi* multi_for(int bw, int fmt) {
    int ngpu = env_int("S2KIT_CUDA_NGPU", 1);
    if (ngpu <= 1 || fmt != S2KIT_COMPLEX || bw < env_int("S2KIT_CUDA_MULTI_MIN_BW", 512)) return NULL;
    if (s2kit_cuda_shard_layout(bw, ngpu, 0, NULL, NULL) < 0) return NULL;
    if (g_multi && g_multi_bw != bw) {
        s2kit_cuda_multi_destroy(g_multi);
        g_multi = NULL;
    }
    if (!g_multi) {
        if (s2kit_cuda_multi_create(&g_multi, bw, ngpu, NULL)) die("multi_create");
        g_multi_bw = bw;
    }
    return g_multi;
}

/* drops every cached plan (frees device memory); not part of the reference API.  Must not race with transforms. */
void s2kit_compat_release(void) {
    pthread_mutex_lock(&g_lock);
    for (int i = 0; i < MAX_CACHED; ++i) entry_destroy(&g_cache[i]);
    pthread_mutex_unlock(&g_lock);
    pthread_mutex_lock(&g_multi_lock);
    if (g_multi) s2kit_cuda_multi_destroy(g_multi);
    g_multi = NULL;
    pthread_mutex_unlock(&g_multi_lock);
}

/* ------------------------------------------------------------------------------------ layout arithmetic */

/* entries of degree l in order m's packed table (cospml.c:250-258) */
int RowSize(const int m, const int l) {
    if (l < m) return 0;
    return ((m % 2) ? (l - 1) : l) / 2 + 1;
}

/* doubles before degree l in order m's packed table (cospml.c:123-134).  H(d) = sum_{j<d} (j/2 + 1); an odd
   order's rows are those of the even order below it shifted by one degree. */
static int half_sum(int d) {
    int q = d / 2;
    return q * (q + 1) + ((d % 2) ? q + 1 : 0);
}
int TableOffset(int m, int l) {
    if (m % 2) return half_sum(l - 1) - half_sum(m - 1);
    return half_sum(l) - half_sum(m);
}

/* doubles in order m's packed table (cospml.c:39-59): all rows m..bw-1 */
int TableSize(const int m, const int bw) {
    if (m >= bw) return 0;
    return TableOffset(m, bw - 1) + RowSize(m, bw - 1);
}

/* cospml.c:107-115 */
int Reduced_SpharmonicTableSize(const int bw, const int m) {
    int total = 0;
    for (int o = 0; o < m; ++o) total += TableSize(o, bw);
    return total;
}

/* cospml.c:85-92: closed-form upper bound up to bw = 512, exact sum above */
int Spharmonic_TableSize(const int bw) {
    if (bw > 512) return Reduced_SpharmonicTableSize(bw, bw);
    return (4 * bw * bw * bw + 6 * bw * bw - 8 * bw) / 24 + bw;
}

/* cospml.c:434-440: theta-space tables of orders m..bw-1, 2bw samples per degree */
int Reduced_Naive_TableSize(const int bw, const int m) {
    int degrees = 0;
    for (int o = m; o < bw; ++o) degrees += bw - o;
    return 2 * bw * degrees;
}

/* entries of cosine index `row` in the transposed table of order m: the degrees l >= max(first, m) of matching
   parity that carry that index.  Equals cospml.c:270-288 for even bw (the reference's odd-bw values are
   inconsistent with its own TableSize, SURVEY.md section 0 trap 4). */
static int transposed_first_degree(int row, int m) {
    if (m % 2) return row >= m ? row + 1 : m + (row % 2);
    return row > m ? row : m + (row % 2);
}
int Transpose_RowSize(const int row, const int m, const int bw) {
    if (row >= bw) return 0;
    if ((m % 2) && row == bw - 1) return 0;
    int first = transposed_first_degree(row, m);
    if (first >= bw) return 0;
    return (bw - 1 - first) / 2 + 1;
}

/* util.c:42-49 */
int IndexOfHarmonicCoeff(const int m, const int l, const int bw) {
    if (m >= 0) return m * bw - (m * (m - 1)) / 2 + (l - m);
    int a = -m;
    /* bw(bw+1)/2 entries of the non-negative orders, then orders -(bw-1) .. -(a+1) */
    return bw * (bw + 1) / 2 + ((bw - 1 - a) * (bw - a)) / 2 + (l - a);
}

/* ------------------------------------------------------------------------------------ setup */

/* chebyshev_nodes.c:16-34 (host helpers some callers use directly, e.g. test/test_DLT_semi.c:54) */
void AcosOfChebyshevNodes(const int n, double* eval_points) {
    const double den = 2. * n;
    for (int i = 0; i < n; ++i) eval_points[i] = (2. * i + 1.) * M_PI / den;
}

void ChebyshevNodes(const int n, double* eval_points) {
    const double den = 2. * n;
    for (int i = 0; i < n; ++i) eval_points[i] = cos((2. * i + 1.) * M_PI / den);
}

void GenerateWeightsForDLT(const int bw, double* weights) { s2k_host_weights(bw, weights); }

/* pmm.c:21-33 (setup seed, libm in the reference's expression order like the other seeds in host_setup.c) */
void Pmm_L2(const int m, double* eval_points, const int n, double* result) {
    double c = sqrt(m + 0.5);
    for (int i = 0; i < m; ++i) c *= sqrt((m - (i / 2.)) / ((double)m - i));
    if (m) c *= pow(2., -m / 2.);
    if (!isfinite(c)) {
        /* m >= 2044: the reference's running product overflows before 2^(-m/2) is applied and it returns NaN.  Only there
           the factor is folded into the product (as s2k_host_seeds does for the device tables); validated against mpmath
           (tests/golden/mp_high_orders.npz).  Smaller orders keep the reference's exact arithmetic. */
        c = sqrt(m + 0.5);
        for (int i = 0; i < m; ++i) c *= sqrt((m - (i / 2.)) / ((double)m - i)) * M_SQRT1_2;
    }
    if (m % 2) c *= -1.;
    for (int i = 0; i < n; ++i) result[i] = c * pow(sin(eval_points[i]), m);
}

/* cospml.c:161-242: generated on the device, exported in the reference's packed layout */
void GenerateCosPmlTable(const int bw, const int m, double* tablespace, double* workspace) {
    (void)workspace;
    Checkout c = ctx_acquire(bw, S2KIT_CUDA_MEMO);
    int rc = s2kit_cuda_table_export(c.plan, m, tablespace);
    ctx_release(c);
    if (rc) die("GenerateCosPmlTable");
}

/* cospml.c:301-362: gather each cosine index's column of the packed table (ascending degree) */
void TransposeCosPmlTable(const int bw, const int m, double* cos_pml_table, double* result) {
    double* out = result;
    for (int row = 0; row < bw; ++row) {
        int count = Transpose_RowSize(row, m, bw);
        int l = transposed_first_degree(row, m);
        for (int i = 0; i < count; ++i, l += 2) *out++ = cos_pml_table[TableOffset(m, l) + row / 2];
    }
}

double** Spharmonic_Pml_Table(const int bw, double* resultspace, double* workspace) {
    double** t = (double**)malloc(sizeof(double*) * bw);
    double* at = resultspace;
    for (int m = 0; m < bw; ++m) {
        t[m] = at;
        GenerateCosPmlTable(bw, m, at, workspace);
        at += TableSize(m, bw);
    }
    return t;
}

double** Transpose_Spharmonic_Pml_Table(double** spharmonic_pml_table, const int bw, double* resultspace) {
    double** t = (double**)malloc(sizeof(double*) * bw);
    double* at = resultspace;
    for (int m = 0; m < bw; ++m) {
        t[m] = at;
        TransposeCosPmlTable(bw, m, spharmonic_pml_table[m], at);
        at += TableSize(m, bw);
    }
    return t;
}

/* pml.c:41-79: theta-space table of order m at the 2bw Chebyshev nodes (host; only ever read by callers, the
   GPU engine uses the seminaive algorithm for every order) */
void GeneratePmlTable(const int bw, const int m, double* pml_table, double* workspace) {
    (void)workspace;
    const int n = 2 * bw;
    double* buf = (double*)malloc(sizeof(double) * 3 * n);
    double *x = buf, *older = buf + n, *cur = buf + 2 * n;
    const double den = 2. * n;
    double c = sqrt(m + 0.5);
    for (int i = 0; i < m; ++i) c *= sqrt((m - (i / 2.)) / ((double)m - i));
    if (m) c *= pow(2., -m / 2.);
    if (m % 2) c *= -1.;
    for (int i = 0; i < n; ++i) {
        double theta = (2. * i + 1.) * M_PI / den;
        x[i] = cos((2. * i + 1.) * M_PI / den);
        older[i] = 0.;
        cur[i] = m ? c * pow(sin(theta), m) : M_SQRT1_2;
    }
    memcpy(pml_table, cur, sizeof(double) * n);
    for (int l = m; l + 1 < bw; ++l) {
        double a = sqrt(((2. * l + 3.) / (2. * l + 1.)) * ((l - m + 1.) / (l + m + 1.))) * ((2. * l + 1.) / (l - m + 1.));
        double cc = 0.;
        if (l)
            cc = -1.0 *
                 sqrt(((2. * l + 3.) / (2. * l - 1.)) * ((l - m + 1.) / (l + m + 1.)) *
                      (((double)l - m) / ((double)l + m))) *
                 ((l + m) / (l - m + 1.));
        double* next = pml_table + (size_t)(l + 1 - m) * n;
        for (int i = 0; i < n; ++i) {
            double t1 = cc * older[i];
            double t2 = cur[i] * x[i];
            double t3 = a * t2;
            next[i] = t3 + t1;
        }
        memcpy(older, cur, sizeof(double) * n);
        memcpy(cur, next, sizeof(double) * n);
    }
    free(buf);
}

/* cospml.c:451-475: cosine tables below the cutoff, theta-space tables from the cutoff on */
double** SemiNaive_Naive_Pml_Table(const int bw, const int m, double* resultspace, double* workspace) {
    double** t = (double**)malloc(sizeof(double*) * (bw + 1));
    double* at = resultspace;
    for (int o = 0; o < bw; ++o) {
        t[o] = at;
        if (o < m) {
            GenerateCosPmlTable(bw, o, at, workspace);
            at += TableSize(o, bw);
        } else {
            GeneratePmlTable(bw, o, at, workspace);
            at += 2 * bw * (bw - o);
        }
    }
    t[bw] = at;
    return t;
}

/* cospml.c:490-518 */
double** Transpose_SemiNaive_Naive_Pml_Table(double** seminaive_naive_pml_table, const int bw, const int m,
                                             double* resultspace, double* workspace) {
    double** t = (double**)malloc(sizeof(double*) * (bw + 1));
    double* at = resultspace;
    for (int o = 0; o < bw; ++o) {
        t[o] = at;
        if (o < m) {
            TransposeCosPmlTable(bw, o, seminaive_naive_pml_table[o], at);
            at += TableSize(o, bw);
        } else {
            GeneratePmlTable(bw, o, at, workspace);
            at += 2 * bw * (bw - o);
        }
    }
    t[bw] = at;
    return t;
}

/* ------------------------------------------------------------------------------------ transforms */

static void run_fst(int variant, double* rdata, double* idata, double* rcoeffs, double* icoeffs, int bw, int fmt) {
    long gs = 4L * bw * bw, cs = (long)bw * bw;
    if (variant == S2KIT_CUDA_MEMO) {
        pthread_mutex_lock(&g_multi_lock);
        s2kit_cuda_multi* mp = multi_for(bw, fmt);
        if (mp) {
            int rc = s2kit_cuda_multi_fst(mp, rdata, idata, rcoeffs, icoeffs);
            pthread_mutex_unlock(&g_multi_lock);
            if (rc) die("FSTSemiMemo (multi-GPU)");
            return;
        }
        pthread_mutex_unlock(&g_multi_lock);
    }
    Checkout c = ctx_acquire(bw, variant);
    int rc = s2kit_cuda_fst(c.plan, rdata, idata, rcoeffs, icoeffs, 1, gs, cs, fmt, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("FSTSemi");
}

static void run_inv(int variant, double* rcoeffs, double* icoeffs, double* rdata, double* idata, int bw, int fmt) {
    long gs = 4L * bw * bw, cs = (long)bw * bw;
    if (variant == S2KIT_CUDA_MEMO) {
        pthread_mutex_lock(&g_multi_lock);
        s2kit_cuda_multi* mp = multi_for(bw, fmt);
        if (mp) {
            int rc = s2kit_cuda_multi_inv_fst(mp, rcoeffs, icoeffs, rdata, idata);
            pthread_mutex_unlock(&g_multi_lock);
            if (rc) die("InvFSTSemiMemo (multi-GPU)");
            return;
        }
        pthread_mutex_unlock(&g_multi_lock);
    }
    Checkout c = ctx_acquire(bw, variant);
    int rc = s2kit_cuda_inv_fst(c.plan, rcoeffs, icoeffs, rdata, idata, 1, cs, gs, fmt, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("InvFSTSemi");
}

static void run_fzt(int variant, double* rdata, double* idata, double* rres, double* ires, int bw, int fmt) {
    Checkout c = ctx_acquire(bw, variant);
    int rc = s2kit_cuda_fzt(c.plan, rdata, idata, rres, ires, 1, 4L * bw * bw, bw, fmt, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("FZTSemi");
}

static void run_conv(int variant, double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
                     double* ires, int bw) {
    Checkout c = ctx_acquire(bw, variant);
    long gs = 4L * bw * bw;
    int rc = s2kit_cuda_conv(c.plan, rdata, idata, rfilter, ifilter, rres, ires, 1, gs, gs, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("ConvOn2SphereSemi");
}

void FSTSemiMemo(double* rdata, double* idata, double* rcoeffs, double* icoeffs, const int bw,
                 double** seminaive_naive_table, double* workspace, DataFormat data_format, const int cutoff,
                 fftw_plan* DCT_plan, fftw_plan* FFT_plan, double* weights) {
    (void)seminaive_naive_table; (void)workspace; (void)cutoff; (void)DCT_plan; (void)FFT_plan; (void)weights;
    run_fst(S2KIT_CUDA_MEMO, rdata, idata, rcoeffs, icoeffs, bw, (int)data_format);
}

void InvFSTSemiMemo(double* rcoeffs, double* icoeffs, double* rdata, double* idata, const int bw,
                    double** transpose_seminaive_naive_table, double* workspace, DataFormat data_format,
                    const int cutoff, fftw_plan* inv_DCT_plan, fftw_plan* inv_FFT_plan) {
    (void)transpose_seminaive_naive_table; (void)workspace; (void)cutoff; (void)inv_DCT_plan; (void)inv_FFT_plan;
    run_inv(S2KIT_CUDA_MEMO, rcoeffs, icoeffs, rdata, idata, bw, (int)data_format);
}

void FZTSemiMemo(double* rdata, double* idata, double* rres, double* ires, const int bw, double* cos_pml_table,
                 double* workspace, const DataFormat data_format, fftw_plan* DCT_plan, double* weights) {
    (void)cos_pml_table; (void)workspace; (void)DCT_plan; (void)weights;
    run_fzt(S2KIT_CUDA_MEMO, rdata, idata, rres, ires, bw, (int)data_format);
}

void ConvOn2SphereSemiMemo(double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
                           double* ires, const int bw, double* workspace) {
    (void)workspace;
    run_conv(S2KIT_CUDA_MEMO, rdata, idata, rfilter, ifilter, rres, ires, bw);
}

void FSTSemiFly(double* rdata, double* idata, double* rcoeffs, double* icoeffs, const int bw, double* workspace,
                DataFormat data_format, const int cutoff, fftw_plan* DCT_plan, fftw_plan* FFT_plan,
                double* weights) {
    (void)workspace; (void)cutoff; (void)DCT_plan; (void)FFT_plan; (void)weights;
    run_fst(S2KIT_CUDA_FLY, rdata, idata, rcoeffs, icoeffs, bw, (int)data_format);
}

void InvFSTSemiFly(double* rcoeffs, double* icoeffs, double* rdata, double* idata, const int bw, double* workspace,
                   DataFormat data_format, const int cutoff, fftw_plan* inv_DCT_plan, fftw_plan* inv_FFT_plan) {
    (void)workspace; (void)cutoff; (void)inv_DCT_plan; (void)inv_FFT_plan;
    run_inv(S2KIT_CUDA_FLY, rcoeffs, icoeffs, rdata, idata, bw, (int)data_format);
}

void FZTSemiFly(double* rdata, double* idata, double* rres, double* ires, const int bw, double* workspace,
                DataFormat data_format, fftw_plan* DCT_plan, double* weights) {
    (void)workspace; (void)DCT_plan; (void)weights;
    run_fzt(S2KIT_CUDA_FLY, rdata, idata, rres, ires, bw, (int)data_format);
}

void ConvOn2SphereSemiFly(double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
                          double* ires, const int bw, double* workspace) {
    (void)workspace;
    run_conv(S2KIT_CUDA_FLY, rdata, idata, rfilter, ifilter, rres, ires, bw);
}

/* seminaive.c:153-198 / 56-115, single column, single order */
void DLTSemi(double* data, const int bw, const int m, double* result, double* workspace, double* cos_pml_table,
             double* weights, fftw_plan* plan) {
    (void)workspace; (void)cos_pml_table; (void)weights; (void)plan;
    Checkout c = ctx_acquire(bw, S2KIT_CUDA_MEMO);
    int rc = s2kit_cuda_dlt_semi(c.plan, data, m, result, 1, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("DLTSemi");
}

void InvDLTSemi(double* coeffs, const int bw, const int m, double* result, double* trans_cos_pml_table,
                double* sin_values, double* workspace, fftw_plan* plan) {
    (void)trans_cos_pml_table; (void)sin_values; (void)workspace; (void)plan;
    Checkout c = ctx_acquire(bw, S2KIT_CUDA_MEMO);
    int rc = s2kit_cuda_inv_dlt_semi(c.plan, coeffs, m, result, 1, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("InvDLTSemi");
}

/* naive.c:35-60 */
void DLTNaive(double* data, const int bw, const int m, double* weights, double* result, double* pml_table,
              double* workspace) {
    (void)workspace;
    if (s2kit_cuda_dlt_naive(data, bw, m, weights, result, pml_table, S2KIT_CUDA_HOST)) die("DLTNaive");
}

/* naive.c:77-95 */
void InvDLTNaive(double* coeffs, const int bw, const int m, double* result, double* pml_table) {
    if (s2kit_cuda_inv_dlt_naive(coeffs, bw, m, result, pml_table, S2KIT_CUDA_HOST)) die("InvDLTNaive");
}

/* util.c:68-103 */
void TransMult(double* rdatacoeffs, double* idatacoeffs, double* rfiltercoeffs, double* ifiltercoeffs, double* rres,
               double* ires, const int bw) {
    Checkout c = ctx_acquire(bw, S2KIT_CUDA_MEMO);
    int rc = s2kit_cuda_trans_mult(c.plan, rdatacoeffs, idatacoeffs, rfiltercoeffs, ifiltercoeffs, rres, ires, 1,
                                   (long)bw * bw, S2KIT_CUDA_HOST);
    ctx_release(c);
    if (rc) die("TransMult");
}
