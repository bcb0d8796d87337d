/*
 * host_setup.c -- the bit-sensitive setup quantities, computed on the HOST with glibc libm.
 *
 * The reference's outputs at bw >= 1024 move by more than the 1e-10 parity tolerance when the Chebyshev
 * nodes change by one ulp (trap 1), so everything that seeds the Legendre tables is
 * evaluated here with the same libm calls and the same expression order as the reference, then uploaded.
 * Built with -ffp-contract=off so no multiply-add is fused.
 *
 *   s2k_host_weights   GenerateWeightsForDLT        src/legendre_transform/weights.c:32-47
 *   s2k_host_nodes     ChebyshevNodes               src/util/chebyshev_nodes.c:29-34
 *   s2k_host_sines     sin of AcosOfChebyshevNodes  src/FST_semi_memo.c:244-246, chebyshev_nodes.c:16-21
 *   s2k_host_seeds     Pmm_L2 (+ 1/sin for odd m)   src/legendre_polynomials/pmm.c:21-33, cospml.c:181-192
 * The FFT/DCT twiddle tables are new (the reference delegates to FFTW) and only need to be accurate.
 */
#include "host_setup.h"

#include <math.h>

void s2k_host_weights(int bw, double* w) {
    const double q = M_PI / (4. * bw);
    for (int j = 0; j < 2 * bw; ++j) {
        const double odd = 2. * j + 1.;
        double series = 0.;
        for (int k = 0; k < bw; ++k) series += 1. / (2. * k + 1.) * sin(odd * (2. * k + 1.) * q);
        series *= 2. * sin(odd * q) / bw;
        w[j] = series;
        w[2 * bw + j] = series * sin(odd * q);
    }
}

void s2k_host_nodes(int bw, double* x) {
    const double den = 2. * bw;
    for (int i = 0; i < bw; ++i) x[i] = cos((2. * i + 1.) * M_PI / den);
}

void s2k_host_sines(int bw, double* s) {
    const int n = 2 * bw;
    const double den = 2. * n;
    for (int j = 0; j < n; ++j) {
        double theta = (2. * j + 1.) * M_PI / den;
        s[j] = sin(theta);
    }
}

/* seeds[m*bw + i] = c_m sin(theta_i)^m, divided by sin(theta_i) when m is odd; theta_i on the bw-point grid */
void s2k_host_seeds(int bw, int m_lo, int m_hi, double* seeds) {
    const double den = 2. * bw;
    for (int m = m_lo; m < m_hi; ++m) {
        double* row = seeds + (long)(m - m_lo) * bw;
        if (m == 0) {
            for (int i = 0; i < bw; ++i) row[i] = M_SQRT1_2;
            continue;
        }
        double c = sqrt(m + 0.5);
        for (int i = 0; i < m; ++i) c *= sqrt((m - (i / 2.)) / ((double)m - i));
        c *= pow(2., -m / 2.);
        if (m % 2) c *= -1.;
        if (!isfinite(c)) {
            /* The reference's running product overflows for m >= 2044 and its tables become NaN (pmm.c:22-30,
               SURVEY.md section 0 trap 3).  Only there, fold the 2^(-m/2) into the product so it stays finite;
               orders m <= 2043 keep the reference's exact arithmetic. */
            c = sqrt(m + 0.5);
            for (int i = 0; i < m; ++i) c *= sqrt((m - (i / 2.)) / ((double)m - i)) * M_SQRT1_2;
            if (m % 2) c *= -1.;
        }
        for (int i = 0; i < bw; ++i) {
            double theta = (2. * i + 1.) * M_PI / den;
            double v = c * pow(sin(theta), m);
            if (m % 2) v /= sin(theta);
            row[i] = v;
        }
    }
}

/* tw[2q] = cos(2 pi q / n), tw[2q+1] = -sin(2 pi q / n), q < n: octant-reduced so the table is exactly symmetric */
void s2k_host_twiddles(int n, double* tw) {
    for (int q = 0; q < n; ++q) {
        double a = 2.0 * M_PI * (double)q / (double)n;
        tw[2 * q] = cos(a);
        tw[2 * q + 1] = -sin(a);
    }
    if (n % 4 == 0) {
        tw[2 * (n / 4)] = 0.0;
        tw[2 * (n / 4) + 1] = -1.0;
        tw[2 * (n / 2)] = -1.0;
        tw[2 * (n / 2) + 1] = 0.0;
        tw[2 * (3 * n / 4)] = 0.0;
        tw[2 * (3 * n / 4) + 1] = 1.0;
    } else if (n % 2 == 0) {
        tw[2 * (n / 2)] = -1.0;
        tw[2 * (n / 2) + 1] = 0.0;
    }
}

/* qt[2q] = cos(pi q / 2n), qt[2q+1] = sin(pi q / 2n), q < 4n */
void s2k_host_quarter(int n, double* qt) {
    for (int q = 0; q < 4 * n; ++q) {
        double a = M_PI * (double)q / (2.0 * (double)n);
        qt[2 * q] = cos(a);
        qt[2 * q + 1] = sin(a);
    }
    qt[2 * n] = 0.0;          /* q = n: cos(pi/2) */
    qt[2 * (3 * n)] = 0.0;    /* q = 3n */
    qt[2 * (2 * n) + 1] = 0.0; /* q = 2n: sin(pi) */
}

/* The DCT kernels read their input in even/odd-reordered positions p -> j(p) = 2p (p < n/2), 2(n-1-p)+1 otherwise;
   these copies of the weights / sines are stored in that order so the loads are contiguous.
   wv[par*n + p] = weights[par*n + j(p)] (par = parity of the order), sv[p] = sines[j(p)], n = 2 bw. */
void s2k_host_reordered(int bw, const double* weights, const double* sines, double* wv, double* sv) {
    const int n = 2 * bw;
    for (int p = 0; p < n; ++p) {
        int j = (p < bw) ? 2 * p : 2 * (n - 1 - p) + 1;
        wv[p] = weights[j];
        wv[n + p] = weights[n + j];
        sv[p] = sines[j];
    }
}
