/* host_setup.h -- host-side (libm) setup quantities; see host_setup.c */
#ifndef S2K_HOST_SETUP_H
#define S2K_HOST_SETUP_H
#ifdef __cplusplus
extern "C" {
#endif
void s2k_host_weights(int bw, double* w);                          /* 4 bw */
void s2k_host_nodes(int bw, double* x);                            /* bw */
void s2k_host_sines(int bw, double* s);                            /* 2 bw */
void s2k_host_seeds(int bw, int m_lo, int m_hi, double* seeds);    /* (m_hi-m_lo) * bw */
void s2k_host_twiddles(int n, double* tw);                         /* 2 n */
void s2k_host_quarter(int n, double* qt);                          /* 8 n */
void s2k_host_reordered(int bw, const double* weights, const double* sines, double* wv, double* sv); /* 4bw, 2bw */
#ifdef __cplusplus
}
#endif
#endif
