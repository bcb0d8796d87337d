/*
 * s2kit.h -- S2kit's public C API, as exported by libs2kit_cuda.so (drop-in boundary).
 *
 * One header for the prototypes the reference spreads over the files of include/s2kit/ (the same file names exist
 * here as forwarding headers); each block cites the
 * declaration it replaces.  Names, argument order and meaning are the reference's; the parameters are named
 * here (the reference leaves them anonymous) and extern "C" guards are added.  Where FFTW's header is not
 * available, `fftw_plan` is declared as an opaque pointer: the GPU engine accepts and ignores the plan
 * arguments.
 *
 * Differences from the reference a caller can observe:
 *  - bandwidths 2 .. 2048 (any value; powers of two >= 16 run on the radix-FFT kernels, others on direct O(n^2)
 *    kernels); bw > 2048 is refused: the void entry points print the reason and abort(), there is no CPU fallback;
 *  - every order uses the seminaive algorithm, `cutoff` is accepted and ignored;
 *  - Transpose_RowSize / TransposeCosPmlTable agree with the reference for every even bw; for odd bw the reference's
 *    row sizes do not add up to TableSize (its odd-bw inverse is wrong, cospml.c:270-288) -- here they do, and the
 *    inverse transform is the exact transpose of the forward one;
 *  - at bw = 2048 the orders |m| >= 2044, where the reference returns NaN (pmm.c:22-30), are finite;
 *  - the entry points may be called from several host threads at once (each call checks a private context out of a
 *    per-bandwidth cache); FSTSemiMemo / InvFSTSemiMemo use several GPUs for one field when S2KIT_CUDA_NGPU > 1.
 */
#ifndef S2KIT_H
#define S2KIT_H

#if defined(S2KIT_USE_FFTW3_H)
#include <fftw3.h>
#elif !defined(FFTW_ESTIMATE) && !defined(ORACLE_FFTW3_STUB_H)
typedef struct fftw_plan_s* fftw_plan;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* include/s2kit/util.h:10-17 */
#ifndef _UTIL_H
typedef enum { COMPLEX = 0, REAL } DataFormat;
#endif
int IndexOfHarmonicCoeff(const int m, const int l, const int bw);
void TransMult(double* rdatacoeffs, double* idatacoeffs, double* rfiltercoeffs, double* ifiltercoeffs, double* rres,
               double* ires, const int bw);

/* include/s2kit/FST_semi_memo.h:8-16 */
void FSTSemiMemo(double* rdata, double* idata, double* rcoeffs, double* icoeffs, const int bw,
                 double** seminaive_naive_table, double* workspace, DataFormat data_format, const int cutoff,
                 fftw_plan* DCT_plan, fftw_plan* FFT_plan, double* weights);
void InvFSTSemiMemo(double* rcoeffs, double* icoeffs, double* rdata, double* idata, const int bw,
                    double** transpose_seminaive_naive_table, double* workspace, DataFormat data_format,
                    const int cutoff, fftw_plan* inv_DCT_plan, fftw_plan* inv_FFT_plan);
void FZTSemiMemo(double* rdata, double* idata, double* rres, double* ires, const int bw, double* cos_pml_table,
                 double* workspace, const DataFormat data_format, fftw_plan* DCT_plan, double* weights);
void ConvOn2SphereSemiMemo(double* rdata, double* idata, double* rfilter, double* ifilter, double* rres,
                           double* ires, const int bw, double* workspace);

/* include/s2kit/FST_semi_fly.h:8-16 */
void FSTSemiFly(double* rdata, double* idata, double* rcoeffs, double* icoeffs, const int bw, double* workspace,
                DataFormat data_format, const int cutoff, fftw_plan* DCT_plan, fftw_plan* FFT_plan, double* weights);
void InvFSTSemiFly(double* rcoeffs, double* icoeffs, double* rdata, double* idata, const int bw, double* workspace,
                   DataFormat data_format, const int cutoff, fftw_plan* inv_DCT_plan, fftw_plan* inv_FFT_plan);
void FZTSemiFly(double* rdata, double* idata, double* rres, double* ires, const int bw, double* workspace,
                DataFormat data_format, fftw_plan* DCT_plan, double* weights);
void ConvOn2SphereSemiFly(double* rdata, double* idata, double* rfilter, double* ifilter, double* rres, double* ires,
                          const int bw, double* workspace);

/* include/s2kit/seminaive.h:6-8 */
void DLTSemi(double* data, const int bw, const int m, double* result, double* workspace, double* cos_pml_table,
             double* weights, fftw_plan* plan);
void InvDLTSemi(double* coeffs, const int bw, const int m, double* result, double* trans_cos_pml_table,
                double* sin_values, double* workspace, fftw_plan* plan);

/* include/s2kit/naive.h:4-6 (theta-space table from GeneratePmlTable; the GPU does the dense products) */
void DLTNaive(double* data, const int bw, const int m, double* weights, double* result, double* pml_table,
              double* workspace);
void InvDLTNaive(double* coeffs, const int bw, const int m, double* result, double* pml_table);

/* include/s2kit/pmm.h:4.  Same libm expression as the reference (bit-identical) wherever the reference is finite; for
 * m >= 2044 the reference's product overflows and it returns NaN -- here the 2^(-m/2) factor is folded into the product
 * so the result is the finite value of the definition (checked against mpmath, tests/golden/mp_high_orders.npz). */
void Pmm_L2(const int m, double* eval_points, const int n, double* result);

/* include/s2kit/chebyshev_nodes.h:4-6 */
void AcosOfChebyshevNodes(const int n, double* eval_points);
void ChebyshevNodes(const int n, double* eval_points);

/* include/s2kit/weights.h:4 */
void GenerateWeightsForDLT(const int bw, double* weights);

/* include/s2kit/cospml.h:6-30, include/s2kit/pml.h:4 */
int TableSize(const int m, const int bw);
int Spharmonic_TableSize(const int bw);
int Reduced_SpharmonicTableSize(const int bw, const int m);
int Reduced_Naive_TableSize(const int bw, const int m);
int TableOffset(int m, int l);
int RowSize(const int m, const int l);
int Transpose_RowSize(const int row, const int m, const int bw);
void GenerateCosPmlTable(const int bw, const int m, double* tablespace, double* workspace);
void TransposeCosPmlTable(const int bw, const int m, double* cos_pml_table, double* result);
void GeneratePmlTable(const int bw, const int m, double* pml_table, double* workspace);
double** Spharmonic_Pml_Table(const int bw, double* resultspace, double* workspace);
double** Transpose_Spharmonic_Pml_Table(double** spharmonic_pml_table, const int bw, double* resultspace);
double** SemiNaive_Naive_Pml_Table(const int bw, const int m, double* resultspace, double* workspace);
double** Transpose_SemiNaive_Naive_Pml_Table(double** seminaive_naive_pml_table, const int bw, const int m,
                                             double* resultspace, double* workspace);

/* not in the reference: frees the per-bandwidth device plans the functions above create lazily */
void s2kit_compat_release(void);

#ifdef __cplusplus
}
#endif
#endif /* S2KIT_H */
