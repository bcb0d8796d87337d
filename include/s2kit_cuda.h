/*
 * s2kit_cuda.h -- thin C-ABI of the B200-native spherical harmonic transform engine.
 *
 * This is the boundary the C host layer (s2kit_b200/csrc/s2kit_compat.c: FSTSemiMemo, InvFSTSemiMemo,
 * FZTSemiMemo, ConvOn2SphereSemiMemo and the -SemiFly twins, same signatures as the reference's
 * include/s2kit/FST_semi_memo.h:8-16 and FST_semi_fly.h:8-16) calls into.  Plain pointers and sizes only.
 * All arithmetic is IEEE FP64.  There is no CPU fallback: every entry point fails with a non-zero code
 * (and s2kit_cuda_last_error() text) when no CUDA device is usable.
 *
 * Conventions shared with the reference:
 *   grid        2bw x 2bw, latitude-major: data[j*2bw + k] = f(theta_j, phi_k), split re / im arrays
 *   coefficients bw*bw per re / im array, order m = 0..bw-1 then -(bw-1)..-1, degree l = |m|..bw-1
 *               (IndexOfHarmonicCoeff, src/util/util.c:42-49)
 *   data_format  S2KIT_COMPLEX = 0, S2KIT_REAL = 1 (include/s2kit/util.h:10-13)
 * Batched calls take `batch` functions; function f starts at base + f*stride (strides in doubles).
 */
#ifndef S2KIT_CUDA_H
#define S2KIT_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s2kit_cuda_plan s2kit_cuda_plan;

enum { S2KIT_CUDA_MEMO = 0, S2KIT_CUDA_FLY = 1 };   /* table policy: resident (Memo) / regenerated per call (Fly) */
enum { S2KIT_COMPLEX = 0, S2KIT_REAL = 1 };
enum { S2KIT_CUDA_HOST = 0, S2KIT_CUDA_DEVICE = 1 }; /* where the data pointers of a call live */

/* kernel kinds reported by s2kit_cuda_profile_get() */
enum {
    S2KIT_K_PHI_FFT_FWD = 0,   /* K1: longitude FFT + scale + transpose to order-major   (FST_semi_memo.c:81-90)  */
    S2KIT_K_DCT_FWD = 1,       /* K2: weights * DCT-II(2bw), first bw outputs            (seminaive.c:162-176)    */
    S2KIT_K_LEGENDRE_FWD = 2,  /* K3: triangular contraction with the cosine tables      (seminaive.c:183-197)    */
    S2KIT_K_LEGENDRE_INV = 3,  /* K4: transposed contraction                             (seminaive.c:74-95)      */
    S2KIT_K_DCT_INV = 4,       /* K5: DCT-III(2bw) * sin(theta)                          (seminaive.c:98-114)     */
    S2KIT_K_PHI_FFT_INV = 5,   /* K6: inverse longitude FFT                              (FST_semi_memo.c:344-350)*/
    S2KIT_K_TABLE_GEN = 6,     /* K7: recurrence + DCT-II(bw) + pack                     (cospml.c:161-242)       */
    S2KIT_K_ZONAL = 7,         /* K8a: m = 0 row sums                                    (FST_semi_memo.c:386-399)*/
    S2KIT_K_SPECTRAL_MUL = 8,  /* K8b: TransMult                                         (util.c:68-103)          */
    S2KIT_K_FUSED_FWD = 9,     /* K2+K3 in one kernel (batched path, bw 64..512)         (seminaive.c:153-198)    */
    S2KIT_K_FUSED_INV = 10,    /* K4+K5 in one kernel                                    (seminaive.c:56-115)     */
    S2KIT_K_COUNT = 11
};

/* ---- plans ------------------------------------------------------------------------------------- */

/* Creates a plan for bandwidth bw on CUDA device `device`.  Computes the bit-sensitive seeds (Chebyshev
 * nodes, P_m^m, quadrature weights: chebyshev_nodes.c, pmm.c, weights.c) on the host with libm, uploads
 * them and (Memo) generates the device-resident cosine tables for orders [0, bw).
 * max_batch bounds the number of functions processed per internal chunk (sizes the device workspace). */
int s2kit_cuda_plan_create(s2kit_cuda_plan** out, int bw, int variant, int max_batch, int device);

/* Same, for rank `rank` of `nranks` cooperating processes that share ONE transform (single large-bw
 * field): the plan keeps tables only for its own orders (m paired with bw-1-m for balance) and its own
 * latitude rings.  The ring<->order exchange itself is done by the caller (NCCL all-to-all) between
 * s2kit_cuda_fst_rings() / s2kit_cuda_fst_orders(). */
int s2kit_cuda_plan_create_sharded(s2kit_cuda_plan** out, int bw, int variant, int device, int rank, int nranks);

/* A second plan object for the same bandwidth that shares `src`'s device tables and constants (read-only after
 * creation) and owns its stream and workspaces.  A plan object runs one transform at a time (every entry point takes
 * the plan's lock); concurrent host threads use one clone each.  `src` must outlive its clones. */
int s2kit_cuda_plan_clone(s2kit_cuda_plan** out, const s2kit_cuda_plan* src, int max_batch);

int s2kit_cuda_plan_destroy(s2kit_cuda_plan* plan);

/* Run on a caller-provided cudaStream_t (e.g. the framework's current stream) instead of the plan's own. */
int s2kit_cuda_plan_set_stream(s2kit_cuda_plan* plan, void* cuda_stream);
void* s2kit_cuda_plan_stream(s2kit_cuda_plan* plan);
int s2kit_cuda_synchronize(s2kit_cuda_plan* plan);

int s2kit_cuda_plan_bw(const s2kit_cuda_plan* plan);
/* bytes of device memory held by the cosine tables (tile-padded private layout): one copy at bw >= 512, a second,
 * tile-transposed copy for the wide batched inverse kernels below that */
size_t s2kit_cuda_plan_table_bytes(const s2kit_cuda_plan* plan);
/* table bytes ONE transform of a Memo plan reads (a single copy of the plan's tiles) */
size_t s2kit_cuda_plan_table_stream_bytes(const s2kit_cuda_plan* plan);

/* ---- transforms (replace FSTSemiMemo/Fly, InvFSTSemiMemo/Fly, FZTSemiMemo/Fly, ConvOn2SphereSemiMemo/Fly) */

/* forward SHT: grids -> coefficients.  where = S2KIT_CUDA_HOST copies in/out (synchronous);
 * S2KIT_CUDA_DEVICE runs asynchronously on the plan's stream. */
int s2kit_cuda_fst(s2kit_cuda_plan* plan, const double* rdata, const double* idata, double* rcoeffs,
                   double* icoeffs, int batch, long data_stride, long coef_stride, int data_format, int where);

/* inverse SHT: coefficients -> grids */
int s2kit_cuda_inv_fst(s2kit_cuda_plan* plan, const double* rcoeffs, const double* icoeffs, double* rdata,
                       double* idata, int batch, long coef_stride, long data_stride, int data_format, int where);

/* zonal (order-0) transform: grids -> bw coefficients per function (stride res_stride) */
int s2kit_cuda_fzt(s2kit_cuda_plan* plan, const double* rdata, const double* idata, double* rres, double* ires,
                   int batch, long data_stride, long res_stride, int data_format, int where);

/* convolution of real fields with a zonal filter, entirely on the device:
 * FST(REAL) -> FZT(REAL) -> TransMult -> InvFST(REAL); filter_stride = 0 shares one filter */
int s2kit_cuda_conv(s2kit_cuda_plan* plan, const double* rdata, const double* idata, const double* rfilter,
                    const double* ifilter, double* rres, double* ires, int batch, long data_stride,
                    long filter_stride, int where);

/* TransMult (src/util/util.c:68-103) on bw*bw coefficient arrays, ComplexMult signs as in the reference */
int s2kit_cuda_trans_mult(s2kit_cuda_plan* plan, const double* rdata, const double* idata, const double* rfilter,
                          const double* ifilter, double* rres, double* ires, int batch, long coef_stride, int where);

/* 1-D transforms of one order m on `ncols` real columns (replace DLTSemi / InvDLTSemi, seminaive.c:56,153):
 * forward: data[ncols][2bw] -> result[ncols][bw-m]; inverse: coeffs[ncols][bw-m] -> result[ncols][2bw] */
int s2kit_cuda_dlt_semi(s2kit_cuda_plan* plan, const double* data, int m, double* result, int ncols, int where);
int s2kit_cuda_inv_dlt_semi(s2kit_cuda_plan* plan, const double* coeffs, int m, double* result, int ncols, int where);

/* The reference's naive algorithm on one column (replace DLTNaive / InvDLTNaive, src/legendre_transform/naive.c:35,77):
 * dense products with the caller's theta-space table pml_table[bw-m][2bw] (GeneratePmlTable, pml.c:41-79).  No plan:
 * nothing is precomputed.  forward: data[2bw], weights[>= 2bw] -> result[bw-m]; inverse: coeffs[bw-m] -> result[2bw] */
int s2kit_cuda_dlt_naive(const double* data, int bw, int m, const double* weights, double* result,
                         const double* pml_table, int where);
int s2kit_cuda_inv_dlt_naive(const double* coeffs, int bw, int m, double* result, const double* pml_table, int where);

/* ---- sharded single-field halves (see s2kit_cuda_plan_create_sharded) ---------------------------- */
/* forward, stage 1: this rank's latitude rings [ring_lo, ring_lo+nrings) -> longitude FFT, written as
 * send blocks: out[dest_rank][part][local order][local ring]; all pointers device memory */
int s2kit_cuda_fst_rings(s2kit_cuda_plan* plan, const double* rdata, const double* idata, double* sendbuf);
/* forward, stage 2: recvbuf[src_rank][part][local order][ring] -> coefficients of this rank's orders,
 * written at their reference positions inside full bw*bw arrays */
int s2kit_cuda_fst_orders(s2kit_cuda_plan* plan, const double* recvbuf, double* rcoeffs, double* icoeffs);
int s2kit_cuda_inv_fst_orders(s2kit_cuda_plan* plan, const double* rcoeffs, const double* icoeffs, double* sendbuf);
int s2kit_cuda_inv_fst_rings(s2kit_cuda_plan* plan, const double* recvbuf, double* rdata, double* idata);
/* Host-only (no GPU needed): which orders (ascending; returns their count, -1 if the split is not supported) and
 * which spectral rows (2bw/nranks entries, -1 = unused slot) rank `rank` of `nranks` owns. */
int s2kit_cuda_shard_layout(int bw, int nranks, int rank, int* orders_out, int* rows_out);
/* geometry of the exchange: doubles per (src,dst) block, rings per rank, orders rows per rank */
int s2kit_cuda_shard_info(const s2kit_cuda_plan* plan, long* block_doubles, int* rings_per_rank, int* rows_per_rank);

/* ---- one transform on several GPUs of one process (multi.cu) ---------------------------------------
 * Single large-bandwidth field (BASELINE configs[4]: FSTSemiMemo, src/FST_semi_memo.c:68-202, at bw = 2048): latitude
 * rings and orders are split over `ngpu` devices (NULL devices = 0..ngpu-1), every GPU keeps only its orders' tables,
 * and the ring <-> order exchange happens inside the DCT kernels as NVLink loads / stores on peer-mapped buffers -- no
 * separate collective, no NCCL.  Needs peer access between the devices.  COMPLEX format.
 * The drop-in layer routes FSTSemiMemo / InvFSTSemiMemo through this when S2KIT_CUDA_NGPU > 1 (bw >= 512). */
typedef struct s2kit_cuda_multi s2kit_cuda_multi;
int s2kit_cuda_multi_create(s2kit_cuda_multi** out, int bw, int ngpu, const int* devices);
int s2kit_cuda_multi_destroy(s2kit_cuda_multi* mp);
int s2kit_cuda_multi_ngpu(const s2kit_cuda_multi* mp);
size_t s2kit_cuda_multi_table_bytes_per_gpu(const s2kit_cuda_multi* mp);
/* host pointers: full 2bw x 2bw grids, full bw*bw coefficient arrays (every entry written); synchronous */
int s2kit_cuda_multi_fst(s2kit_cuda_multi* mp, const double* rdata, const double* idata, double* rcoeffs,
                         double* icoeffs);
int s2kit_cuda_multi_inv_fst(s2kit_cuda_multi* mp, const double* rcoeffs, const double* icoeffs, double* rdata,
                             double* idata);
/* device-resident form: GPU g's latitude rings [2bw/ngpu][2bw] and its full-size coefficient arrays (only the owned
 * orders are read / written) live in the plan's own buffers; `iters` transforms run back to back and
 * *ms_per_transform is the device time (CUDA events on every GPU's stream, maximum over the GPUs). */
int s2kit_cuda_multi_buffers(s2kit_cuda_multi* mp, int g, int* device, double** ring_r, double** ring_i,
                             double** coef_r, double** coef_i);
int s2kit_cuda_multi_run(s2kit_cuda_multi* mp, int inverse, int iters, double* ms_per_transform);

/* ---- tables -------------------------------------------------------------------------------------- */
/* Copies order m's table to host memory in the reference's packed layout (GenerateCosPmlTable,
 * cospml.c:161-242; TableSize(m,bw) doubles).  Memo plans only. */
int s2kit_cuda_table_export(s2kit_cuda_plan* plan, int m, double* host_out);
/* Regenerates and exports one order without a resident table (used by the Fly host API). */
int s2kit_cuda_table_generate(s2kit_cuda_plan* plan, int m, double* host_out);

/* ---- measurement --------------------------------------------------------------------------------- */
/* Per-kernel-kind CUDA-event timing on the launching stream (enable, run, synchronize, get). */
int s2kit_cuda_profile_enable(s2kit_cuda_plan* plan, int on);
int s2kit_cuda_profile_get(s2kit_cuda_plan* plan, double* ms_per_kind, long* launches_per_kind);
int s2kit_cuda_profile_reset(s2kit_cuda_plan* plan);
/* FP64 peak of this device, measured: dense DFMA and DMMA (mma.sync m8n8k4 f64) micro-kernels, TFLOP/s */
int s2kit_cuda_measure_fp64_peak(int device, double* fma_tflops, double* dmma_tflops);
/* HBM copy bandwidth (read + write bytes / s) over `bytes` of device memory, GB/s */
int s2kit_cuda_measure_copy_bw(int device, size_t bytes, double* gbs);

/* pinned host memory helpers for the host-pointer API */
void* s2kit_cuda_host_alloc(size_t bytes);
void s2kit_cuda_host_free(void* p);

const char* s2kit_cuda_last_error(void);
const char* s2kit_cuda_version(void);

#ifdef __cplusplus
}
#endif
#endif /* S2KIT_CUDA_H */
