// s2k_internal.cuh -- plan object and kernel launch prototypes shared by the .cu files.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/s2kit_cuda.h"

namespace s2k {

// ---- private table layout -------------------------------------------------------------------------
// Order m's cosine table T_m (rows l = m..bw-1) is split by parity p = (l-m)&1 into two
// lower-trapezoidal blocks: row r <-> l = m+p+2r, column c <-> cosine index k = 2c+p, row r has
// len0+r entries.  Each block is cut into 8x8 tiles stored row-tile-major; a tile holds 64 doubles in
// DMMA fragment order: element (row i, col j) sits at 2*(4*i + (j&3)) + (j>>2), so that lane L of a warp
// reads with ONE 128-bit load the two A-fragment values (row L/4, cols L%4 and L%4+4) of the two
// mma.m8n8k4 k-steps covering the tile.  Entries outside the trapezoid are zero.
constexpr int S2K_MAX_PEERS = 8;

struct BlockMeta {
    int rt_base;  // index of this block's first row tile in rt_start[]
    int rows;     // R_p
    int len0;     // entries in row 0
    int nrt;      // number of row tiles = ceil(rows / 8)
};

__host__ __device__ inline int tile_elem_offset(int i, int j) { return 2 * (4 * i + (j & 3)) + (j >> 2); }

// Cosine planes X / V store the bw cosine indices of a column split by parity: even k first, then odd k.  The
// contraction kernels keep the two parities in separate panels, so both sides move contiguous runs.
__host__ __device__ inline int cos_slot(int k, int bw) { return (k & 1) * ((bw + 1) / 2) + (k >> 1); }

// How a spectral plane is addressed.  The default describes the ordinary [part][order row][latitude] plane; the
// sharded single-field path points the same kernels at all-to-all send / receive blocks instead
// ([peer][part][local row][local ring], shard.cu).
struct PlaneView {
    const long* rowbase;  // K1/K6: offset (doubles) of order row m' inside the plane; null -> m' * n
    const int* rowlist;   // K2/K5: local row index -> order row m'; null -> all rows of the plane
    long part_stride;     // offset of the imaginary plane
    long lrow_stride;     // K2/K5: doubles between consecutive (local) rows
    long seg_stride;      // K2/K5: a row of 2bw latitudes is cut into segments of (seg_mask+1) entries, this far apart
    int seg_shift, seg_mask;
    int nrings;           // K1/K6: latitude rows handled by this launch
    // Latitude order inside a row.  0: natural (j).  1: the DCT kernels' even/odd-reordered order -- position i holds
    // latitude 2i (i < bw) or 2(2bw-1-i)+1 (i >= bw), i.e. exactly the sequence K2 feeds its FFT and K5's FFT emits, so
    // K2 loads / K5 stores contiguous runs.  Only the TMA variants of K1 / K6 write / read it (a tile of 8 positions is
    // 8 even or 8 odd grid rows); it is private to one fst / inv_fst call.
    int lat_perm = 0;
    // K2/K5 on a single field split over the GPUs of one process (multi.cu): segment s of a row lives in the memory of
    // peer s (PeerSegs below, passed to the PEER instantiations of the DCT kernels as a separate argument so that the
    // ordinary instantiations keep their parameters in constant memory without dynamic indexing)
    const struct PeerSegs* peers = nullptr;  // host-side only: the launchers pick the PEER kernels when set
};

struct PeerSegs {
    const double* ptr[S2K_MAX_PEERS];  // peer-mapped base of segment s (this rank's block inside peer s's ring buffer)
};

// address (doubles from the plane base of the local row) of latitude j in a row cut into segments
__host__ __device__ inline long seg_offset(const PlaneView& pv, int j) {
    return (long)(j >> pv.seg_shift) * pv.seg_stride + (j & pv.seg_mask);
}

struct ProfileSlot {
    cudaEvent_t a, b;
    int kind;
};

}  // namespace s2k

struct s2kit_cuda_plan {
    int bw = 0, n = 0, variant = 0, device = 0, chunk = 1;
    bool fast = false;  // power-of-two bandwidth >= 16: radix FFT kernels; otherwise direct O(n^2) kernels
    bool l2_persist = false;  // persisting-L2 window on Memo tables that fit (S2KIT_CUDA_L2PERSIST=1 enables)
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    int sm_count = 148;         // multiProcessorCount of `device` (persistent-grid sizes)
    // One transform at a time per plan object: every public entry point holds this lock for the whole call (the
    // workspaces and the stream are per plan).  Callers that want concurrency use one clone per thread
    // (s2kit_cuda_plan_clone: shared tables, private workspaces), which is what the drop-in layer does.
    std::mutex* mu = nullptr;
    bool shares_tables = false;  // a clone: tables / constants belong to the plan it was cloned from
    bool own_table = true;       // d_table is this object's allocation (false for Memo clones)
    // sub-batches of one device call on two streams (S2KIT_CUDA_SPLIT): HBM-bound and FP64-bound kernels of different
    // sub-batches overlap
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int nsplit = 1;
    void* host_pipe = nullptr;   // HostPipe (plan.cu): copy streams + events of the host-pointer pipeline, created once

    // sharding (single-field multi-GPU); nranks == 1 for ordinary plans
    int rank = 0, nranks = 1;
    std::vector<int> my_orders;  // orders owned by this rank (ascending)
    void* shard = nullptr;       // ShardState (shard.cu) of a sharded plan

    // host-computed constants on the device
    double* d_weights = nullptr;   // 4 bw   (weights.c:32-47)
    double* d_sin = nullptr;       // 2 bw   sin((2j+1) pi / 4bw)
    double* d_wv = nullptr;        // 4 bw   weights in the DCT kernels' even/odd-reordered load order
    double* d_sv = nullptr;        // 2 bw   sines in that order
    double2* d_tw_n = nullptr;     // n      (cos, -sin)(2 pi q / n)
    double2* d_tw_b = nullptr;     // bw     (cos, -sin)(2 pi q / bw)
    double2* d_q_n = nullptr;      // 4n     (cos, sin)(pi q / 2n)
    double2* d_q_b = nullptr;      // 4bw    (cos, sin)(pi q / 2bw)
    double* d_nodes = nullptr;     // bw     cos((2i+1) pi / 2bw)
    double* d_seeds = nullptr;     // bw*bw  seed[m][i] = P~_m^m(theta_i) (/ sin theta_i for odd m)
    double2* d_rec = nullptr;      // bw*bw  rec[m*bw + l] = (a_l^m, c_l^m)  (l2_norms.c:16-38)

    // table (private tiled layout)
    double* d_table = nullptr;    // tiles in A-fragment order (forward); Fly: the scratch ring, either order
    double* d_table_t = nullptr;  // Memo: the same tiles in B-fragment order (inverse), second half of one allocation;
                                  // == d_table when the plan keeps a single copy (table_single)
    bool table_single = false;    // bw >= 512: one copy, the inverse reads A-order tiles (S2KIT_CUDA_TABLE_COPIES=2 forces two)
    size_t table_tiles = 0;             // tiles resident in d_table
    std::vector<uint64_t> h_order_start;  // [bw+1] tile offset of each order in a full table
    std::vector<s2k::BlockMeta> h_meta;   // [2*bw]
    std::vector<uint32_t> h_rt_start;     // row-tile starts relative to the order's start
    uint64_t* d_order_start = nullptr;
    s2k::BlockMeta* d_meta = nullptr;
    uint32_t* d_rt_start = nullptr;
    // work lists of the uniform-warp kernels (kernels_uni.cu): four queues per order, heaviest unit first
    int* d_sub_off = nullptr;              // [4 bw + 1]
    unsigned short* d_sub_list = nullptr;  // forward: parity | row tile << 1 | pair << 12
    int n_sub_list = 0;
    int* d_isub_off = nullptr;
    unsigned short* d_isub_list = nullptr;  // inverse: parity | column tile << 1 | pair << 12
    int n_isub_list = 0;
    int* d_iq_off = nullptr;               // persistent inverse contraction (kernels_flow.cu)
    unsigned short* d_iq_list = nullptr;   // parity | quad << 1 (four adjacent column tiles)
    int n_iq_list = 0;
    // table-generator work units (order, first degree)
    int* d_units = nullptr;  // pairs (m, l0)
    std::vector<int> h_units;
    std::vector<int> h_unit_first;  // first unit of each order, [bw+1]
    // half-grid generator (bw >= 1024): recurrence state at the first degree of every unit, [unit][prev, cur][bw/2]
    double* d_ckpt = nullptr;
    int* d_unit_first = nullptr;
    bool own_ckpt = false;  // this object allocated them (a clone made before the first generation allocates its own)
    // Fly: scratch table for a group of orders
    size_t fly_tiles = 0;

    // workspace for `chunk` functions
    double* d_S = nullptr;  // spectral planes  [chunk][2][n][n]
    CUtensorMap tma_S;      // d_S as a 3-D tensor (latitude, order row, plane) with 8 x 256 x 1 boxes, 64-byte swizzle
    bool tma_S_ok = false;
    double* d_T = nullptr;  // ring-major staging of the longitude transforms at n = 4096 [chunk][2][rings][n]; on first use
    double* d_X = nullptr;  // cosine planes    [chunk][n][2][bw]
    double* d_coef = nullptr;  // [chunk][2][bw*bw]   conv intermediates / staging
    double* d_coef2 = nullptr;
    double* d_filt = nullptr;  // [chunk][2][bw]
    // double-buffered staging for host-pointer calls (host_pipeline, plan.cu)
    double* d_stage = nullptr;
    size_t stage_doubles = 0;
    size_t table_bytes = 0;

    // profiling
    bool profiling = false;
    std::vector<s2k::ProfileSlot> prof_slots;
    size_t prof_used = 0;
    double prof_ms[S2KIT_K_COUNT] = {0};
    long prof_launches[S2KIT_K_COUNT] = {0};
};

namespace s2k {

// Opt a kernel into > 48 KB of dynamic shared memory, once per (kernel, device): cudaFuncSetAttribute on every launch
// costs tens of microseconds of host time and made the launch path CPU-bound.
cudaError_t ensure_smem(const void* kernel, size_t bytes);

// RAII-less profiling bracket: begin returns a slot index (or -1)
int prof_begin(s2kit_cuda_plan* p, int kind);
void prof_end(s2kit_cuda_plan* p, int slot);

// the ordinary [part][order row][latitude] plane of bandwidth n/2
PlaneView default_view(int n);
// true when K1 / K6 can use their TMA variants on the plan's own spectral workspace (then lat_perm = 1 is allowed)
bool tma_planes_ok(const s2kit_cuda_plan* p, int nfun);

// ---- launchers (each checks cudaGetLastError and returns it) -----------------------------------------
// K1 / K6: longitude FFT.  S layout [f][part][order row][latitude]
cudaError_t launch_phi_fft_fwd(s2kit_cuda_plan* p, const double* rdata, const double* idata, long stride,
                               double* S, int nfun, int data_format, const PlaneView* view = nullptr);
cudaError_t launch_phi_fft_inv(s2kit_cuda_plan* p, const double* G, double* rdata, double* idata, long stride,
                               int nfun, int data_format, const PlaneView* view = nullptr);
// K2 / K5: DCT stages.  X layout [f][order row][part][cos_slot(k)]
cudaError_t launch_dct_fwd(s2kit_cuda_plan* p, const double* S, double* X, int nfun, int row_lo, int row_hi,
                           int data_format, const PlaneView* view = nullptr);
cudaError_t launch_dct_inv(s2kit_cuda_plan* p, const double* V, double* G, int nfun, int row_lo, int row_hi,
                           int data_format, const PlaneView* view = nullptr);
// K3 / K4: Legendre contraction for orders [m_lo, m_hi); table_shift = tile offset subtracted from order starts
// order_list (device, optional): the launch covers orders order_list[0 .. m_hi-m_lo) instead of [m_lo, m_hi)
cudaError_t launch_legendre_fwd(s2kit_cuda_plan* p, const double* table, uint64_t table_shift, const double* X,
                                double* rco, double* ico, long coef_stride, int nfun, int m_lo, int m_hi,
                                int data_format, const int* order_list = nullptr);
cudaError_t launch_legendre_inv(s2kit_cuda_plan* p, const double* table, uint64_t table_shift, const double* rco,
                                const double* ico, long coef_stride, double* V, int nfun, int m_lo, int m_hi,
                                int data_format, const int* order_list = nullptr);
// K7: table generation for orders [m_lo, m_hi) into `table` (tile layout, pre-zeroed by the launcher)
cudaError_t launch_table_gen(s2kit_cuda_plan* p, double* table, uint64_t table_shift, int m_lo, int m_hi,
                             int transposed = 0);
// tile layout -> reference packed layout for one order
cudaError_t launch_table_unpack(s2kit_cuda_plan* p, const double* table, uint64_t table_shift, int m, double* out);
// recurrence coefficients (a_l^m, c_l^m) for all (m, l)
cudaError_t launch_rec_coeffs(s2kit_cuda_plan* p);
// K8
cudaError_t launch_zonal_rowsum(s2kit_cuda_plan* p, const double* rdata, const double* idata, long stride, double* S,
                                int nfun);
cudaError_t launch_spectral_mul(s2kit_cuda_plan* p, const double* rd, const double* id, long coef_stride,
                                const double* rf, const double* ifl, long filt_stride, double* rres, double* ires,
                                long res_stride, int nfun);
// DLTNaive / InvDLTNaive products with a caller-provided theta-space table (all pointers device memory)
cudaError_t launch_naive_dlt(const double* data, const double* weights, const double* pml, double* result, int size,
                             int rows, cudaStream_t st);
cudaError_t launch_naive_inv_dlt(const double* coeffs, const double* pml, double* result, int size, int rows,
                                 cudaStream_t st);
int table_unit_rows(int bw);
// persistent warp-specialised K2+K3 for batched launches (kernels_pipe.cu)
bool fwd_pipe_supported(const s2kit_cuda_plan* p, int nfun, int data_format);
cudaError_t launch_fwd_pipe(s2kit_cuda_plan* p, const double* table, uint64_t table_shift, const double* S, double* rco,
                            double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format, int lat_perm);

// uniform-warp successor of the persistent kernel at bw = 256 (kernels_uni.cu); S2KIT_CUDA_UNI=0 falls back to k_fwd_pipe
bool fwd_uni_supported(const s2kit_cuda_plan* p, int nfun, int data_format);
cudaError_t launch_fwd_uni(s2kit_cuda_plan* p, const double* table, uint64_t table_shift, const double* S, double* rco,
                           double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format, int lat_perm);

bool inv_uni_supported(const s2kit_cuda_plan* p, int nfun, int data_format);
cudaError_t launch_inv_uni(s2kit_cuda_plan* p, const double* table_t, uint64_t table_shift, const double* rco,
                           const double* ico, long coef_stride, double* G, int nfun, int m_lo, int m_hi, int data_format,
                           int lat_perm);

// K4 as a persistent kernel at bw = 256, batched (kernels_flow.cu); S2KIT_CUDA_FLOW=0 falls back to k_legendre_inv
bool inv_flow_supported(const s2kit_cuda_plan* p, int nfun, int data_format);
cudaError_t launch_inv_flow(s2kit_cuda_plan* p, const double* table_t, uint64_t table_shift, const double* rco,
                            const double* ico, long coef_stride, double* V, int nfun, int m_lo, int m_hi, int data_format);

bool fwd_pipe_fused();  // default: the DCT runs inside the persistent kernel; S2KIT_CUDA_PIPE=1: K2 + streamed K3
// K3 as a persistent kernel with cp.async-streamed table tiles (kernels_pipe.cu); X = K2's cosine planes
cudaError_t launch_leg_fwd_stream(s2kit_cuda_plan* p, const double* table, uint64_t table_shift, const double* X,
                                  double* rco, double* ico, long coef_stride, int nfun, int m_lo, int m_hi,
                                  int data_format);

// K5 on the one-warp 512-point FFT (kernels_fft16.cu); S2KIT_CUDA_FFT16=0 disables
bool fft16_enabled();
cudaError_t launch_dct_inv16(s2kit_cuda_plan* p, const double* V, double* G, int nfun, int lo, int hi, const PlaneView& pv);

// peaks
cudaError_t measure_fp64(double* fma_tflops, double* dmma_tflops);
cudaError_t measure_copy(size_t bytes, double* gbs);

}  // namespace s2k
