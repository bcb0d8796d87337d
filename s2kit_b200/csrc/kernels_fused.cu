// kernels_fused.cu -- the fused per-order kernels of the batched path (bw <= 512, power of two).
//
// forward  K2+K3:  weights . DCT-II(2bw) . triangular contraction      (DLTSemi, seminaive.c:153-198, inside the
//                  m-loops of FSTSemiMemo, FST_semi_memo.c:96-108,175-201)
// inverse  K4+K5:  transposed contraction . DCT-III(2bw) . sin(theta)  (InvDLTSemi, seminaive.c:56-115, inside
//                  InvFSTSemiMemo, FST_semi_memo.c:252-266,312-332)
//
// One CTA owns one order m and NC columns = (function, +m / -m, re / im).  The cosine-domain panel that the
// unfused kernels pass through HBM (2 MiB per function and direction at bw = 256, written once and read once)
// never leaves shared memory: forward, the DCT post-processing writes straight into the MMA B-operand panel;
// inverse, the accumulators are parked in the panel and consumed by the DCT-III rounds.  Per function and
// direction this removes 4 MiB of the 17 MiB of HBM traffic and one kernel launch.
#include "s2k_fft.cuh"
#include "s2k_legendre.cuh"

namespace s2k {

constexpr int FUSED_THREADS = LEG_WARPS * 32;

// ------------------------------------------------------------------------------------------------ forward
template <int N, int NC>
__global__ void __launch_bounds__(FUSED_THREADS, 2) k_fused_fwd(
    const double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t table_shift,
    const BlockMeta* __restrict__ meta, const uint32_t* __restrict__ rt_start, const double* __restrict__ S,
    const double* __restrict__ weights, const double2* __restrict__ tw, const double2* __restrict__ qtab,
    double* __restrict__ rco, double* __restrict__ ico, long coef_stride, int nfun, int m_lo, int real_fmt) {
    constexpr int B = N / 2, T8 = N / 8, G = FUSED_THREADS / T8, NP = fft_padded_len(N);
    constexpr int NFFT = NC / 2;  // one complex FFT serves the re and im column of one (function, sign)
    static_assert(FUSED_THREADS % T8 == 0 && NFFT % G == 0, "FFT groups must tile the CTA");
    extern __shared__ double smem[];
    const int CS = panel_stride(B);
    const int m = m_lo + blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cols_per_fn = real_fmt ? 2 : 4;
    const int NF = NC / cols_per_fn;
    const int f0 = blockIdx.x * NF;
    double* Xs = smem;                                   // [2][NC][CS]  MMA B-operand panel
    double2* ex = reinterpret_cast<double2*>(smem + 2 * NC * CS);  // [G][NP] FFT exchange rows
    uint32_t* srt = reinterpret_cast<uint32_t*>(ex + G * NP);

    prefetch_order_l2(table + (order_start[m] - table_shift) * 64, order_start[m + 1] - order_start[m], tid, FUSED_THREADS);
    const BlockMeta mb0 = meta[2 * m], mb1 = meta[2 * m + 1];
    const int total = mb0.nrt + mb1.nrt;
    for (int i = tid; i < total; i += FUSED_THREADS)
        srt[i] = rt_start[(i < mb0.nrt ? mb0.rt_base : mb1.rt_base - mb0.nrt) + i];
    // panel slots beyond the bw/2 cosine indices of a parity are only ever multiplied by zero table padding,
    // but must not hold NaN garbage
    for (int i = tid; i < 2 * NC * (CS - B / 2); i += FUSED_THREADS) {
        int col2 = i / (CS - B / 2), c = B / 2 + i % (CS - B / 2);
        Xs[col2 * CS + c] = 0.0;
    }

    // ---- DCT rounds: G transforms at a time
    const int g = tid / T8, t = tid % T8;
    double2* sx = ex + g * NP;
    const double* w = weights + ((m & 1) ? N : 0);
    const double s_all = 1.0 / sqrt(2.0 * (double)N);  // seminaive.c:174
#pragma unroll 1
    for (int round = 0; round < NFFT / G; ++round) {
        const int q = round * G + g;  // transform index: columns 2q (re) and 2q+1 (im)
        const int fl = real_fmt ? q : (q >> 1), sgn = real_fmt ? 0 : (q & 1);
        const int f = f0 + fl;
        const bool live = (f < nfun) && !(sgn && m == 0);
        const int mp = sgn ? N - m : m;
        double xr[8], xi[8];
        if (live) {
            const double* Sr = S + ((long)f * 2 * N + mp) * N;
            const double* Si = Sr + (long)N * N;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int p = t + e * T8;
                int j = (p < B) ? 2 * p : 2 * (N - 1 - p) + 1;
                double wj = __ldg(w + p);  // load order (s2k_host_reordered)
                xr[e] = __ldg(Sr + j) * wj;
                xi[e] = __ldg(Si + j) * wj;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) xr[e] = xi[e] = 0.0;
        }
        fft_block<N>(xr, xi, sx, t, g, tw);
        fft_sync<N>(g);
#pragma unroll
        for (int e = 0; e < 8; ++e) sx[fft_pad(fft_out_index<N>(e, t))] = make_double2(xr[e], xi[e]);
        fft_sync<N>(g);
        double* x_re = Xs + (2 * q) * CS;  // parity 0 block; parity 1 block is NC*CS further
        for (int k = t; k < B; k += T8) {
            int nk = (N - k) & (N - 1);
            const double2 za = sx[fft_pad(k)], zb = sx[fft_pad(nk)];
            const double ar = za.x, ai = za.y, br = zb.x, bi = zb.y;
            double2 qq = __ldg(qtab + k);
            double y1 = qq.x * (ar + br) + qq.y * (ai - bi);
            double y2 = qq.x * (ai + bi) - qq.y * (ar - br);
            if (k == 0) {
                y1 *= 0.70710678118654752440;  // seminaive.c:173
                y2 *= 0.70710678118654752440;
            }
            double* dst = x_re + (k & 1) * NC * CS + (k >> 1);
            dst[0] = y1 * s_all;
            dst[CS] = y2 * s_all;
        }
        // the next round's fft_block synchronises (per transform) before it overwrites the exchange row
    }
    __syncthreads();

    // ---- contraction (same main loop as k_legendre_fwd)
    const double* tbase = table + (order_start[m] - table_shift) * 64 + lane * 2;
    const int gq = lane >> 2, q4 = lane & 3;
    const double sgn_neg = (m & 1) ? -1.0 : 1.0;
    const int base_pos = coef_base(m, B), base_neg = coef_base(-m, B);
    for (int round = 0; round * LEG_WARPS < 2 * mb0.nrt; ++round) {
        const int q = snake_item(round, warp, LEG_WARPS);
        const int p = q & 1;
        const BlockMeta mb = p ? mb1 : mb0;
        const int rt = mb0.nrt - 1 - (q >> 1);  // mb0.nrt >= mb1.nrt
        if (q >= 2 * mb0.nrt || rt >= mb.nrt) continue;
        const int ctn = tiles_in_row(mb, rt);
        const double* tp = tbase + (uint64_t)srt[(p ? mb0.nrt : 0) + rt] * 64;
        const double* xp = Xs + (p * NC + gq) * CS + q4;
        double acc[NC / 8][2];
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) acc[j][0] = acc[j][1] = 0.0;
        fwd_row_tile<NC>(tp, xp, CS, ctn, acc);

        const int r = 8 * rt + gq;
        if (r < mb.rows) {
            const int off = p + 2 * r;  // l - m
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int col = 8 * j + 2 * q4 + e;
                    int fl = col / cols_per_fn, sub = col % cols_per_fn;
                    int f = f0 + fl;
                    if (f >= nfun) continue;
                    double v = acc[j][e];
                    int part = sub & 1;
                    double* dst = (part ? ico : rco) + (long)f * coef_stride;
                    if (real_fmt) {
                        dst[base_pos + off] = v;
                        if (m > 0) dst[base_neg + off] = part ? -sgn_neg * v : sgn_neg * v;  // FST_semi_memo.c:131-145
                    } else if (sub >> 1) {
                        if (m > 0) dst[base_neg + off] = sgn_neg * v;  // FST_semi_memo.c:181-186
                    } else {
                        dst[base_pos + off] = v;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ inverse
// ITEMS = (parity, column tile) work items per warp; their accumulators stay in registers until every warp
// has finished reading the coefficient panel, which is then reused for the cosine-domain result.
template <int N, int NC>
__global__ void __launch_bounds__(FUSED_THREADS, 2) k_fused_inv(
    const double* __restrict__ table, const uint64_t* __restrict__ order_start, uint64_t table_shift,
    const BlockMeta* __restrict__ meta, const uint32_t* __restrict__ rt_start, const double* __restrict__ rco,
    const double* __restrict__ ico, long coef_stride, const double* __restrict__ sinv, const double2* __restrict__ tw,
    const double2* __restrict__ qtab, double* __restrict__ Gout, double out_scale, int nfun, int m_lo, int real_fmt) {
    constexpr int B = N / 2, T8 = N / 8, G = FUSED_THREADS / T8, NP = fft_padded_len(N);
    constexpr int NFFT = NC / 2;
    constexpr int NCT = (B / 2 + 7) / 8;                                   // column tiles per parity
    constexpr int ITEMS = (2 * NCT + LEG_WARPS - 1) / LEG_WARPS;
    constexpr int VS = B + 4;                                              // column stride of the result panel
    extern __shared__ double smem[];
    const int CS = panel_stride(B);
    const int m = m_lo + blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cols_per_fn = real_fmt ? 2 : 4;
    const int NF = NC / cols_per_fn;
    const int f0 = blockIdx.x * NF;
    double* Cs = smem;                 // [2][NC][CS] coefficient panel, later [NC][VS] cosine-domain panel
    double2* ex = reinterpret_cast<double2*>(smem + 2 * NC * CS);  // [G][NP]
    uint32_t* srt = reinterpret_cast<uint32_t*>(ex + G * NP);
    static_assert(NC * VS <= 2 * NC * (B / 2 + 4), "result panel must fit in the coefficient panel");

    prefetch_order_l2(table + (order_start[m] - table_shift) * 64, order_start[m + 1] - order_start[m], tid, FUSED_THREADS);
    const BlockMeta mb0 = meta[2 * m], mb1 = meta[2 * m + 1];
    for (int i = tid; i < mb0.nrt + mb1.nrt; i += FUSED_THREADS)
        srt[i] = rt_start[(i < mb0.nrt ? mb0.rt_base : mb1.rt_base - mb0.nrt) + i];
    const int base_pos = coef_base(m, B), base_neg = coef_base(-m, B);
    for (int col = warp; col < NC; col += LEG_WARPS) {
        int fl = col / cols_per_fn, sub = col % cols_per_fn;
        int sgn = real_fmt ? 0 : (sub >> 1), part = sub & 1;
        int f = f0 + fl;
        double* d0 = Cs + col * CS;
        double* d1 = Cs + (NC + col) * CS;
        if (f >= nfun || (sgn && m == 0)) {
            for (int c = lane; c < CS; c += 32) d0[c] = d1[c] = 0.0;
            continue;
        }
        const double* src = (part ? ico : rco) + (long)f * coef_stride + (sgn ? base_neg : base_pos);
        const int cnt = B - m, h0 = (cnt + 1) / 2, h1 = cnt / 2;
        for (int o = lane; o < cnt; o += 32) cp_async8(((o & 1) ? d1 : d0) + (o >> 1), src + o);
        for (int c = h0 + lane; c < CS; c += 32) d0[c] = 0.0;
        for (int c = h1 + lane; c < CS; c += 32) d1[c] = 0.0;
    }
    cp_async_wait_all();
    __syncthreads();

    const double* tbase = table + (order_start[m] - table_shift) * 64 + lane * 2;  // B-fragment-ordered tiles
    const int gq = lane >> 2, q4 = lane & 3;
    double acc[ITEMS][NC / 8][2];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
#pragma unroll
        for (int j = 0; j < NC / 8; ++j) acc[it][j][0] = acc[it][j][1] = 0.0;
        const int q = snake_item(it, warp, LEG_WARPS);
        if (q < 2 * NCT) {
            const int p = q & 1, ct = q >> 1;
            inv_col_tile<NC>(tbase, srt + (p ? mb0.nrt : 0), p ? mb1 : mb0, ct, Cs + (p * NC + gq) * CS + q4, CS,
                             acc[it]);
        }
    }
    __syncthreads();  // every warp is done with the coefficient panel
    double* Vs = Cs;  // [NC][VS], natural cosine index
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int q = snake_item(it, warp, LEG_WARPS);
        if (q < 2 * NCT) {
            const int p = q & 1, ct = q >> 1;
#pragma unroll
            for (int j = 0; j < NC / 8; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    int k = 2 * (8 * ct + 2 * q4 + e) + p;
                    if (k < B) Vs[(8 * j + gq) * VS + k] = acc[it][j][e];
                }
            }
        }
    }
    __syncthreads();

    // ---- DCT-III rounds (same arithmetic as k_dct_inv)
    const int g = tid / T8, t = tid % T8;
    double2* sx = ex + g * NP;
    const double c_rest = 1.0 / sqrt(2.0 * (double)N);  // 0.5/sqrt(bw), seminaive.c:72
    const double c_zero = 1.0 / sqrt((double)N);        // seminaive.c:98
#pragma unroll 1
    for (int round = 0; round < NFFT / G; ++round) {
        const int q = round * G + g;
        const int fl = real_fmt ? q : (q >> 1), sgn = real_fmt ? 0 : (q & 1);
        const int f = f0 + fl;
        const bool live = (f < nfun) && !(sgn && m == 0);
        const int mp = sgn ? N - m : m;
        const double* Va = Vs + (2 * q) * VS;
        const double* Vb = Va + VS;
        double xr[8], xi[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int k = t + e * T8;
            double wr = 0.0, wi = 0.0;
            if (k != B) {
                int src = k < B ? k : N - k;
                double sc = (src == 0) ? c_zero : c_rest;
                double a = Va[src] * sc, b = Vb[src] * sc;
                double2 qq = __ldg(qtab + k);
                double ur = (k < B) ? a : b, ui = (k < B) ? b : -a;
                wr = qq.x * ur - qq.y * ui;
                wi = qq.x * ui + qq.y * ur;
            }
            xr[e] = wi;
            xi[e] = wr;
        }
        fft_block<N>(xr, xi, sx, t, g, tw);
        if (live) {
            double sign = (sgn && (m & 1)) ? -out_scale : out_scale;  // (-1)^m for negative orders
            double* Gr = Gout + ((long)f * 2 * N + mp) * N;
            double* Gi = Gr + (long)N * N;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int i = fft_out_index<N>(e, t);
                int j = (i < B) ? 2 * i : 2 * (N - 1 - i) + 1;
                double s = (m & 1) ? __ldg(sinv + i) * sign : sign;  // output order (s2k_host_reordered)
                Gr[j] = xi[e] * s;
                Gi[j] = xr[e] * s;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
template <int N, int NC>
static size_t fused_smem() {
    constexpr int T8 = N / 8, G = FUSED_THREADS / T8;
    return sizeof(double) * (2 * NC * panel_stride(N / 2) + G * 2 * fft_padded_len(N)) + sizeof(uint32_t) * (N / 16 + 8);
}

template <int N, int NC>
static cudaError_t fused_fwd_launch(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* S,
                                    double* rco, double* ico, long coef_stride, int nfun, int m_lo, int m_hi,
                                    int real_fmt) {
    size_t smem = fused_smem<N, NC>();
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_fused_fwd<N, NC>), smem);
    if (e != cudaSuccess) return e;
    int NF = NC / (real_fmt ? 2 : 4);
    dim3 grid((nfun + NF - 1) / NF, m_hi - m_lo);
    k_fused_fwd<N, NC><<<grid, FUSED_THREADS, smem, p->stream>>>(table, p->d_order_start, shift, p->d_meta,
                                                                 p->d_rt_start, S, p->d_wv, p->d_tw_n, p->d_q_n,
                                                                 rco, ico, coef_stride, nfun, m_lo, real_fmt);
    return cudaGetLastError();
}

template <int N, int NC>
static cudaError_t fused_inv_launch(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* rco,
                                    const double* ico, long coef_stride, double* G, int nfun, int m_lo, int m_hi,
                                    int real_fmt) {
    size_t smem = fused_smem<N, NC>();
    cudaError_t e = ensure_smem(reinterpret_cast<const void*>(k_fused_inv<N, NC>), smem);
    if (e != cudaSuccess) return e;
    int NF = NC / (real_fmt ? 2 : 4);
    dim3 grid((nfun + NF - 1) / NF, m_hi - m_lo);
    k_fused_inv<N, NC><<<grid, FUSED_THREADS, smem, p->stream>>>(
        table, p->d_order_start, shift, p->d_meta, p->d_rt_start, rco, ico, coef_stride, p->d_sv, p->d_tw_n, p->d_q_n, G,
        1.0 / sqrt(2.0 * M_PI), nfun, m_lo, real_fmt);
    return cudaGetLastError();
}

// the fused path exists for power-of-two bandwidths 64..512 and batches that fill a 32-column panel
bool fused_supported(const s2kit_cuda_plan* p, int nfun, int data_format) {
    if (!p->fast || p->bw < 64 || p->bw > 512 || !p->fuse) return false;
    int cols = nfun * (data_format == S2KIT_REAL ? 2 : 4);
    return cols >= 16;
}

#define S2K_FUSED_DISPATCH(CALL)                                   \
    switch (p->n) {                                                \
        case 128: e = CALL(128, 32); break;                        \
        case 256: e = CALL(256, 32); break;                        \
        case 512: e = CALL(512, 32); break;                        \
        case 1024: e = CALL(1024, 16); break;                      \
        default: e = cudaErrorInvalidValue;                        \
    }

cudaError_t launch_fused_fwd(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* S, double* rco,
                             double* ico, long coef_stride, int nfun, int m_lo, int m_hi, int data_format) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    int real_fmt = data_format == S2KIT_REAL;
    int slot = prof_begin(p, S2KIT_K_FUSED_FWD);
    cudaError_t e;
#define CALL(NN, NCC) fused_fwd_launch<NN, NCC>(p, table, shift, S, rco, ico, coef_stride, nfun, m_lo, m_hi, real_fmt)
    S2K_FUSED_DISPATCH(CALL)
#undef CALL
    prof_end(p, slot);
    return e;
}

cudaError_t launch_fused_inv(s2kit_cuda_plan* p, const double* table, uint64_t shift, const double* rco,
                             const double* ico, long coef_stride, double* G, int nfun, int m_lo, int m_hi,
                             int data_format) {
    if (m_hi <= m_lo || nfun <= 0) return cudaSuccess;
    int real_fmt = data_format == S2KIT_REAL;
    int slot = prof_begin(p, S2KIT_K_FUSED_INV);
    cudaError_t e;
#define CALL(NN, NCC) fused_inv_launch<NN, NCC>(p, table, shift, rco, ico, coef_stride, G, nfun, m_lo, m_hi, real_fmt)
    S2K_FUSED_DISPATCH(CALL)
#undef CALL
    prof_end(p, slot);
    return e;
}

}  // namespace s2k
