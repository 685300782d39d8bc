// xrft_b200 -- the two FFT pass kernels and their IO policies.
//
//   rows_kernel  (K-A): batched 1-D FFT along the CONTIGUOUS axis.  One CTA = SEQ sequences.
//   cols_kernel  (K-B): batched 1-D FFT along a STRIDED axis.  One CTA = a tile of C adjacent
//                       columns x all L rows; points of a tile are interleaved in smem (SI = C).
//
// IO policies supply load() (global -> registers, with the fused prologue) and store()
// (registers/smem -> global, with the fused epilogue).  Reference call sites replaced:
//   np.fft.fftn/rfftn/ifftn/irfftn            xrft/xrft.py:439-444, 615
//   detrend subtract + window multiply        xrft/detrend.py:55,100-113 ; xrft/xrft.py:96-103
//   fftshift, phase ramp, x prod(dx)          xrft/xrft.py:446-447, 462-472
//   |F|^2, F conj(G), angle, psd scalings     xrft/xrft.py:740-748, 825-833, 865-869
//   radial-bin sum                            xrft/xrft.py:895-906
#pragma once
#include <cuda.h>
#include "fft_core.cuh"

namespace xrftb {

// register budget: keep >= 512 threads resident per SM (<= 128 registers/thread)
#ifndef XRFTB_TARGET_THREADS
#define XRFTB_TARGET_THREADS 512
#endif
constexpr int min_blocks_for(int threads) { return threads >= XRFTB_TARGET_THREADS ? 1 : (XRFTB_TARGET_THREADS / threads > 8 ? 8 : XRFTB_TARGET_THREADS / threads); }
constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n / 2); }

enum : int { EPI_COMPLEX = 0, EPI_POWER = 1, EPI_CROSS = 2, EPI_PHASE = 3, EPI_BINS_POWER = 4, EPI_BINS_CROSS = 5,
             EPI_CROSS_AND_PHASE = 6 /* internal: CROSS with the optional second output (desc.out2) */ };

__device__ __forceinline__ float xatan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double xatan2(double y, double x) { return atan2(y, x); }

// =============================================================================================
// K-A : rows
// =============================================================================================
template <typename T, int LOG2L, int LOGE, int SEQ, class IO>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * SEQ, min_blocks_for((1 << (LOG2L - LOGE)) * SEQ))
rows_kernel(IO io, const cplx<T>* __restrict__ tw, long nseq) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, NT = G_::NT;
    constexpr int SEQ_STRIDE = G_::LPAD + IO::kSeqSkew;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    const int s = threadIdx.x / NT, u = threadIdx.x % NT;
    cplx<T>* sm = smem + s * SEQ_STRIDE;
    const long ngroups = (nseq + SEQ - 1) / SEQ;
    // software pipeline: the global loads of the NEXT group are issued before the store phase of the current one
    cplx<T> raw[E];
    if ((long)blockIdx.x < ngroups) io.template fetch<LOG2L, LOGE>((long)blockIdx.x * SEQ + s, (long)blockIdx.x * SEQ + s < nseq, u, raw);
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long seq = grp * SEQ + s;
        const bool active = seq < nseq;
        const long nxt = grp + gridDim.x;
        if (threadIdx.x == 0 && nxt + gridDim.x < ngroups) io.template prefetch<LOG2L, SEQ>((nxt + gridDim.x) * SEQ, nseq);
        cplx<T> v[1][E];
        io.template prologue<LOG2L, LOGE>(seq, active, u, raw, v[0]);
        block_fft<T, LOG2L, LOGE, 1, 1>(v, u, sm, 0, tw);
        io.template store_a<LOG2L, LOGE, SEQ>(grp, seq, active, u, s, v[0], smem, SEQ_STRIDE, nseq);
        if (nxt < ngroups) io.template fetch<LOG2L, LOGE>(nxt * SEQ + s, nxt * SEQ + s < nseq, u, raw);
        io.template store_b<LOG2L, LOGE, SEQ>(grp, seq, active, u, s, v[0], smem, SEQ_STRIDE, nseq);
    }
}

// ---- plain C2C (forward, or inverse through conj(FFT(conj x)) * scale) -----------------------
template <typename T> struct RowsC2C {
    static constexpr int kSeqSkew = 0;
    const cplx<T>* in; cplx<T>* out; long in_stride, out_stride; int inverse; T scale;
    // ---- hooks of Bluestein's chirp-z on the power-of-two rows (c2c_pass; all off by default): only the first in_len points
    // of a source row exist (the rest of the row is zero), source point j is conjugated (in_conj) and multiplied by in_ramp[j];
    // output point k is multiplied by out_ramp[k], conjugated (out_conj), and stored only for k < out_len
    int in_len = 0, in_conj = 0, out_conj = 0, out_len = 0;
    const cplx<T>* in_ramp = nullptr;
    const cplx<T>* out_ramp = nullptr;

    template <int LOG2L, int SEQ> __device__ __forceinline__ void prefetch(long, long) const {}

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void fetch(long seq, bool active, int u, cplx<T> (&raw)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        const cplx<T>* p = in + seq * in_stride + u;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            raw[q] = mk<T>(0, 0);
            if (active && (in_len == 0 || u + q * NT < in_len)) raw[q] = p[q * NT];
        }
    }
    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void prologue(long, bool, int u, cplx<T> (&raw)[1 << LOGE], cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            cplx<T> x = raw[q];
            if (in_conj) x.y = -x.y;
            if (in_ramp != nullptr && (in_len == 0 || u + q * NT < in_len)) x = cmul(x, __ldg(in_ramp + u + q * NT));
            if (inverse) x.y = -x.y;
            v[q] = x;
        }
    }
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_a(long, long seq, bool active, int u, int, cplx<T> (&v)[1 << LOGE], cplx<T>*, int, long) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        if (!active) return;
        cplx<T>* p = out + seq * out_stride;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int k = final_index<LOG2L, LOGE>(u, g, t);
                if (out_len > 0 && k >= out_len) continue;
                cplx<T> x = v[g + t * G];
                if (inverse) x.y = -x.y;
                x = cscale(x, scale);
                if (out_ramp != nullptr) x = cmul(x, __ldg(out_ramp + k));
                if (out_conj) x.y = -x.y;
                p[k] = x;
            }
    }
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_b(long, long, bool, int, int, cplx<T> (&)[1 << LOGE], cplx<T>*, int, long) const {}
};

// R2C split of the packed half-length transform: Z = FFT_M(x[2n] + i x[2n+1]), N = 2M:
//   X[k] = E + w_N^k O,  E = (Z[k] + conj Z[M-k]) / 2,  O = -i (Z[k] - conj Z[M-k]) / 2,  k in [0, M]
template <typename T>
__device__ __forceinline__ cplx<T> r2c_split(cplx<T> zk, cplx<T> zm, cplx<T> w, T h = (T)0.5) {   // h = 1/2 (x a real row factor)
    cplx<T> e = mk<T>(h * (zk.x + zm.x), h * (zk.y - zm.y));
    cplx<T> d = mk<T>(h * (zk.x - zm.x), h * (zk.y + zm.y));  // (Z[k] - conj Z[M-k]) / 2
    cplx<T> o = mul_mi(d);
    return cadd(e, cmul(w, o));
}

// ---- fused real-input row pass -----------------------------------------------------------------
// prologue : (x - plane) * wy[iy] * wx[ix]       (detrend in fp64 -- SURVEY F6 -- window in T)
// epilogue : R2C split, then either the natural half-spectrum [seq][M+1]   (logC < 0)
//            or the BLOCKED intermediate [batch][tile][Ny][C] consumed by cols_kernel (logC >= 0):
//            each CTA writes SEQ*C*sizeof(cplx) contiguous bytes per tile, each column tile is
//            one contiguous Ny*C*sizeof(cplx) chunk for the column pass.
//            Ny and C are powers of two: all index math is shifts and masks.
template <typename T> struct RowsR2CFused {
    static constexpr int kSeqSkew = 4;  // rows skewed by 4 points: conflict-free cross-row gathers
    const T* in; long in_row_stride;    // real input, element stride between consecutive rows
    int logNy;                           // log2(rows per batch item) (0 for 1-D)
    int detrend;                         // 0 none | 1 constant | 2 linear
    const double* moments;               // [batch][4] : S, S0(unused), Sy, Sx -- sum and centred first moments
    const T* wy; const T* wx;            // window vectors (nullable)
    cplx<T>* out; int logC; long out_seq_stride;  // natural: stride per seq ; blocked: unused
    const cplx<T>* tw_r2c;               // exp(-2 pi i k / N), k in [0, M]
    // row-line detrend (rows2_kernel): instead of the global plane (which needs a separate moments pass over the input),
    // each row subtracts ITS OWN exactly representable line ph(j) = A0 + B j and records (A0, B, sum r, sum (j-jc) r) of
    // the residual r; the difference between the row lines and the least-squares plane is added back in the column pass
    // (ColsFused::fix), where it is a rank-2 update of the half-spectrum.  rowstats == nullptr: global-plane prologue.
    float4* rowstats;
    // ---- hooks of the padded / scaled 2-D real transform (xrftb_fft2r, natural layout only; all off by default) ----
    // xrft.pad folded into the loads: the sequences enumerate only the rows_in input rows of every item (the padding rows are
    // never transformed); real element j of a padded row is in[j - col_lo] for col_lo <= j < col_hi and zero elsewhere;
    // row r of item b lands in output row b * rows_out + row_off + r.  The half spectrum leaves multiplied by out_ramp[k] * out_scale.
    long rows_in = 0, rows_out = 0, row_off = 0;
    int col_lo = 0, col_hi = 0;
    const cplx<T>* out_ramp = nullptr;
    T out_scale = 1;

    __device__ __forceinline__ long out_row(long seq) const {
        if (rows_in == 0) return seq;
        const long b = seq / rows_in;
        return b * rows_out + row_off + (seq - b * rows_in);
    }

    template <int LOG2L, int SEQ> __device__ __forceinline__ void prefetch(long seq0, long nseq) const {
        // the SEQ input rows of an upcoming group are one contiguous block when the rows are dense (full rows, or the
        // unpadded rows of a zero-padded transform)
        const long dense = col_hi > 0 ? (long)(col_hi - col_lo) : (long)(2 << LOG2L);
        const unsigned bytes = (unsigned)(dense * sizeof(T)) * SEQ;
        if (in_row_stride == dense && seq0 + SEQ <= nseq && bytes % 16 == 0 && ((in_row_stride * sizeof(T) * seq0) % 16) == 0)
            prefetch_l2_bulk(in + seq0 * in_row_stride, bytes);
    }

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void fetch(long seq, bool active, int u, cplx<T> (&raw)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        if (col_hi > 0) {   // zero-padded row: predicated scalar loads
            const T* rowp = in + seq * in_row_stride - col_lo;
#pragma unroll
            for (int q = 0; q < (1 << LOGE); ++q) {
                const int j0 = 2 * (u + q * NT);
                raw[q] = mk<T>(0, 0);
                if (active && j0 >= col_lo && j0 < col_hi) raw[q].x = rowp[j0];
                if (active && j0 + 1 >= col_lo && j0 + 1 < col_hi) raw[q].y = rowp[j0 + 1];
            }
            return;
        }
        const cplx<T>* p = reinterpret_cast<const cplx<T>*>(in + seq * in_row_stride) + u;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            raw[q] = mk<T>(0, 0);
            if (active) raw[q] = p[q * NT];
        }
    }

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void prologue(long seq, bool active, int u, cplx<T> (&raw)[1 << LOGE], cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        constexpr int Nx = 2 << LOG2L;
        const int Ny = 1 << logNy;
        const long b = seq >> logNy;
        const int iy = (int)(seq & (Ny - 1));
        double p0 = 0.0, cx = 0.0;
        if (detrend && active) {
            const double* m = moments + b * 4;
            const double npts = (double)Ny * (double)Nx;
            p0 = m[0] / npts;
            if (detrend == 2) {
                // least-squares plane on a full regular grid == centred first moments (orthogonal regressors)
                const double vy = (double)Nx * ((double)Ny * ((double)Ny * Ny - 1.0) / 12.0);
                const double vx = (double)Ny * ((double)Nx * ((double)Nx * Nx - 1.0) / 12.0);
                const double cy = Ny > 1 ? m[2] / vy : 0.0;
                cx = m[3] / vx;
                p0 += cy * ((double)iy - 0.5 * (Ny - 1)) + cx * ((double)(2 * u) - 0.5 * (Nx - 1));
            }
        }
        const double pstep = cx * (double)(2 * NT);
        const T wrow = (wy != nullptr && active) ? wy[iy] : (T)1;
        const cplx<T>* pw = reinterpret_cast<const cplx<T>*>(wx) + u;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            cplx<T> y = raw[q];
            if (detrend) {
                const double pq = p0 + (double)q * pstep;  // independent per q: no serial fp64 dependency chain
                y.x = (T)((double)y.x - pq);
                y.y = (T)((double)y.y - (pq + cx));
            }
            if (wx != nullptr) {
                cplx<T> w = __ldg(pw + q * NT);
                y.x *= w.x * wrow; y.y *= w.y * wrow;
            } else {
                y.x *= wrow; y.y *= wrow;
            }
            v[q] = y;
        }
    }

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ cplx<T> split_at(const cplx<T>* smr, int k) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int M = G_::L;
        cplx<T> zk = smr[padded<G_::LOGPAD>(k & (M - 1))];
        cplx<T> zm = smr[padded<G_::LOGPAD>((M - k) & (M - 1))];
        return r2c_split<T>(zk, zm, __ldg(tw_r2c + k));
    }

    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_a(long, long, bool, int u, int s, cplx<T> (&v)[1 << LOGE], cplx<T>* smem, int seq_stride, long) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        cplx<T>* sm = smem + s * seq_stride;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) sm[padded<G_::LOGPAD>(final_index<LOG2L, LOGE>(u, g, t))] = v[g + t * G];
        __syncthreads();
    }

    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_b(long grp, long seq, bool active, int u, int s, cplx<T> (&)[1 << LOGE],
                                            cplx<T>* smem, int seq_stride, long nseq) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int M = G_::L, NT = G_::NT, NTHR = NT * SEQ, PADW = 1 << G_::LOGPAD;
        constexpr int LOGSEQ = ilog2c(SEQ);
        cplx<T>* sm = smem + s * seq_stride;
        if (logC < 0) {
            if (active) {
                cplx<T>* p = out + out_row(seq) * out_seq_stride;
                if (out_ramp != nullptr || out_scale != (T)1) {
                    for (int k = u; k <= M; k += NT) {
                        cplx<T> x = cscale(split_at<LOG2L, LOGE>(sm, k), out_scale);
                        if (out_ramp != nullptr) x = cmul(x, __ldg(out_ramp + k));
                        p[k] = x;
                    }
                } else {
                    for (int k = u; k <= M; k += NT) p[k] = split_at<LOG2L, LOGE>(sm, k);
                }
            }
        } else {
            // thread -> (tile t, row s2, column c), c fastest: SEQ*C consecutive threads write one contiguous run
            const int C = 1 << logC;
            const int Ny = 1 << logNy;
            const int ntile = (M >> logC) + 1;
            const int tstep = NTHR >> (logC + LOGSEQ);                // tiles covered per sweep of the CTA
            if (tstep >= 1 && ((tstep << logC) % PADW) == 0 && M % (tstep << logC) == 0) {
                // k advances by KS = tstep*C (a multiple of the pad width) per sweep: every index below is affine
                const int c = threadIdx.x & (C - 1);
                const int s2 = (threadIdx.x >> logC) & (SEQ - 1);
                const long seq2 = grp * SEQ + s2;
                if (seq2 < nseq) {
                    const long b = seq2 >> logNy;
                    const int iy = (int)(seq2 & (Ny - 1));
                    const cplx<T>* smr = smem + s2 * seq_stride;
                    const int KS = tstep << logC, KSP = KS + KS / PADW;
                    const int t0 = threadIdx.x >> (logC + LOGSEQ);
                    const int k0 = (t0 << logC) + c;                  // 0 <= k0 < KS
                    const long ostep = ((long)Ny << logC) * tstep;
                    cplx<T>* po = out + (((b * ntile + t0) << logNy) + iy) * C + c;
                    const cplx<T>* pk = smr + padded<G_::LOGPAD>(k0);
                    const cplx<T>* pm = smr + padded<G_::LOGPAD>((M - k0) & (M - 1));   // k0 == 0 -> Z[0]
                    const cplx<T>* ptw = tw_r2c + k0;
                    const int nsweep = M / KS;                         // tiles [0, M/C) : all k < M
                    {   // first sweep: k0 may be 0 (its mirror index wraps), handled by the pm above
                        *po = r2c_split<T>(*pk, *pm, __ldg(ptw));
                        pm = smr + padded<G_::LOGPAD>(M - k0 - KS > 0 ? M - k0 - KS : 0);
                        pk += KSP; ptw += KS; po += ostep;
                    }
#pragma unroll 4
                    for (int i = 1; i < nsweep; ++i) {
                        *po = r2c_split<T>(*pk, *pm, __ldg(ptw));
                        pk += KSP; pm -= KSP; ptw += KS; po += ostep;
                    }
                    // last tile (t = M/C): column M (Nyquist) then zero padding
                    if (t0 == 0) {
                        cplx<T> z0 = smr[0];
                        cplx<T>* pl = out + (((b * ntile + (M >> logC)) << logNy) + iy) * C + c;
                        *pl = (c == 0) ? mk<T>(z0.x - z0.y, 0) : mk<T>(0, 0);
                    }
                }
            } else {
                for (int w = threadIdx.x; w < ntile * SEQ * C; w += NTHR) {
                    const int c = w & (C - 1);
                    const int s2 = (w >> logC) & (SEQ - 1);
                    const int t = w >> (logC + LOGSEQ);
                    const long seq2 = grp * SEQ + s2;
                    if (seq2 >= nseq) continue;
                    const long b = seq2 >> logNy;
                    const int iy = (int)(seq2 & (Ny - 1));
                    const int k = (t << logC) + c;
                    out[(((b * ntile + t) << logNy) + iy) * C + c] =
                        (k <= M) ? split_at<LOG2L, LOGE>(smem + s2 * seq_stride, k) : mk<T>(0, 0);
                }
            }
        }
        __syncthreads();
    }
};

// ---- fused real-input row pass, TWO rows per thread (the hot 2-D path) ------------------------------------
// Same math as rows_kernel<RowsR2CFused> in blocked mode; each thread owns the same 16 points of two adjacent
// rows, interleaved in shared memory ([pad(o)][2]) so every exchange access is one 128-bit LDS/STS for both
// rows and the stage twiddles are generated once per pair.  Needs Ny % (2*PAIRS) == 0 (rows of a CTA group are
// consecutive rows of one item).
template <typename T, int LOG2L, int LOGE, int PAIRS, bool ROWLINE>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * PAIRS, min_blocks_for((1 << (LOG2L - LOGE)) * PAIRS))
rows2_kernel(RowsR2CFused<T> io, const cplx<T>* __restrict__ tw, long nseq) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, NT = G_::NT, SEQ = 2 * PAIRS, NTHR = NT * PAIRS, M = G_::L, Nx = 2 * M;
    constexpr int PAIR_STRIDE = 2 * G_::LPAD + 8;
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R, PADW = 1 << G_::LOGPAD, LOGSEQ = ilog2c(SEQ);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    const int pr = threadIdx.x / NT, u = threadIdx.x % NT;
    cplx<T>* sm = smem + pr * PAIR_STRIDE;
    const long ngroups = nseq / SEQ;
    const int Ny = 1 << io.logNy;
    const int logC = io.logC, C = 1 << logC;
    const int ntile = (M >> logC) + 1;
    // two CTAs per SM interleave their load / transform / store phases; the next group is L2-prefetched
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long nxt = grp + gridDim.x;
        if (threadIdx.x == 0 && nxt < ngroups) io.template prefetch<LOG2L, SEQ>(nxt * SEQ, nseq);
        cplx<T> v[2][E];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const cplx<T>* p = reinterpret_cast<const cplx<T>*>(io.in + (grp * SEQ + 2 * pr + r) * io.in_row_stride) + u;
#pragma unroll
            for (int q = 0; q < E; ++q) v[r][q] = p[q * NT];
        }
        // ---- prologue (row-line variant): fp32 line subtract with exact line values, residual sums, window
        if constexpr (ROWLINE) {
            if constexpr (sizeof(T) == 4) {
            constexpr int RL = NT < 32 ? NT : 32, WPP = NT / RL;      // lanes per reduction segment, partial slots per pair
            float* part = reinterpret_cast<float*>(smem + PAIRS * PAIR_STRIDE);   // [PAIRS][WPP][4] partial sums, then [PAIRS][4] lines
            float* lines = part + PAIRS * WPP * 4;
            const long seq0 = grp * SEQ + 2 * pr;
            const float jf0 = (float)(2 * u);
            const float jc = 0.5f * (float)(Nx - 1);
            const cplx<T>* pw = reinterpret_cast<const cplx<T>*>(io.wx) + u;
            // residual sums per row: sx, sy (even / odd columns) and the q-weighted tx, ty; column j = 2u + 2 NT q (+1), so
            //   sum r = sx + sy ;  sum (j - jc) r = (2u - jc) sx + (2u + 1 - jc) sy + 2 NT (tx + ty)
            // The row's window factor w_y(i) is NOT applied here: the R2C split folds it into its 1/2 (store phase).
            float A0[2], B[2], sx[2] = {0.f, 0.f}, sy[2] = {0.f, 0.f}, tx[2] = {0.f, 0.f}, ty[2] = {0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float* rowp = io.in + (seq0 + r) * io.in_row_stride;
                const float x0 = __ldg(rowp), x1 = __ldg(rowp + (Nx - 1));
                const float m = fmaxf(fabsf(x0), fabsf(x1));
                A0[r] = 0.f; B[r] = 0.f;
                if (m > 1e-30f && m < 1e30f) {
                    // quantum Q = 2^(floor(log2 m) - 21): A0, B are multiples of Q and |A0 + B j| < 2^(floor(log2 m) + 2),
                    // so every line value is representable: fmaf(B, j, A0) and (that + B) return them exactly
                    const int p2b = __float_as_int(m) & 0x7f800000;
                    const float p2 = __int_as_float(p2b);
                    const float Q = p2 * 4.76837158203125e-07f, iQ = __int_as_float((275 << 23) - p2b);   // 2^21 / p2, exactly
                    A0[r] = rintf(x0 * iQ) * Q;
                    B[r] = rintf((x1 - x0) * (1.0f / (float)(Nx - 1)) * iQ) * Q;
                }
            }
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const float jf = jf0 + (float)(2 * NT * q);
                cplx<T> w = mk<T>(1, 1);
                if (io.wx != nullptr) w = __ldg(pw + q * NT);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    cplx<T> y = v[r][q];
                    const float ph = fmaf(B[r], jf, A0[r]);
                    y.x -= ph;
                    y.y -= ph + B[r];
                    sx[r] += y.x; sy[r] += y.y;
                    tx[r] = fmaf((float)q, y.x, tx[r]);
                    ty[r] = fmaf((float)q, y.y, ty[r]);
                    y.x *= w.x; y.y *= w.y;
                    v[r][q] = y;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float s2 = sx[r] + sy[r];
                float t2 = fmaf(jf0 - jc, sx[r], fmaf(jf0 + 1.f - jc, sy[r], (float)(2 * NT) * (tx[r] + ty[r])));
#pragma unroll
                for (int off = RL / 2; off > 0; off >>= 1) {
                    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, off);
                }
                if ((u & (RL - 1)) == 0) {
                    part[(pr * WPP + u / RL) * 4 + 2 * r] = s2;
                    part[(pr * WPP + u / RL) * 4 + 2 * r + 1] = t2;
                }
                if (u == 0) { lines[pr * 4 + 2 * r] = A0[r]; lines[pr * 4 + 2 * r + 1] = B[r]; }
            }
            }
        } else
        // ---- prologue: fp64 plane subtract, window
        {
            const long seq0 = grp * SEQ + 2 * pr;
            const long b = seq0 >> io.logNy;
            const int iy0 = (int)(seq0 & (Ny - 1));
            double pm = 0.0, cx = 0.0, cy = 0.0;
            if (io.detrend) {
                const double* m = io.moments + b * 4;
                const double npts = (double)Ny * (double)Nx;
                pm = m[0] / npts;
                if (io.detrend == 2) {
                    const double vy = (double)Nx * ((double)Ny * ((double)Ny * Ny - 1.0) / 12.0);
                    const double vx = (double)Ny * ((double)Nx * ((double)Nx * Nx - 1.0) / 12.0);
                    cy = Ny > 1 ? m[2] / vy : 0.0;
                    cx = m[3] / vx;
                    pm += cx * ((double)(2 * u) - 0.5 * (Nx - 1));
                }
            }
            const double pstep = cx * (double)(2 * NT);
            const cplx<T>* pw = reinterpret_cast<const cplx<T>*>(io.wx) + u;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double p0 = pm + cy * ((double)(iy0 + r) - 0.5 * (Ny - 1));
                const T wrow = io.wy != nullptr ? io.wy[iy0 + r] : (T)1;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    cplx<T> y = v[r][q];
                    if (io.detrend) {
                        const double pq = p0 + (double)q * pstep;
                        y.x = (T)((double)y.x - pq);
                        y.y = (T)((double)y.y - (pq + cx));
                    }
                    if (io.wx != nullptr) {
                        cplx<T> w = __ldg(pw + q * NT);
                        y.x *= w.x * wrow; y.y *= w.y * wrow;
                    } else {
                        y.x *= wrow; y.y *= wrow;
                    }
                    v[r][q] = y;
                }
            }
        }
        block_fft<T, LOG2L, LOGE, 2, 2>(v, u, sm, 1, tw);
        // ---- store_a: both rows' packed spectra Z -> smem (natural order, interleaved)
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                cplx<T>* p = sm + padded<G_::LOGPAD>(final_index<LOG2L, LOGE>(u, g, t)) * 2;
                p[0] = v[0][g + t * G];
                p[1] = v[1][g + t * G];
            }
        __syncthreads();
        if constexpr (ROWLINE) {
            if constexpr (sizeof(T) == 4) {
                // the partial sums became visible at the barrier above; one thread per row folds them
                constexpr int RL = NT < 32 ? NT : 32, WPP = NT / RL;
                const float* part = reinterpret_cast<const float*>(smem + PAIRS * PAIR_STRIDE);
                const float* lines = part + PAIRS * WPP * 4;
                if (threadIdx.x < SEQ) {
                    const int p_ = threadIdx.x >> 1, r_ = threadIdx.x & 1;
                    float a = 0.f, b_ = 0.f;
#pragma unroll
                    for (int w = 0; w < WPP; ++w) { a += part[(p_ * WPP + w) * 4 + 2 * r_]; b_ += part[(p_ * WPP + w) * 4 + 2 * r_ + 1]; }
                    io.rowstats[grp * SEQ + threadIdx.x] = make_float4(lines[p_ * 4 + 2 * r_], lines[p_ * 4 + 2 * r_ + 1], a, b_);
                }
            }
        }
        // ---- store_b: R2C split + blocked store; thread -> (tile t, row s2, column c), c fastest
        {
            const int tstep = NTHR >> (logC + LOGSEQ);
            const int c = threadIdx.x & (C - 1);
            const int s2 = (threadIdx.x >> logC) & (SEQ - 1);
            const long seq2 = grp * SEQ + s2;
            const long b = seq2 >> io.logNy;
            const int iy = (int)(seq2 & (Ny - 1));
            const cplx<T>* smr = smem + (s2 >> 1) * PAIR_STRIDE + (s2 & 1);
            // row-line variant: the row's window factor w_y(iy) rides on the 1/2 of the split
            T hrow = (T)0.5;
            if constexpr (ROWLINE) { if (io.wy != nullptr) hrow *= io.wy[iy]; }
            if (tstep >= 1 && ((tstep << logC) % PADW) == 0 && M % (tstep << logC) == 0) {
                const int KS = tstep << logC, KSP2 = 2 * (KS + KS / PADW);
                const int t0 = threadIdx.x >> (logC + LOGSEQ);
                const int k0 = (t0 << logC) + c;
                const long ostep = ((long)Ny << logC) * tstep;
                cplx<T>* po = io.out + (((b * ntile + t0) << io.logNy) + iy) * C + c;
                const cplx<T>* pk = smr + 2 * padded<G_::LOGPAD>(k0);
                const cplx<T>* pmr = smr + 2 * padded<G_::LOGPAD>((M - k0) & (M - 1));
                const cplx<T>* ptw = io.tw_r2c + k0;
                const int nsweep = M / KS;
                *po = r2c_split<T>(*pk, *pmr, __ldg(ptw), hrow);
                pmr = smr + 2 * padded<G_::LOGPAD>(M - k0 - KS > 0 ? M - k0 - KS : 0);
                pk += KSP2; ptw += KS; po += ostep;
#pragma unroll 4
                for (int i = 1; i < nsweep; ++i) {
                    *po = r2c_split<T>(*pk, *pmr, __ldg(ptw), hrow);
                    pk += KSP2; pmr -= KSP2; ptw += KS; po += ostep;
                }
                if (t0 == 0) {
                    cplx<T> z0 = smr[0];
                    cplx<T>* pl = io.out + (((b * ntile + (M >> logC)) << io.logNy) + iy) * C + c;
                    *pl = (c == 0) ? mk<T>(((T)2 * hrow) * (z0.x - z0.y), 0) : mk<T>(0, 0);
                }
            } else {
                for (int w = threadIdx.x; w < ntile * SEQ * C; w += NTHR) {
                    const int c2 = w & (C - 1);
                    const int s3 = (w >> logC) & (SEQ - 1);
                    const int t = w >> (logC + LOGSEQ);
                    const long seq3 = grp * SEQ + s3;
                    const long b3 = seq3 >> io.logNy;
                    const int iy3 = (int)(seq3 & (Ny - 1));
                    const int k = (t << logC) + c2;
                    const cplx<T>* sr = smem + (s3 >> 1) * PAIR_STRIDE + (s3 & 1);
                    cplx<T> r = mk<T>(0, 0);
                    T h3 = (T)0.5;
                    if constexpr (ROWLINE) { if (io.wy != nullptr) h3 *= io.wy[iy3]; }
                    if (k <= M) r = r2c_split<T>(sr[2 * padded<G_::LOGPAD>(k & (M - 1))], sr[2 * padded<G_::LOGPAD>((M - k) & (M - 1))], __ldg(io.tw_r2c + k), h3);
                    io.out[(((b3 * ntile + t) << io.logNy) + iy3) * C + c2] = r;
                }
            }
        }
        __syncthreads();
    }
}

// ---- C2R: half spectrum [seq][M+1] -> real [seq][N], numpy irfft semantics (scale = 2/N folded in) --
//   Z[k] = E + i O,  E = (X[k] + conj X[M-k]) / 2,  O = w_N^{-k} (X[k] - conj X[M-k]) / 2
//   z = IFFT_M(Z) = conj(FFT(conj Z)) ;  x[2n] = Re z[n], x[2n+1] = Im z[n]
template <typename T> struct RowsC2R {
    static constexpr int kSeqSkew = 0;
    const cplx<T>* in; long in_stride; T* out; long out_stride; T scale; const cplx<T>* tw_r2c;
    // ---- hooks of xrftb_fft2r (all off by default): the half spectrum is multiplied by in_ramp[k] on its way in (before the
    // imaginary parts of DC and Nyquist are dropped, like spectral_post -> irfftn); real output element n is stored at position
    // (n + out_roll) % N; with out_hi > 0 only positions in [out_lo, out_hi) are written, at out[pos - out_lo] (crop)
    const cplx<T>* in_ramp = nullptr;
    int out_roll = 0, out_lo = 0, out_hi = 0;

    // the half-spectrum rows of an upcoming group -> L2 (the loads of a row happen at the start of its transform)
    template <int LOG2L, int SEQ> __device__ __forceinline__ void prefetch(long seq0, long nseq) const {
        const unsigned bytes = (unsigned)(in_stride * sizeof(cplx<T>)) * SEQ;
        if (seq0 + SEQ <= nseq && bytes % 16 == 0 && ((in_stride * sizeof(cplx<T>) * seq0) % 16) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0)
            prefetch_l2_bulk(in + seq0 * in_stride, bytes);
    }

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void fetch(long, bool, int, cplx<T> (&)[1 << LOGE]) const {}
    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void prologue(long seq, bool active, int u, cplx<T> (&)[1 << LOGE], cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, M = 1 << LOG2L;
        const cplx<T>* p = in + seq * in_stride;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            const int k = u + q * NT;
            cplx<T> z = mk<T>(0, 0);
            if (active) {
                cplx<T> xk = p[k], xm = p[M - k];
                if (in_ramp != nullptr) { xk = cmul(xk, __ldg(in_ramp + k)); xm = cmul(xm, __ldg(in_ramp + (M - k))); }
                if (k == 0) { xk.y = 0; xm.y = 0; }  // numpy/pocketfft ignore Im of DC and Nyquist
                cplx<T> e = mk<T>((T)0.5 * (xk.x + xm.x), (T)0.5 * (xk.y - xm.y));
                cplx<T> d = mk<T>((T)0.5 * (xk.x - xm.x), (T)0.5 * (xk.y + xm.y));
                cplx<T> o = cmulc(d, __ldg(tw_r2c + k));  // * w_N^{-k}
                z = cadd(e, mul_pi(o));
                z.y = -z.y;  // conj for the inverse-through-forward trick
            }
            v[q] = z;
        }
    }
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_b(long, long, bool, int, int, cplx<T> (&)[1 << LOGE], cplx<T>*, int, long) const {}
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_a(long, long seq, bool active, int u, int, cplx<T> (&v)[1 << LOGE], cplx<T>*, int, long) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        if (!active) return;
        T* p = out + seq * out_stride;
        if (out_roll != 0 || out_hi > 0) {
            constexpr int N = 2 << LOG2L;
            const int lo = out_hi > 0 ? out_lo : 0, hi = out_hi > 0 ? out_hi : N;
            const bool pairs = ((out_roll | lo) & 1) == 0 && (hi & 1) == 0 && (out_stride & 1) == 0;   // pairs stay aligned pairs
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const cplx<T> x = v[g + t * G];
                    const int pos = (2 * final_index<LOG2L, LOGE>(u, g, t) + out_roll) & (N - 1);
                    if (pairs) {
                        if (pos >= lo && pos < hi) *reinterpret_cast<cplx<T>*>(p + (pos - lo)) = mk<T>(x.x * scale, -x.y * scale);
                    } else {
                        const int pos1 = (pos + 1) & (N - 1);
                        if (pos >= lo && pos < hi) p[pos - lo] = x.x * scale;
                        if (pos1 >= lo && pos1 < hi) p[pos1 - lo] = -x.y * scale;
                    }
                }
            return;
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                cplx<T> x = v[g + t * G];
                cplx<T> r = mk<T>(x.x * scale, -x.y * scale);
                *reinterpret_cast<cplx<T>*>(p + 2 * final_index<LOG2L, LOGE>(u, g, t)) = r;
            }
    }
};

// pass 2 of the columns-first order, TWO rows per thread (like rows2_kernel): each thread owns the same E points of two
// consecutive half-spectrum rows, interleaved in shared memory ([pad(o)][2]) so that every exchange is one 128-bit
// access for both rows, the stage twiddles are generated once per pair and the completion vector ag is loaded once.
// Epilogue straight from registers: |F|^2 * scale to output row ky and, reversed, to row -ky (fully coalesced).
template <typename T> struct RowsC2CPower;
template <typename T, int LOG2L, int LOGE, int PAIRS>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * PAIRS, min_blocks_for((1 << (LOG2L - LOGE)) * PAIRS))
rows2c_power_kernel(RowsC2CPower<T> io, const cplx<T>* __restrict__ tw, long nseq) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, NT = G_::NT, SEQ = 2 * PAIRS, Nx = 1 << LOG2L;
    constexpr int PAIR_STRIDE = 2 * G_::LPAD + 8;
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    const int pr = threadIdx.x / NT, u = threadIdx.x % NT;
    cplx<T>* sm = smem + pr * PAIR_STRIDE;
    const long ngroups = (nseq + SEQ - 1) / SEQ;
    const int Ny = 1 << io.logNy;
    const int sy = io.shift_y ? Ny / 2 : 0, sx = io.shift_x ? Nx / 2 : 0;
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long nxt = grp + gridDim.x;
        if (threadIdx.x == 0 && nxt < ngroups) io.template prefetch<LOG2L, SEQ>(nxt * SEQ, nseq);
        const long seq0 = grp * SEQ + 2 * pr;
        cplx<T> v[2][E];
        long bb[2]; int kys[2]; bool act[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            act[r] = seq0 + r < nseq;
            const cplx<T>* p = io.in + (seq0 + r) * (long)Nx + u;
#pragma unroll
            for (int q = 0; q < E; ++q) v[r][q] = act[r] ? p[q * NT] : mk<T>(0, 0);
            bb[r] = (seq0 + r) / io.H;
            kys[r] = (int)((seq0 + r) - bb[r] * io.H);
        }
        if (io.ag != nullptr) {
            // completion of the column-line detrend; both rows usually belong to the same item: one ag load serves both
            cplx<T> W[2], J[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) { W[r] = mk<T>(0, 0); J[r] = mk<T>(0, 0); if (act[r]) { W[r] = __ldg(io.wj + 2 * kys[r]); J[r] = __ldg(io.wj + 2 * kys[r] + 1); } }
            const cplx<T>* pa0 = io.ag + (bb[0] << LOG2L) + u;
            const bool same = bb[1] == bb[0];
            const cplx<T>* pa1 = io.ag + ((act[1] ? bb[1] : bb[0]) << LOG2L) + u;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const cplx<T> a0 = __ldg(pa0 + q * NT);
                const cplx<T> a1 = same ? a0 : __ldg(pa1 + q * NT);
                v[0][q].x += a0.x * W[0].x + a0.y * J[0].x; v[0][q].y += a0.x * W[0].y + a0.y * J[0].y;
                v[1][q].x += a1.x * W[1].x + a1.y * J[1].x; v[1][q].y += a1.x * W[1].y + a1.y * J[1].y;
            }
        }
        block_fft<T, LOG2L, LOGE, 2, 2>(v, u, sm, 1, tw);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!act[r]) continue;
            const int ky = kys[r];
            T* rowd = io.out + ((bb[r] << io.logNy) + ((ky + sy) & (Ny - 1))) * (long)Nx;
            T* rowm = io.out + ((bb[r] << io.logNy) + ((Ny - ky + sy) & (Ny - 1))) * (long)Nx;
            const bool self = (ky == 0) || (2 * ky == Ny);   // see RowsC2CPower::store_a
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const cplx<T> f = v[r][g + t * G];
                    const T val = (f.x * f.x + f.y * f.y) * io.scale;
                    const int kx = final_index<LOG2L, LOGE>(u, g, t);
                    if (!self || 2 * kx <= Nx) rowd[(kx + sx) & (Nx - 1)] = val;
                    if (!self || (kx > 0 && 2 * kx < Nx)) rowm[(Nx - kx + sx) & (Nx - 1)] = val;
                }
        }
        __syncthreads();   // the exchange buffer is reused by the next group
    }
}

// Pass 2 of the columns-first order when pass 1 leaves the PACKED column spectra Z[batch][Ny][M] (M = Nx/2 packed columns,
// ColsR2CPack::zout): row ky of the half-spectrum is A_c[ky] = Z[ky][c] + conj Z[Ny-ky][c] (real column 2c) and
// B_c[ky] = -i (Z[ky][c] - conj Z[Ny-ky][c]) (real column 2c+1) -- the factors 1/2 and w_x were applied in pass 1 -- and
// both land in the thread that loaded (Z[ky][c], Z[Ny-ky][c]): the separation costs four additions and no exchange.
// The row transform of the interleaved sequence (A_0, B_0, A_1, B_1, ...) of length Nx is computed as two M-point
// transforms (one block_fft call, the two sequences interleaved in shared memory) and a final radix-2 step in registers:
// F[k] = FA[k] + w^k FB[k], F[k + M] = FA[k] - w^k FB[k], w = exp(-2 pi i / Nx) (table staged in shared memory once per CTA).
// ROWS half-spectrum rows per CTA, NT = M / E threads each.
template <typename T> struct RowsZPower {
    const cplx<T>* z; T* out; int logNy; int H; int shift_y, shift_x; T scale;
    const cplx<T>* ag; const cplx<T>* wj;   // column-line detrend completion (see RowsC2CPower); nullptr = none
    const cplx<T>* tw2;                     // exp(-2 pi i k / Nx), k in [0, Nx): twiddle_fft(log2 Nx)
};
template <typename T, int LOG2M, int LOGE, int ROWS>
__global__ void __launch_bounds__((1 << (LOG2M - LOGE)) * ROWS, min_blocks_for((1 << (LOG2M - LOGE)) * ROWS))
rowsz_power_kernel(RowsZPower<T> io, const cplx<T>* __restrict__ tw, long nseq) {
    using G_ = Geometry<LOG2M, LOGE>;
    constexpr int E = G_::E, NT = G_::NT, M = 1 << LOG2M, Nx = 2 * M;
    constexpr int ROW_STRIDE = 2 * G_::LPAD + 8;
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* smw = smem + ROWS * ROW_STRIDE;   // [M] radix-2 twiddles
    const int r = threadIdx.x / NT, u = threadIdx.x % NT;
    cplx<T>* sm = smem + r * ROW_STRIDE;
    for (int k = threadIdx.x; k < M; k += NT * ROWS) smw[k] = __ldg(io.tw2 + k);
    __syncthreads();
    const long ngroups = (nseq + ROWS - 1) / ROWS;
    const int Ny = 1 << io.logNy;
    const int sy = io.shift_y ? Ny / 2 : 0, sx = io.shift_x ? Nx / 2 : 0;
    // (item, row) of this thread's sequence, advanced by the grid stride without a division per group
    const long gstep = (long)gridDim.x * ROWS;
    const long bstep = gstep / io.H;
    const int kstep = (int)(gstep - bstep * io.H);
    long seq = (long)blockIdx.x * ROWS + r;
    long b = seq / io.H;
    int ky = (int)(seq - b * io.H);
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, seq += gstep, b += bstep, ky += kstep) {
        if (ky >= io.H) { ky -= io.H; ++b; }
        const long nxt = grp + gridDim.x;
        if (threadIdx.x < 2 * ROWS && nxt < ngroups) {   // next group's rows -> L2
            const long s2 = nxt * ROWS + (threadIdx.x >> 1);
            if (s2 < nseq) {
                const long b2 = s2 / io.H;
                const int k2 = (int)(s2 - b2 * io.H);
                const int row = (threadIdx.x & 1) ? ((Ny - k2) & (Ny - 1)) : k2;
                prefetch_l2_bulk(io.z + ((b2 << io.logNy) + row) * (long)M, (unsigned)(M * sizeof(cplx<T>)));
            }
        }
        const bool act = seq < nseq;
        cplx<T> v[2][E];
        {
            const cplx<T>* pa = io.z + ((b << io.logNy) + ky) * (long)M + u;
            const cplx<T>* pb = io.z + ((b << io.logNy) + ((Ny - ky) & (Ny - 1))) * (long)M + u;
            cplx<T> za[E], zb[E];
#pragma unroll
            for (int q = 0; q < E; ++q) { za[q] = act ? pa[q * NT] : mk<T>(0, 0); zb[q] = act ? pb[q * NT] : mk<T>(0, 0); }
#pragma unroll
            for (int q = 0; q < E; ++q) {
                v[0][q] = mk<T>(za[q].x + zb[q].x, za[q].y - zb[q].y);
                v[1][q] = mk<T>(za[q].y + zb[q].y, zb[q].x - za[q].x);
            }
        }
        if (io.ag != nullptr && act) {
            const cplx<T> W = __ldg(io.wj + 2 * ky), J = __ldg(io.wj + 2 * ky + 1);
            const cplx<T>* pa = io.ag + b * (long)Nx + 2 * u;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                cplx<T> a0, a1;
                if constexpr (sizeof(T) == 4) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(pa + 2 * q * NT));
                    a0 = mk<T>(a.x, a.y); a1 = mk<T>(a.z, a.w);
                } else {
                    a0 = __ldg(pa + 2 * q * NT); a1 = __ldg(pa + 2 * q * NT + 1);
                }
                v[0][q].x += a0.x * W.x + a0.y * J.x; v[0][q].y += a0.x * W.y + a0.y * J.y;
                v[1][q].x += a1.x * W.x + a1.y * J.x; v[1][q].y += a1.x * W.y + a1.y * J.y;
            }
        }
        block_fft<T, LOG2M, LOGE, 2, 2>(v, u, sm, 1, tw);
        if (act) {
            T* rowd = io.out + ((b << io.logNy) + ((ky + sy) & (Ny - 1))) * (long)Nx;
            T* rowm = io.out + ((b << io.logNy) + ((Ny - ky + sy) & (Ny - 1))) * (long)Nx;
            // rows 0 and Ny/2 are their own mirror image (rowm == rowd): their kx <= Nx/2 half is written and mirrored inside
            // the row, so the result is exactly symmetric like every other row pair
            const bool self = (ky == 0) || (2 * ky == Ny);
            if (!self) {
                // k = u + c with c = g NT + t Ns known at compile time and k < M, sx in {0, M}: the four destinations are
                // base pointers plus immediate offsets (no index arithmetic per point).  Only k = 0 wraps in a mirrored row.
                constexpr int NsL = 1 << G_::LOGNS_LAST;
                T* pd0 = rowd + u + sx;                         // kx = k       at (k + sx) mod Nx
                T* pd1 = rowd + u + ((M + sx) & (Nx - 1));      // kx = k + M   at (k + M + sx) mod Nx
                T* pm0 = rowm + ((sx ? M : Nx) - u);            // mirror of kx = k      at (Nx - k + sx) mod Nx = pm0[-c]
                T* pm1 = rowm + ((sx ? Nx : M) - u);            // mirror of kx = k + M  at (M - k + sx) mod Nx  = pm1[-c]
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int t = 0; t < R; ++t) {
                        const int c = g * NT + t * NsL;
                        const cplx<T> fa = v[0][g + t * G];
                        const cplx<T> wb = cmul(v[1][g + t * G], smw[u + c]);
                        const cplx<T> f0 = cadd(fa, wb), f1 = csub(fa, wb);
                        const T p0 = (f0.x * f0.x + f0.y * f0.y) * io.scale;
                        const T p1 = (f1.x * f1.x + f1.y * f1.y) * io.scale;
                        pd0[c] = p0;
                        pd1[c] = p1;
                        if (c == 0) {   // k = u: u = 0 is the one point whose mirror index wraps
                            rowm[(Nx - u + sx) & (Nx - 1)] = p0;
                            rowm[(M - u + sx) & (Nx - 1)] = p1;
                        } else {
                            pm0[-c] = p0;
                            pm1[-c] = p1;
                        }
                    }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int t = 0; t < R; ++t) {
                        const int k = final_index<LOG2M, LOGE>(u, g, t);
                        const cplx<T> fa = v[0][g + t * G];
                        const cplx<T> wb = cmul(v[1][g + t * G], smw[k]);
                        const cplx<T> f0 = cadd(fa, wb), f1 = csub(fa, wb);
                        const T p0 = (f0.x * f0.x + f0.y * f0.y) * io.scale;   // kx = k
                        const T p1 = (f1.x * f1.x + f1.y * f1.y) * io.scale;   // kx = k + M
                        rowd[(k + sx) & (Nx - 1)] = p0;
                        if (k > 0) rowm[(Nx - k + sx) & (Nx - 1)] = p0;
                        if (k == 0) rowd[(k + M + sx) & (Nx - 1)] = p1;
                    }
            }
        }
    }
}

// Two-field pass 2 of the columns-first z-mode chain (BASELINE config 3): cross spectrum F1 conj(F2), its phase, or BOTH from
// one read of the packed column spectra Z1, Z2 of the two fields (each written by the unchanged pass 1).  Measured on B200
// (2048^2 x 64): 117 / 112 GPoints/s per field against 78 / 82 of the rows-first two-field chain.
// One group of NT threads per (row, field): it separates the real columns of its field's rows ky / Ny-ky, transforms the
// row exactly like rowsz_power_kernel and leaves F_f[0 .. Nx) in its own (by then free) exchange buffer; after one
// barrier the 2 NT threads of the row combine F1 conj(F2) cooperatively and write rows ky and -ky (conjugate / negated
// angle) with coalesced stores.
template <typename T> struct RowsZCross {
    const cplx<T>* z1; const cplx<T>* z2; void* out; void* out2 /* phase, EPI_CROSS_AND_PHASE */; int logNy; int H; int shift_y, shift_x; T scale;
    const cplx<T>* ag1; const cplx<T>* ag2; const cplx<T>* wj;   // column-line detrend completion per field (nullptr = none)
    const cplx<T>* tw2;                                          // exp(-2 pi i k / Nx), k in [0, Nx)
};
template <int LOG2M, int LOGE, int ROWS, int MODE>
__global__ void __launch_bounds__((1 << (LOG2M - LOGE)) * 2 * ROWS, min_blocks_for((1 << (LOG2M - LOGE)) * 2 * ROWS))
rowszx_kernel(RowsZCross<float> io, const float2* __restrict__ tw, long nseq) {
    using T = float;
    using G_ = Geometry<LOG2M, LOGE>;
    constexpr int E = G_::E, NT = G_::NT, M = 1 << LOG2M, Nx = 2 * M;
    constexpr int ROW_STRIDE = 2 * G_::LPAD + 8;   // >= Nx complex: the buffer later holds the field's whole transformed row
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R;
    static_assert(MODE == EPI_CROSS || MODE == EPI_PHASE || MODE == EPI_CROSS_AND_PHASE, "two-field modes only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* smw = smem + 2 * ROWS * ROW_STRIDE;   // [M] radix-2 twiddles
    const int r = threadIdx.x / (2 * NT), f = (threadIdx.x / NT) & 1, u = threadIdx.x % NT;
    cplx<T>* sm = smem + (2 * r + f) * ROW_STRIDE;
    for (int k = threadIdx.x; k < M; k += 2 * NT * ROWS) smw[k] = __ldg(io.tw2 + k);
    __syncthreads();
    const long ngroups = (nseq + ROWS - 1) / ROWS;
    const int Ny = 1 << io.logNy;
    const int sy = io.shift_y ? Ny / 2 : 0, sx = io.shift_x ? Nx / 2 : 0;
    const cplx<T>* zf = f ? io.z2 : io.z1;
    const cplx<T>* agf = f ? io.ag2 : io.ag1;
    // (item, row) of this thread's sequence, advanced by the grid stride without a division per group
    const long gstep = (long)gridDim.x * ROWS;
    const long bstep = gstep / io.H;
    const int kstep = (int)(gstep - bstep * io.H);
    long seq = (long)blockIdx.x * ROWS + r;
    long b = seq / io.H;
    int ky = (int)(seq - b * io.H);
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x, seq += gstep, b += bstep, ky += kstep) {
        if (ky >= io.H) { ky -= io.H; ++b; }
        const long nxt = grp + gridDim.x;
        if (threadIdx.x < 4 * ROWS && nxt < ngroups) {   // next group's rows (ky and Ny - ky of both fields) -> L2
            const long s2 = nxt * ROWS + (threadIdx.x >> 2);
            if (s2 < nseq) {
                const long b2 = s2 / io.H;
                const int k2 = (int)(s2 - b2 * io.H);
                const int row = (threadIdx.x & 1) ? ((Ny - k2) & (Ny - 1)) : k2;
                prefetch_l2_bulk(((threadIdx.x & 2) ? io.z2 : io.z1) + ((b2 << io.logNy) + row) * (long)M, (unsigned)(M * sizeof(cplx<T>)));
            }
        }
        const bool act = seq < nseq;
        cplx<T> v[2][E];
        {
            const cplx<T>* pa = zf + ((b << io.logNy) + ky) * (long)M + u;
            const cplx<T>* pb = zf + ((b << io.logNy) + ((Ny - ky) & (Ny - 1))) * (long)M + u;
            cplx<T> za[E], zb[E];
#pragma unroll
            for (int q = 0; q < E; ++q) { za[q] = act ? pa[q * NT] : mk<T>(0, 0); zb[q] = act ? pb[q * NT] : mk<T>(0, 0); }
#pragma unroll
            for (int q = 0; q < E; ++q) {
                v[0][q] = mk<T>(za[q].x + zb[q].x, za[q].y - zb[q].y);
                v[1][q] = mk<T>(za[q].y + zb[q].y, zb[q].x - za[q].x);
            }
        }
        if (agf != nullptr && act) {
            const cplx<T> W = __ldg(io.wj + 2 * ky), J = __ldg(io.wj + 2 * ky + 1);
            const cplx<T>* pa = agf + b * (long)Nx + 2 * u;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(pa + 2 * q * NT));
                v[0][q].x += a.x * W.x + a.y * J.x; v[0][q].y += a.x * W.y + a.y * J.y;
                v[1][q].x += a.z * W.x + a.w * J.x; v[1][q].y += a.z * W.y + a.w * J.y;
            }
        }
        block_fft<T, LOG2M, LOGE, 2, 2>(v, u, sm, 1, tw);   // ends with a barrier: the exchange buffers are free
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int k = final_index<LOG2M, LOGE>(u, g, t);
                const cplx<T> fa = v[0][g + t * G];
                const cplx<T> wb = cmul(v[1][g + t * G], smw[k]);
                sm[k] = cadd(fa, wb);
                sm[k + M] = csub(fa, wb);
            }
        __syncthreads();
        if (act) {
            const cplx<T>* s1 = smem + (2 * r) * ROW_STRIDE;
            const cplx<T>* s2 = s1 + ROW_STRIDE;
            const int tt = threadIdx.x % (2 * NT);
            const long rd = ((b << io.logNy) + ((ky + sy) & (Ny - 1))) * (long)Nx;
            const long rm = ((b << io.logNy) + ((Ny - ky + sy) & (Ny - 1))) * (long)Nx;
            const bool self = (ky == 0) || (2 * ky == Ny);   // the row mirrors into itself: kx <= Nx/2 computed, the rest mirrored
            constexpr int J = Nx / (2 * NT);
            if (!self) {
                // kx = tt + j 2NT with tt < 2NT and sx in {0, M}: kx < M exactly for j < J/2, so every destination is a base
                // pointer plus an immediate offset; only kx = 0 and kx = M can wrap in the mirrored row (index 0)
                const long dlo = rd + tt + sx, dhi = rd + tt - sx;                       // (kx + sx) mod Nx for kx < M / kx >= M
                const long mlo = rm + ((sx ? M : Nx) - tt), mhi = rm + ((Nx + sx) - tt);   // (Nx - kx + sx) mod Nx likewise
                const long m0 = rm + ((Nx - tt + sx) & (Nx - 1)), mM = rm + ((Nx - (tt + M) + sx) & (Nx - 1));
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int off = j * (2 * NT);
                    const cplx<T> c = cscale(cmulc(s1[tt + off], s2[tt + off]), io.scale);
                    const long pd = (j < J / 2 ? dlo : dhi) + off;
                    const long pm = j == 0 ? m0 : j == J / 2 ? mM : (j < J / 2 ? mlo : mhi) - off;
                    if constexpr (MODE == EPI_PHASE || MODE == EPI_CROSS_AND_PHASE) {
                        T* o = reinterpret_cast<T*>(MODE == EPI_PHASE ? io.out : io.out2);
                        const T ph = xatan2(c.y, c.x);
                        o[pd] = ph;
                        o[pm] = -ph;
                    }
                    if constexpr (MODE == EPI_CROSS || MODE == EPI_CROSS_AND_PHASE) {
                        cplx<T>* o = reinterpret_cast<cplx<T>*>(io.out);
                        o[pd] = c;
                        o[pm] = cconj(c);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int kx = tt + j * (2 * NT);
                    const cplx<T> c = cscale(cmulc(s1[kx], s2[kx]), io.scale);
                    const int pd = (kx + sx) & (Nx - 1), pm = (Nx - kx + sx) & (Nx - 1);
                    const bool wd = 2 * kx <= Nx, wm = kx > 0 && 2 * kx < Nx;
                    if constexpr (MODE == EPI_PHASE || MODE == EPI_CROSS_AND_PHASE) {
                        T* o = reinterpret_cast<T*>(MODE == EPI_PHASE ? io.out : io.out2);
                        const T ph = xatan2(c.y, c.x);
                        if (wd) o[rd + pd] = ph;
                        if (wm) o[rm + pm] = -ph;
                    }
                    if constexpr (MODE == EPI_CROSS || MODE == EPI_CROSS_AND_PHASE) {
                        cplx<T>* o = reinterpret_cast<cplx<T>*>(io.out);
                        if (wd) o[rd + pd] = c;
                        if (wm) o[rm + pm] = cconj(c);
                    }
                }
            }
        }
        __syncthreads();   // the buffers are scattered into again by the next group
    }
}

// =============================================================================================
// K-B : columns
// =============================================================================================
template <typename T, int LOG2L, int LOGE, int C, int V, class IO>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * (C / V), min_blocks_for((1 << (LOG2L - LOGE)) * (C / V)))
cols_kernel(const __grid_constant__ IO io, const cplx<T>* __restrict__ tw, long ntiles) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, CG = C / V;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    const int cg = threadIdx.x % CG, u = threadIdx.x / CG;
    cplx<T>* sm = smem + cg * V;
    io.template init<LOG2L, LOGE, C, V>(smem);
    // software pipeline: the loads of the NEXT tile are issued between staging the epilogue values in shared memory
    // (after which the registers are dead) and the cooperative store loop, so their latency hides behind the stores
    cplx<T> v[V][E];
    if ((long)blockIdx.x < ntiles) io.template load<LOG2L, LOGE, C, V>((long)blockIdx.x, u, cg, v, 0);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long nxt = tile + gridDim.x;
        if (threadIdx.x == 0 && nxt + gridDim.x < ntiles) io.template prefetch<LOG2L, C>(nxt + gridDim.x);
        io.tma_reads_done();  // staging buffer (aliases the exchange buffer) may be overwritten from here on
        {
            cplx<T> ag_[E];
            io.template fix_fetch<LOG2L, LOGE>(tile, u, ag_);
            io.template fix_apply<LOG2L, LOGE, C, V>(tile, u, cg, ag_, v, reinterpret_cast<float*>(smem + G_::LPAD * C));
        }
        block_fft<T, LOG2L, LOGE, V, C>(v, u, sm, 1, tw);
        io.template store_a<LOG2L, LOGE, C, V>(tile, u, cg, v, smem);
        if (nxt < ntiles) io.template load<LOG2L, LOGE, C, V>(nxt, u, cg, v, 0);
        io.template store_b<LOG2L, LOGE, C, (1 << (LOG2L - LOGE)) * (C / V)>(tile, smem);
    }
    io.tma_drain();
}

// Two-field variant (cross spectrum / phase / complex bins): each thread owns the SAME column of both fields, the two
// fields are interleaved in shared memory ([pad(l)][C][2]) and transformed by one block_fft call (shared twiddles,
// 128-bit exchanges); F1 and F2 of a cell end up in the same thread, so F1 conj(F2) needs no extra exchange.
template <typename T, int LOG2L, int LOGE, int C, class IO>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * C, min_blocks_for((1 << (LOG2L - LOGE)) * C))
cols2f_kernel(const __grid_constant__ IO io, const cplx<T>* __restrict__ tw, long ntiles) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    const int c = threadIdx.x % C, u = threadIdx.x / C;
    cplx<T>* sm = smem + 2 * c;
    io.template init<LOG2L, LOGE, C, 1>(smem);
    cplx<T> v[2][E];
    if ((long)blockIdx.x < ntiles) io.template load2f<LOG2L, LOGE, C>((long)blockIdx.x, u, c, v);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long nxt = tile + gridDim.x;
        if (threadIdx.x == 0 && nxt + gridDim.x < ntiles) io.template prefetch<LOG2L, C>(nxt + gridDim.x);
        block_fft<T, LOG2L, LOGE, 2, 2 * C>(v, u, sm, 1, tw);
        io.template store_a2f<LOG2L, LOGE, C>(tile, u, c, v, smem);
        if (nxt < ntiles) io.template load2f<LOG2L, LOGE, C>(nxt, u, c, v);
        io.template store_b<LOG2L, LOGE, C, (1 << (LOG2L - LOGE)) * C>(tile, smem);
    }
}

// Asynchronous variant of cols_kernel for tiles that are ONE contiguous chunk of global memory (the blocked
// intermediate): the next tile is brought into shared memory by bulk copies (TMA engine) that stay in flight while
// the current tile is transformed and stored, instead of by register loads issued just before the store phase.
// Shared memory: landing buffer L [LPAD][C] (also exchange #0), half-size buffer X [LPAD][C/2] (later exchanges,
// one register sequence at a time, and the epilogue staging), one mbarrier, then the IO's histogram.
template <typename T, int LOG2L, int LOGE, int C, class IO>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * (C / 2), min_blocks_for((1 << (LOG2L - LOGE)) * (C / 2)))
cols_async_kernel(const __grid_constant__ IO io, const cplx<T>* __restrict__ tw, long ntiles) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, CG = C / 2, NT = G_::NT, L = 1 << LOG2L, NTHR = NT * CG;
    static_assert(LOG2L > LOGE, "needs at least one exchange");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx<T>* smL = reinterpret_cast<cplx<T>*>(smem_raw);
    cplx<T>* smX = smL + G_::LPAD * C;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smX + G_::LPAD * CG);
    float* extra = reinterpret_cast<float*>(bar + 2);   // IO scratch (histogram / partial sums) behind the barrier
    const int cg = threadIdx.x % CG, u = threadIdx.x / CG;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init_fence(); }
    io.template init<LOG2L, LOGE, C, 2>(smX);   // ends with a barrier when there is a histogram
    __syncthreads();
    // the landing buffer was last touched through the generic proxy (exchange #0); the barrier before the issue orders those
    // accesses before this thread, whose proxy fence orders them before the bulk copies that re-fill the buffer
    auto issue = [&](long tile) { fence_proxy_async_smem(); io.template issue_load<LOG2L, C>(tile, smL, bar); };
    if (threadIdx.x == 0 && (long)blockIdx.x < ntiles) issue((long)blockIdx.x);
    unsigned phase = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long nxt = tile + gridDim.x;
        cplx<T> ag_[E];
        io.template fix_fetch<LOG2L, LOGE>(tile, u, ag_);   // issued before the wait: their latency hides behind it
        float4 wc4 = make_float4(1.f, 1.f, 1.f, 1.f);
        float* extra_t = extra;   // IO scratch of this tile (ColsR2CPack: alternate tiles use alternate copies, no barrier needed)
        if constexpr (IO::kSplitEpilogue) { wc4 = io.template col_fetch<C>(tile, cg); extra_t += phase ? IO::kExtraHalf : 0; }
        mbar_wait(bar, phase);
        phase ^= 1u;
        cplx<T> v[2][E];
        {
            const cplx<T>* pl = smL + u * C + cg * 2;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                v[0][q] = pl[q * (NT * C)];
                v[1][q] = pl[q * (NT * C) + 1];
            }
        }
        io.template fix_apply<LOG2L, LOGE, C, 2>(tile, u, cg, ag_, v, extra_t, smL, wc4);
        io.tma_reads_done();  // asynchronous stores of the previous tile have finished reading the staging buffer X
        StagesAsync<T, LOG2L, LOGE, 0, C>::run(v, u, smL + cg * 2, smX + cg, tw,
                                               [&]() { if (threadIdx.x == 0 && nxt < ntiles) issue(nxt); });
        if constexpr (IO::kSplitEpilogue) {
            if (io.zout != nullptr) {
                if (io.ztma) io.template store_z_tma<LOG2L, LOGE, C, NTHR>(tile, u, cg, v, smX, extra_t);
                else io.template store_z<LOG2L, LOGE, C, NTHR>(tile, u, cg, v, extra_t);
            }
            else io.template store_split<LOG2L, LOGE, C, NTHR>(tile, u, cg, v, smX, extra_t);
        } else {
            io.template store_a<LOG2L, LOGE, C, 2>(tile, u, cg, v, smX);
            io.template store_b<LOG2L, LOGE, C, NTHR>(tile, smX);
        }
    }
    io.tma_drain();
}

// ---- plain strided C2C on a [A][L][B] row-major view (in-place safe) ---------------------------
template <typename T> struct ColsC2C {
    static constexpr bool kTwoFields = false;
    static constexpr bool kBins = false;
    static constexpr int kExtraSmemBytes = 0;
    const cplx<T>* in; cplx<T>* out; long B; long tiles_per_row; int inverse; T scale;
    // Hooks of the four-step decomposition n = n1 n2 of a transform too long for one CTA (c2c_pass): the twiddle product and
    // the final transposition ride on the stores of the two strided passes instead of being passes of their own.
    //   step A (FFT over i1 on the view [A][n1][n2 B]): output point k1 of column b is multiplied by w_n^(k1 i2), i2 = b / tw4_div
    //   step B (FFT over i2 on the view [A n1][n2][B]): output point k2 of item (a, k1) is stored at row k1 + n1 k2 of item a
    const cplx<T>* tw4 = nullptr;   // exp(-2 pi i m / n), m in [0, min(n, 8192)); nullptr = no twiddle
    long tw4_div = 1;
    int tr_n1 = 0;                  // n1 of the transposed store; 0 = natural order
    const cplx<T>* tw4_hi = nullptr;   // n > 8192: exp(-2 pi i 8192 h / n), h in [0, n / 8192); w^m = tw4_hi[m >> 13] tw4[m & 8191]
    // ---- hooks of xrftb_fft2r on the strided axis of length hook_n (0 = off).  Element (l, column b) of this pass is row
    // r = l * row_mul + b / row_div of the full axis (row_mul = n2, row_div = B in step A of a four-step pass; 1 and 0 = "no
    // division" in a single pass).  Loads: row r reads SOURCE row (r + in_roll) % hook_n of a source array with row pitch
    // (row_div ? row_div : B); source rows outside [in_lo, in_hi) are zero and never read (in_hi = 0: all valid); the element
    // is multiplied by in_ramp[source row].  Stores (last pass): output row R is multiplied by out_ramp[R] and stored at row
    // (R + out_roll) % hook_n; with out_hi > 0 only stored rows in [out_lo, out_hi) are written, into items of out_hi - out_lo rows.
    long hook_n = 0, row_mul = 1, row_div = 0;
    long in_roll = 0, in_lo = 0, in_hi = 0;
    const cplx<T>* in_ramp = nullptr;
    const cplx<T>* out_ramp = nullptr;
    long out_roll = 0, out_lo = 0, out_hi = 0;
    // Bluestein's chirp-z (c2c_pass): the source holds only in_rows (< hook_n) rows per item, is conjugated before the ramp
    // (in_conj), and the result is conjugated after the ramp (out_conj)
    long in_rows = 0;
    int in_conj = 0, out_conj = 0;
    __device__ __forceinline__ bool in_hooks() const { return hook_n > 0 && (in_roll != 0 || in_hi > 0 || in_ramp != nullptr || in_rows > 0 || in_conj); }
    __device__ __forceinline__ bool out_hooks() const { return hook_n > 0 && (out_roll != 0 || out_hi > 0 || out_ramp != nullptr || out_conj); }

    template <int LOG2L, int C> __device__ __forceinline__ void prefetch(long) const {}
    template <int LOG2L, int LOGE, int C, int V> __device__ __forceinline__ void init(cplx<T>*) const {}
    __device__ __forceinline__ void tma_reads_done() const {}
    __device__ __forceinline__ void tma_drain() const {}
    template <int LOG2L, int LOGE> __device__ __forceinline__ void fix_fetch(long, int, cplx<T> (&)[1 << LOGE]) const {}
    // Everything that CONSUMES the loaded values (input ramp, conjugation of the inverse-through-forward trick) happens here,
    // at the start of the tile's transform -- not in load(), which runs one tile ahead: values consumed right after their
    // load would stall the thread on the memory latency that the software pipeline is there to hide.
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void fix_apply(long tile, int u, int cg, const cplx<T> (&)[1 << LOGE], cplx<T> (&v)[V][1 << LOGE], float* = nullptr, const cplx<T>* = nullptr, float4 = make_float4(1.f, 1.f, 1.f, 1.f)) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        if (hook_n > 0 && in_conj) {
#pragma unroll
            for (int vv = 0; vv < V; ++vv)
#pragma unroll
                for (int q = 0; q < (1 << LOGE); ++q) v[vv][q].y = -v[vv][q].y;
        }
        if (hook_n > 0 && in_ramp != nullptr) {
            const long a = tile / tiles_per_row;
            const long b0 = (tile - a * tiles_per_row) * C + cg * V;
            const int n = (int)hook_n, lo = (int)in_lo, hi = in_hi > 0 ? (int)in_hi : n;
            const int rstep = NT * (int)row_mul;
#pragma unroll
            for (int vv = 0; vv < V; ++vv) {
                const unsigned bb = (unsigned)(b0 + vv);
                const int i2 = row_div ? (int)(bb / (unsigned)row_div) : 0;
                const int r0 = u * (int)row_mul + i2 + (int)in_roll;
#pragma unroll
                for (int q = 0; q < (1 << LOGE); ++q) {
                    int rs = r0 + q * rstep;
                    if (rs >= n) rs -= n;
                    if (rs >= lo && rs < hi) v[vv][q] = cmul(v[vv][q], __ldg(in_ramp + rs));   // (rows outside are zero; the ramp ends at hi)
                }
            }
        }
        if (inverse) {
#pragma unroll
            for (int vv = 0; vv < V; ++vv)
#pragma unroll
                for (int q = 0; q < (1 << LOGE); ++q) v[vv][q].y = -v[vv][q].y;
        }
    }

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void load(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], int) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, L = 1 << LOG2L;
        const long a = tile / tiles_per_row;
        const long b0 = (tile - a * tiles_per_row) * C + cg * V;
        if (in_hooks()) {
            // 32-bit index arithmetic inside an item (the launcher guarantees hook_n * pitch < 2^31): the row of element q is
            // r0 + q * rstep, wrapped once by the roll
            const int pitch = (int)(row_div ? row_div : B);
            const int n = (int)hook_n, lo = (int)in_lo, hi = in_hi > 0 ? (int)in_hi : n;
            const cplx<T>* base = in + a * (in_rows > 0 ? in_rows : hook_n) * pitch;
            const int rstep = NT * (int)row_mul;
#pragma unroll
            for (int vv = 0; vv < V; ++vv) {
                const unsigned bb = (unsigned)(b0 + vv);
                const int i2 = row_div ? (int)(bb / (unsigned)row_div) : 0;
                const int bcol = row_div ? (int)(bb - (unsigned)i2 * (unsigned)row_div) : (int)bb;
                const int r0 = u * (int)row_mul + i2 + (int)in_roll;
                const bool colok = (long)bb < B;
#pragma unroll
                for (int q = 0; q < (1 << LOGE); ++q) {
                    int rs = r0 + q * rstep;
                    if (rs >= n) rs -= n;
                    cplx<T> x = mk<T>(0, 0);
                    if (colok && rs >= lo && rs < hi) x = base[rs * pitch + bcol];
                    v[vv][q] = x;
                }
            }
            return;
        }
        const cplx<T>* p = in + a * L * B + b0 + (long)u * B;
        const long qstep = (long)NT * B;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
#pragma unroll
            for (int vv = 0; vv < V; ++vv) {
                cplx<T> x = mk<T>(0, 0);
                if (b0 + vv < B) x = p[q * qstep + vv];
                v[vv][q] = x;
            }
        }
    }
    template <int LOG2L, int LOGE, int C, int NTHR> __device__ __forceinline__ void store_b(long, cplx<T>*) const {}
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store_a(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], cplx<T>*) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, L = 1 << LOG2L;
        const long a = tile / tiles_per_row;
        const long b0 = (tile - a * tiles_per_row) * C + cg * V;
        cplx<T>* p = out + a * L * B + b0;
        long ostride = B;
        if (tr_n1) {
            const long ao = a / tr_n1;
            p = out + (ao * tr_n1 * L + (a - ao * tr_n1)) * B + b0;
            ostride = (long)tr_n1 * B;
        }
        int i2[V];   // (b0 + vv < n2 B < 2^31, o * i2 < n <= 2^26: 32-bit arithmetic)
#pragma unroll
        for (int vv = 0; vv < V; ++vv) i2[vv] = tw4 ? (int)((unsigned)(b0 + vv) / (unsigned)tw4_div) : 0;
        if (out_hooks()) {   // last pass of the strided axis: factor, roll and crop ride on the stores
            const long ao = tr_n1 ? (long)((unsigned)a / (unsigned)tr_n1) : a;
            const int k1 = tr_n1 ? (int)(a - ao * tr_n1) : 0;
            const int rstep = tr_n1 ? tr_n1 : 1;
            const int n = (int)hook_n;
            const int lo = out_hi > 0 ? (int)out_lo : 0, hi = out_hi > 0 ? (int)out_hi : n;
            const int Bi = (int)B;
            cplx<T>* ob = out + ao * (long)(hi - lo) * B + b0;
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const int Rr = k1 + rstep * final_index<LOG2L, LOGE>(u, g, t);
                    int Rs = Rr + (int)out_roll;
                    if (Rs >= n) Rs -= n;
                    if (Rs < lo || Rs >= hi) continue;
                    cplx<T> f = mk<T>(1, 0);
                    if (out_ramp != nullptr) f = __ldg(out_ramp + Rr);
#pragma unroll
                    for (int vv = 0; vv < V; ++vv) {
                        if (b0 + vv < B) {
                            cplx<T> x = v[vv][g + t * G];
                            if (inverse) x.y = -x.y;
                            x = cscale(x, scale);
                            if (out_ramp != nullptr) x = cmul(x, f);
                            if (out_conj) x.y = -x.y;
                            ob[(Rs - lo) * Bi + vv] = x;
                        }
                    }
                }
            return;
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const long o = final_index<LOG2L, LOGE>(u, g, t);
#pragma unroll
                for (int vv = 0; vv < V; ++vv) {
                    if (b0 + vv < B) {
                        cplx<T> x = v[vv][g + t * G];
                        if (inverse) x.y = -x.y;
                        x = cscale(x, scale);
                        if (tw4) {
                            const int m = (int)o * i2[vv];
                            cplx<T> w = tw4_hi ? cmul(__ldg(tw4_hi + (m >> 13)), __ldg(tw4 + (m & 8191))) : __ldg(tw4 + m);
                            if (inverse) w.y = -w.y;
                            x = cmul(x, w);
                        }
                        p[o * ostride + vv] = x;
                    }
                }
            }
    }
};

// =============================================================================================
// "columns first" order of the fused 2-D real transform (full-width real-valued epilogues: power spectrum).
//   pass 1  cols_kernel<ColsR2CPack>  : FFT along y of the REAL input, two adjacent real columns packed into one complex
//           sequence (re = column 2c, im = column 2c+1); detrend + window fused into the load; the two spectra are
//           separated after the transform and rows ky in [0, Ny/2] of the half-spectrum are written row-major,
//           [batch][Ny/2+1][Nx] complex, in 2C*8-byte segments.
//   pass 2  rows_kernel<RowsC2CPower> : complex FFT along x of every half-spectrum row (contiguous), |F|^2 * scale,
//           then the row is written twice, fully coalesced: to output row ky, and reversed to row -ky
//           (out[-ky][-kx] = out[ky][kx] for real input).  No separate Hermitian mirror pass, no 16-byte scatter.
// =============================================================================================
template <typename T> struct ColsR2CPack {
    static constexpr bool kTwoFields = false;
    static constexpr bool kBins = false;
    // column-line partial sums [warps][2C][2] + lines [2C][2] (floats), two copies used by alternate tiles (kExtraHalf floats each)
    static constexpr int kExtraSmemBytes = 9216;
    static constexpr int kExtraHalf = 1152;
    const T* in;            // [batch][Ny][Nx] real
    int Nx;                 // row length (elements)
    int tiles_per_item;     // Nx / (2 C): a power of two = 1 << log_tpi (tiles are split into (item, column block) by shifts)
    int detrend;            // 0 none | 1 constant | 2 linear
    const double* moments;  // global-plane variant: [batch][4] : S, -, Sy, Sx (moments_kernel)
    const T* wy; const T* wx;
    cplx<T>* out;           // [batch][Ny/2+1][Nx]
    // column-line variant (float32; no moments pass): every column subtracts its own exactly representable line
    // ph_j(i) = A0 + B i and records (A0, B, sum_i r, sum_i (i - ic) r) of the residual in colstats[batch][Nx]; the
    // difference to the least-squares plane is added to the half-spectrum rows in pass 2 (RowsC2CPower::prologue).
    float4* colstats;
    // cols_async_kernel: tile = Ny rows x 2C reals, fetched as boxes of box_rows rows through a 2-D tensor map of the
    // input viewed as [batch * Ny][Nx]; a dense box row = 2C reals = C packed points, i.e. the [ky][C] layout the kernel reads
    int box_rows;
    alignas(64) CUtensorMap tmap;
    // cols_async_kernel, "z" mode (zout != nullptr): the packed column spectra Z[batch][Ny][Nx/2] are stored as they leave the
    // registers -- no staging, no separation of the two real columns (RowsZPower does it in its loads); w_x / 2 is then
    // applied here, to the real and imaginary part of the packed input, which commutes with the column transform
    cplx<T>* zout;
    // z mode, rows of >= 32 bytes: Z is written by TMA tensor stores (SASS UTMASTG) from the half-size buffer, half a tile
    // at a time -- the LSU never sees the 4096 scattered 32-byte row segments of a tile.  ztmap describes Z as
    // [batch * Ny][Nx] floats with the box of the input map.
    int ztma;
    int zbox_rows;          // rows per box of ztmap (divides Ny / 4)
    alignas(64) CUtensorMap ztmap;
    int log_tpi;            // log2(tiles_per_item), set by cols_r2c_pack
    static constexpr bool kSplitEpilogue = true;
    template <int LOG2L, int LOGE, int C, int NTHR>
    __device__ __forceinline__ void store_z_tma(long tile, int u, int cg, cplx<T> (&v)[2][1 << LOGE], cplx<T>* smX, const float* extra) const {
        using G_ = Geometry<LOG2L, LOGE>;
        // the tile leaves in NPIECE pieces of PR rows through two alternating halves of the buffer: piece p is staged while
        // the stores of piece p-1 are still reading the other half
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, Ny = 1 << LOG2L, NPIECE = 4, PR = Ny / NPIECE;
        const long b = tile >> log_tpi;
        const int t0 = (int)(tile & (long)(tiles_per_item - 1));
        write_colstats<C, NTHR>(b, t0 * (2 * C), extra);
#pragma unroll
        for (int pc = 0; pc < NPIECE; ++pc) {
            cplx<T>* buf = smX + (pc & 1) * (PR * C);
            if (pc >= 2) {   // the stores of piece pc-2 have finished reading this half
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncthreads();
            }
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const int ky = final_index<LOG2L, LOGE>(u, g, t);
                    if (ky / PR == pc) {
                        const cplx<T> a = v[0][g + t * G], c = v[1][g + t * G];
                        if constexpr (sizeof(T) == 4)
                            *reinterpret_cast<float4*>(buf + (ky - pc * PR) * C + cg * 2) = make_float4(a.x, a.y, c.x, c.y);
                    }
                }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (TMA) reads
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned sbase = (unsigned)__cvta_generic_to_shared(buf);
                const int row0 = (int)(b << LOG2L) + pc * PR;
                for (int r0 = 0; r0 < PR; r0 += zbox_rows)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                 :: "l"(&ztmap), "r"(t0 * (2 * C)), "r"(row0 + r0), "r"(sbase + (unsigned)(r0 * C * sizeof(cplx<T>))) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    // z mode: w_x of the thread's four real columns, RAW -- the load is issued ahead of the wait for the tile and nothing may
    // consume it before (fix_apply halves it); other modes: ones
    template <int C> __device__ __forceinline__ float4 col_fetch(long tile, int cg) const {
        float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
        if constexpr (sizeof(T) == 4) {
            if (zout != nullptr && wx != nullptr) {
                const long b = tile >> log_tpi;
                const int x0 = (int)(tile & (long)(tiles_per_item - 1)) * (2 * C) + cg * 4;
                w = __ldg(reinterpret_cast<const float4*>(wx + x0));
            }
        }
        return w;
    }
    template <int LOG2L, int LOGE, int C, int NTHR>
    __device__ __forceinline__ void store_z(long tile, int u, int cg, cplx<T> (&v)[2][1 << LOGE], const float* extra) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        const long b = tile >> log_tpi;
        const int t0 = (int)(tile & (long)(tiles_per_item - 1));
        write_colstats<C, NTHR>(b, t0 * (2 * C), extra);
        const int M = Nx >> 1;
        cplx<T>* ob = zout + (b << LOG2L) * (long)M + t0 * C + cg * 2;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int ky = final_index<LOG2L, LOGE>(u, g, t);
                const cplx<T> a = v[0][g + t * G], c = v[1][g + t * G];
                if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(ob + (long)ky * M) = make_float4(a.x, a.y, c.x, c.y);
                else { ob[(long)ky * M] = a; ob[(long)ky * M + 1] = c; }
            }
    }
    template <int LOG2L, int C> __device__ __forceinline__ void issue_load(long tile, cplx<T>* smL, uint64_t* bar) const {
        constexpr int Ny = 1 << LOG2L;
        const long b = tile >> log_tpi;
        const int x0 = (int)(tile & (long)(tiles_per_item - 1)) * (2 * C);
        const int row0 = (int)(b << LOG2L);
        mbar_expect_tx(bar, (unsigned)(Ny * C * sizeof(cplx<T>)));
        for (int r0 = 0; r0 < Ny; r0 += box_rows) tensor_load_2d_g2s(smL + r0 * C, &tmap, x0, row0 + r0, bar);
    }

    template <int LOG2L, int C> __device__ __forceinline__ void prefetch(long) const {}
    template <int LOG2L, int LOGE, int C, int V> __device__ __forceinline__ void init(cplx<T>*) const {}
    // z mode with tensor stores: the previous tile's stores may still be reading the half-size buffer; only thread 0 waits,
    // the barriers of exchange #0 (which precede every write of that buffer) publish it to the CTA
    __device__ __forceinline__ void tma_reads_done() const {
        if (ztma && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __device__ __forceinline__ void tma_drain() const {
        if (ztma && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    // window factors of the rows this thread owns, fetched ahead of use
    template <int LOG2L, int LOGE> __device__ __forceinline__ void fix_fetch(long, int u, cplx<T> (&a)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        if (wy != nullptr) {
            const T* p = wy + u;
#pragma unroll
            for (int q = 0; q < (1 << LOGE); ++q) a[q].x = __ldg(p + q * NT);
        } else {
#pragma unroll
            for (int q = 0; q < (1 << LOGE); ++q) a[q].x = (T)1;
        }
    }

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void load(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], int) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        const long b = tile >> log_tpi;
        const int x0 = (int)(tile & (long)(tiles_per_item - 1)) * (2 * C) + cg * (2 * V);
        const T* p = in + ((b << LOG2L) + u) * (long)Nx + x0;
        const long qstep = (long)NT * Nx;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            if constexpr (V == 2 && sizeof(T) == 4) {
                const float4 x = *reinterpret_cast<const float4*>(p + q * qstep);
                v[0][q] = mk<T>(x.x, x.y);
                v[1][q] = mk<T>(x.z, x.w);
            } else {
#pragma unroll
                for (int vv = 0; vv < V; ++vv) v[vv][q] = *reinterpret_cast<const cplx<T>*>(p + q * qstep + 2 * vv);
            }
        }
    }
    // detrend + w_y(i) on the freshly loaded tile: rows u + q NT, real columns x0 + 2 vv (+1).  w_x(j) is applied when the
    // two spectra of a packed column are separated (store_b).  `extra` = the partial-sum area behind the exchange buffer.
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void fix_apply(long tile, int u, int cg, const cplx<T> (&a)[1 << LOGE], cplx<T> (&v)[V][1 << LOGE], float* extra,
                                              const cplx<T>* landed = nullptr, float4 wc4 = make_float4(1.f, 1.f, 1.f, 1.f)) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, Ny = 1 << LOG2L, CG = C / V;
        const long b = tile >> log_tpi;
        const int x0 = (int)(tile & (long)(tiles_per_item - 1)) * (2 * C) + cg * (2 * V);
        if (colstats != nullptr) {
            if constexpr (sizeof(T) == 4 && V == 2) {
                if (detrend) {
                    // first and last row of the thread's four real columns: from the landed tile ([row][C] packed points, dense)
                    // when there is one, else from global memory
                    float4 e0, e1;
                    if (landed != nullptr) {
                        e0 = *reinterpret_cast<const float4*>(landed + cg * 2);
                        e1 = *reinterpret_cast<const float4*>(landed + (Ny - 1) * C + cg * 2);
                    } else {
                        const float* top = in + (b << LOG2L) * (long)Nx + x0;
                        e0 = *reinterpret_cast<const float4*>(top);
                        e1 = *reinterpret_cast<const float4*>(top + (long)(Ny - 1) * Nx);
                    }
                    const float x0v[4] = {e0.x, e0.y, e0.z, e0.w}, x1v[4] = {e1.x, e1.y, e1.z, e1.w};
                    float A0[4], B[4], sx[4] = {0.f, 0.f, 0.f, 0.f}, tq[4] = {0.f, 0.f, 0.f, 0.f};
                    const float hf = zout != nullptr ? .5f : 1.f;   // z mode: w_x / 2 (exact halving)
                    const float wc[4] = {hf * wc4.x, hf * wc4.y, hf * wc4.z, hf * wc4.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // same exact-line construction as the row-line prologue of rows2_kernel, without a branch: outside
                        // (1e-30, 1e30) both the quantum Q and its reciprocal are zero, i.e. A0 = B = 0
                        const float m = fmaxf(fabsf(x0v[k]), fabsf(x1v[k]));
                        const bool ok = m > 1e-30f && m < 1e30f;
                        const int p2b = __float_as_int(m) & 0x7f800000;
                        // 2^21 / p2 for a power of two p2 = 2^(e-127): the exponent field 127 + 21 - (e - 127), no division
                        const float Q = ok ? __int_as_float(p2b) * 4.76837158203125e-07f : 0.f;
                        const float iQ = ok ? __int_as_float((275 << 23) - p2b) : 0.f;
                        A0[k] = rintf(x0v[k] * iQ) * Q;
                        B[k] = rintf((x1v[k] - x0v[k]) * (1.0f / (float)(Ny > 1 ? Ny - 1 : 1)) * iQ) * Q;
                    }
                    const float if0 = (float)u;
#pragma unroll
                    for (int q = 0; q < (1 << LOGE); ++q) {
                        const float fi = __fadd_rn(if0, (float)(q * NT));   // (one FADD; not an integer add and a conversion)
                        const float wrow = a[q].x;
                        float r[4] = {v[0][q].x, v[0][q].y, v[1][q].x, v[1][q].y};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            r[k] -= fmaf(B[k], fi, A0[k]);
                            sx[k] += r[k];
                            tq[k] = fmaf((float)q, r[k], tq[k]);
                            r[k] = r[k] * wrow * wc[k];
                        }
                        v[0][q] = mk<T>(r[0], r[1]);
                        v[1][q] = mk<T>(r[2], r[3]);
                    }
                    // column sums over all rows: lanes that share cg (stride CG inside the warp), then across warps via `extra`
                    const float ic = 0.5f * (float)(Ny - 1);
                    constexpr int NTHR = NT * CG, NW = NTHR >= 32 ? NTHR / 32 : 1, LW = NTHR >= 32 ? 32 : NTHR;
                    constexpr unsigned LMASK = NTHR >= 32 ? 0xffffffffu : ((1u << (NTHR & 31)) - 1u);
                    float* part = extra;                       // [NW][CG][4][2]
                    float* lines = extra + NW * CG * 8;        // [CG][4][2]
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float s2 = sx[k];
                        float t2 = fmaf(if0 - ic, sx[k], (float)NT * tq[k]);
#pragma unroll
                        for (int off = CG; off < LW; off <<= 1) {
                            s2 += __shfl_xor_sync(LMASK, s2, off);
                            t2 += __shfl_xor_sync(LMASK, t2, off);
                        }
                        if ((threadIdx.x & (LW - 1)) < CG) {
                            float* pp = part + (((threadIdx.x / LW) * CG + cg) * 4 + k) * 2;
                            pp[0] = s2; pp[1] = t2;
                        }
                        if (u == 0) { lines[(cg * 4 + k) * 2] = A0[k]; lines[(cg * 4 + k) * 2 + 1] = B[k]; }
                    }
                    return;
                }
            }
        }
        double p0 = 0.0, cx = 0.0, cy = 0.0;
        if (detrend) {
            const double* m = moments + b * 4;
            const double npts = (double)Ny * (double)Nx;
            p0 = m[0] / npts;
            if (detrend == 2) {
                const double vy = (double)Nx * ((double)Ny * ((double)Ny * Ny - 1.0) / 12.0);
                const double vx = (double)Ny * ((double)Nx * ((double)Nx * Nx - 1.0) / 12.0);
                cy = Ny > 1 ? m[2] / vy : 0.0;
                cx = m[3] / vx;
                p0 += cy * ((double)u - 0.5 * (Ny - 1)) + cx * ((double)x0 - 0.5 * (Nx - 1));
            }
        }
        const double pstep = cy * (double)NT;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            const T wrow = a[q].x;
            const double pq = p0 + (double)q * pstep;
#pragma unroll
            for (int vv = 0; vv < V; ++vv) {
                cplx<T> y = v[vv][q];
                if (detrend) {
                    y.x = (T)((double)y.x - (pq + cx * (double)(2 * vv)));
                    y.y = (T)((double)y.y - (pq + cx * (double)(2 * vv + 1)));
                }
                y.x *= wrow;
                y.y *= wrow;
                if constexpr (V == 2 && sizeof(T) == 4) {   // z mode: w_x / 2 of the two real columns (1 otherwise: exact)
                    const float hf = zout != nullptr ? .5f : 1.f;
                    y.x *= hf * (vv == 0 ? wc4.x : wc4.z);
                    y.y *= hf * (vv == 0 ? wc4.y : wc4.w);
                }
                v[vv][q] = y;
            }
        }
    }
    // packed spectra Z -> staging [ky][C] (natural order)
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store_a(long, int u, int cg, cplx<T> (&v)[V][1 << LOGE], cplx<T>* smem) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int ky = final_index<LOG2L, LOGE>(u, g, t);
#pragma unroll
                for (int vv = 0; vv < V; ++vv) smem[ky * C + cg * V + vv] = v[vv][g + t * G];
            }
        __syncthreads();
    }
    // Epilogue of cols_async_kernel: the staging buffer holds only half a tile, so the rows are unpacked in two rounds.
    // With Q = Ny/4: round 1 stages ky in [0,Q) and [3Q,Ny) (slots ky, ky-2Q) and unpacks k in [0,Q) (partner Ny-k);
    // round 2 stages ky in [Q,3Q] (slot ky-Q) and unpacks k in [Q,2Q].  Row segments stay 2C*sizeof(cplx) bytes wide.
    template <int LOG2L, int LOGE, int C, int NTHR>
    __device__ __forceinline__ void store_split(long tile, int u, int cg, cplx<T> (&v)[2][1 << LOGE], cplx<T>* smX, const float* extra) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, Ny = 1 << LOG2L, Q = Ny / 4, H = Ny / 2 + 1;
        const long b = tile >> log_tpi;
        const int x0 = (int)(tile & (long)(tiles_per_item - 1)) * (2 * C);
        write_colstats<C, NTHR>(b, x0, extra);
        cplx<T>* ob = out + (b * H) * (long)Nx + x0;
        constexpr bool FIXED_C = (NTHR % C == 0);
        int c = threadIdx.x & (C - 1);
        T ha = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c) : (T)1);
        T hb = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c + 1) : (T)1);
#pragma unroll
        for (int round = 0; round < 2; ++round) {
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int t = 0; t < R; ++t) {
                    const int ky = final_index<LOG2L, LOGE>(u, g, t);
                    const bool outer = (ky < Q) || (ky >= 3 * Q);
                    const bool mine = round == 0 ? outer : (!outer || ky == 3 * Q);
                    if (mine) {
                        const int slot = round == 0 ? (ky < Q ? ky : ky - 2 * Q) : ky - Q;
                        smX[slot * C + cg * 2] = v[0][g + t * G];
                        smX[slot * C + cg * 2 + 1] = v[1][g + t * G];
                    }
                }
            __syncthreads();
            const int k0 = round == 0 ? 0 : Q;
            const int nk = round == 0 ? Q : Q + 1;
            if constexpr (FIXED_C && Q % (NTHR / C) == 0) {
                // Affine form of the loop below: a thread keeps its packed column c and walks the rows kr0 + it STEP.  In both
                // rounds Z[k] sits in slot kr and its partner Z[Ny-k] in slot 2Q - kr (k = 0 pairs with itself), so every
                // shared-memory access is a base pointer plus an immediate offset and the output pointer advances by a constant.
                constexpr int STEP = NTHR / C, SWEEPS = Q / STEP;
                const int kr0 = threadIdx.x / C;
                const cplx<T>* pa = smX + kr0 * C + c;
                const cplx<T>* pb = smX + (2 * Q - kr0) * C + c;
                cplx<T>* po = ob + (long)(k0 + kr0) * Nx + 2 * c;
                const long ostep = (long)STEP * Nx;
#pragma unroll
                for (int it = 0; it <= SWEEPS; ++it) {
                    if (it < SWEEPS || (round == 1 && kr0 == 0)) {   // the extra sweep is row k = 2Q (the Nyquist row), round 1 only
                        const cplx<T> za = pa[it * (STEP * C)];
                        cplx<T> zb = pb[-(it * (STEP * C))];
                        if (round == 0 && it == 0 && kr0 == 0) zb = za;
                        const cplx<T> A = mk<T>(ha * (za.x + zb.x), ha * (za.y - zb.y));
                        const cplx<T> B = mk<T>(hb * (za.y + zb.y), hb * (zb.x - za.x));
                        cplx<T>* q = po + it * ostep;
                        if constexpr (sizeof(T) == 4) {
                            *reinterpret_cast<float4*>(q) = make_float4(A.x, A.y, B.x, B.y);
                        } else {
                            q[0] = A; q[1] = B;
                        }
                    }
                }
            } else
#pragma unroll 4
            for (int idx = threadIdx.x; idx < nk * C; idx += NTHR) {
                const int kr = idx / C;
                if constexpr (!FIXED_C) {
                    c = idx - kr * C;
                    ha = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c) : (T)1);
                    hb = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c + 1) : (T)1);
                }
                const int ky = k0 + kr;
                const int sa = round == 0 ? ky : ky - Q;
                const int sb = round == 0 ? (ky == 0 ? 0 : Ny - ky - 2 * Q) : Ny - ky - Q;
                const cplx<T> za = smX[sa * C + c];
                const cplx<T> zb = smX[sb * C + c];
                const cplx<T> A = mk<T>(ha * (za.x + zb.x), ha * (za.y - zb.y));
                const cplx<T> B = mk<T>(hb * (za.y + zb.y), hb * (zb.x - za.x));
                cplx<T>* po = ob + (long)ky * Nx + 2 * c;
                if constexpr (sizeof(T) == 4) {
                    *reinterpret_cast<float4*>(po) = make_float4(A.x, A.y, B.x, B.y);
                } else {
                    po[0] = A; po[1] = B;
                }
            }
            __syncthreads();
        }
    }
    // fold the per-warp column sums (visible after any barrier behind fix_apply) and record the column lines
    template <int C, int NTHR>
    __device__ __forceinline__ void write_colstats(long b, int x0, const float* extra) const {
        if (colstats != nullptr && detrend) {
            if constexpr (sizeof(T) == 4) {
                constexpr int CG = C / 2, NW = NTHR >= 32 ? NTHR / 32 : 1;
                const float* part = extra;
                const float* lines = part + NW * CG * 8;
                for (int col = threadIdx.x; col < 2 * C; col += NTHR) {   // real column x0 + col, col = cg * 4 + k
                    float s2 = 0.f, t2 = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) { s2 += part[(w * 2 * C + col) * 2]; t2 += part[(w * 2 * C + col) * 2 + 1]; }
                    colstats[b * Nx + x0 + col] = make_float4(lines[col * 2], lines[col * 2 + 1], s2, t2);
                }
            }
        }
    }
    // separate the two real columns of every packed column: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = -i (Z[k] - conj Z[N-k]) / 2,
    // rows k in [0, Ny/2], each scaled by its column's w_x; one thread writes both spectra of one (k, packed column):
    // 2*sizeof(cplx) contiguous bytes, C threads cover one 2C*sizeof(cplx)-byte row segment
    template <int LOG2L, int LOGE, int C, int NTHR>
    __device__ __forceinline__ void store_b(long tile, cplx<T>* smem) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int Ny = 1 << LOG2L, H = Ny / 2 + 1;
        const long b = tile >> log_tpi;
        const int x0 = (int)(tile & (long)(tiles_per_item - 1)) * (2 * C);
        write_colstats<C, NTHR>(b, x0, reinterpret_cast<const float*>(smem + G_::LPAD * C));
        cplx<T>* ob = out + (b * H) * (long)Nx + x0;
        constexpr int ITEMS = H * C;
        constexpr bool FIXED_C = (NTHR % C == 0);       // then the packed column of a thread is the same in every sweep
        int c = threadIdx.x & (C - 1);
        T ha = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c) : (T)1);
        T hb = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c + 1) : (T)1);
#pragma unroll 4
        for (int idx = threadIdx.x; idx < ITEMS; idx += NTHR) {
            const int ky = idx / C;
            if constexpr (!FIXED_C) {
                c = idx - ky * C;
                ha = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c) : (T)1);
                hb = (T)0.5 * (wx != nullptr ? __ldg(wx + x0 + 2 * c + 1) : (T)1);
            }
            const cplx<T> za = smem[ky * C + c];
            const cplx<T> zb = smem[((Ny - ky) & (Ny - 1)) * C + c];
            const cplx<T> A = mk<T>(ha * (za.x + zb.x), ha * (za.y - zb.y));
            const cplx<T> B = mk<T>(hb * (za.y + zb.y), hb * (zb.x - za.x));
            cplx<T>* po = ob + (long)ky * Nx + 2 * c;
            if constexpr (sizeof(T) == 4) {
                *reinterpret_cast<float4*>(po) = make_float4(A.x, A.y, B.x, B.y);
            } else {
                po[0] = A; po[1] = B;
            }
        }
        __syncthreads();
    }
};

// pass 2: rows of the half-spectrum [batch][H][Nx] complex -> power spectrum rows ky and -ky of out [batch][Ny][Nx] real
template <typename T> struct RowsC2CPower {
    static constexpr int kSeqSkew = 0;
    const cplx<T>* in; T* out; int logNy; int H; int shift_y, shift_x; T scale;
    // column-line detrend completion: row ky of item b gets ag[b][j].x * wj[2 ky] + ag[b][j].y * wj[2 ky + 1] added at column j
    // (ag = w_x(j) (alpha_j, gamma_j), wj = transforms of w_y(i) and w_y(i)(i - ic)); nullptr = none
    const cplx<T>* ag; const cplx<T>* wj;

    template <int LOG2L, int SEQ> __device__ __forceinline__ void prefetch(long seq0, long nseq) const {
        constexpr unsigned row_bytes = (unsigned)((1u << LOG2L) * sizeof(cplx<T>));
        if (seq0 + SEQ <= nseq) prefetch_l2_bulk(in + seq0 * (long)(1 << LOG2L), row_bytes * SEQ);
    }
    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void fetch(long seq, bool active, int u, cplx<T> (&raw)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        const cplx<T>* p = in + seq * (long)(1 << LOG2L) + u;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            raw[q] = mk<T>(0, 0);
            if (active) raw[q] = p[q * NT];
        }
    }
    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void prologue(long seq, bool active, int u, cplx<T> (&raw)[1 << LOGE], cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        if (ag != nullptr && active) {
            const long b = seq / H;
            const int ky = (int)(seq - b * H);
            const cplx<T> W = __ldg(wj + 2 * ky), J = __ldg(wj + 2 * ky + 1);
            const cplx<T>* pa = ag + (b << LOG2L) + u;
#pragma unroll
            for (int q = 0; q < (1 << LOGE); ++q) {
                const cplx<T> a = __ldg(pa + q * NT);
                v[q] = mk<T>(raw[q].x + a.x * W.x + a.y * J.x, raw[q].y + a.x * W.y + a.y * J.y);
            }
        } else {
#pragma unroll
            for (int q = 0; q < (1 << LOGE); ++q) v[q] = raw[q];
        }
    }
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_a(long, long seq, bool active, int u, int, cplx<T> (&v)[1 << LOGE], cplx<T>*, int, long) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, Nx = 1 << LOG2L;
        if (!active) return;
        const int Ny = 1 << logNy;
        const long b = seq / H;
        const int ky = (int)(seq - b * H);
        const int sy = shift_y ? Ny / 2 : 0, sx = shift_x ? Nx / 2 : 0;
        T* rowd = out + ((b << logNy) + ((ky + sy) & (Ny - 1))) * (long)Nx;
        T* rowm = out + ((b << logNy) + ((Ny - ky + sy) & (Ny - 1))) * (long)Nx;
        // rows 0 and Ny/2 are their own mirror image (rowm == rowd): their kx <= Nx/2 half is written and mirrored inside
        // the row, so the result is exactly symmetric like every other row pair
        const bool self = (ky == 0) || (2 * ky == Ny);
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const cplx<T> f = v[g + t * G];
                const T val = (f.x * f.x + f.y * f.y) * scale;
                const int kx = final_index<LOG2L, LOGE>(u, g, t);
                if (!self || 2 * kx <= Nx) rowd[(kx + sx) & (Nx - 1)] = val;
                if (!self || (kx > 0 && 2 * kx < Nx)) rowm[(Nx - kx + sx) & (Nx - 1)] = val;
            }
    }
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store_b(long, long, bool, int, int, cplx<T> (&)[1 << LOGE], cplx<T>*, int, long) const {}
};

// ---- fused column pass of the 2-D real transform -------------------------------------------------
// input  : blocked half-spectrum intermediate(s) [batch][ntile][Ny][C] written by RowsR2CFused
// output : by MODE, at fftshift-ed positions, Hermitian-mirrored to the full Nx width when full != 0.
// The epilogue value of every (ky, c) of the tile is staged in shared memory (the exchange buffer is
// free after the last gather), then rows of the tile are written cooperatively: one thread owns one
// output row segment (C direct cells + C mirrored cells), so index math is per row, not per cell.
struct EpilogueDesc {
    int logNx;             // log2 of the real-axis length
    int full;              // 0: half spectrum, width Nx/2+1 (real_dim semantics); 1: full width, mirror cells written here;
                           // 2: full width, only the direct cells (kx <= Nx/2) written here, mirror_fill_kernel completes the rest
    int shift_y, shift_x;  // fftshift on output (shift_x requires full)
    double scale;
    const void* ramp_y;    // complex<T>[Ny]  (unshifted index), nullable
    const void* ramp_x;    // complex<T>[W]   (unshifted index), nullable
    const void* weight_x;  // T[Nx/2+1] real one-sided weights (half mode), nullable
    void* out;             // complex<T> or T, [batch][Ny][W]
    const int* lut;        // bins: int32 [Ny][W] bin of each OUTPUT cell (negative = skip)
    double* bins;          // [batch][nbins] (power) or [batch][nbins][2] (cross)
    int nbins;
    int use_tma;           // POWER: write the direct cells with TMA tensor stores (tmap describes out as [rows][W] float)
    int tma_box_rows;      // rows per TMA box (<= 256, divides Ny/2 so a box never straddles the fftshift wrap)
    const void* fix_ag;    // row-line detrend completion (see ColsFused::ag / wj); nullptr = none
    const void* fix_wj;
    int lut_symmetric;     // bins: lut[-ky][-kx] == lut[ky][kx] for every cell (true for radial bins): mirror cells reuse the bin
};

template <typename T> __device__ __forceinline__ T conj_of(T v) { return v; }
__device__ __forceinline__ float2 conj_of(float2 v) { v.y = -v.y; return v; }
__device__ __forceinline__ double2 conj_of(double2 v) { v.y = -v.y; return v; }
template <typename T> __device__ __forceinline__ T wscale(T v, T w) { return v * w; }
__device__ __forceinline__ float2 wscale(float2 v, float w) { v.x *= w; v.y *= w; return v; }
__device__ __forceinline__ double2 wscale(double2 v, double w) { v.x *= w; v.y *= w; return v; }

template <typename T, int MODE> struct ColsFused {
    static constexpr bool kTwoFields = (MODE == EPI_CROSS || MODE == EPI_PHASE || MODE == EPI_BINS_CROSS);
    static constexpr bool kBins = (MODE == EPI_BINS_POWER || MODE == EPI_BINS_CROSS);
    static constexpr int kExtraSmemBytes = 0;
    static constexpr bool kCplxStage = (MODE == EPI_COMPLEX || MODE == EPI_CROSS || MODE == EPI_PHASE || MODE == EPI_BINS_CROSS);
    static constexpr bool kCplxOut = (MODE == EPI_COMPLEX || MODE == EPI_CROSS);
    using StageT = typename std::conditional<kCplxStage, cplx<T>, T>::type;
    using OutT = typename std::conditional<kCplxOut, cplx<T>, T>::type;
    // per-tile partial sums: fp32 shared atomics are native; a tile adds a few hundred values per bin (rel. error ~1e-6,
    // inside the fp32 tolerance); the cross-tile / cross-item accumulation in global memory is fp64
    using HistT = T;
    const cplx<T>* in1; const cplx<T>* in2; int ntile; EpilogueDesc d;
    int hist_off;                  // bins modes: byte offset of the CTA histogram from the staging-buffer base (set by the launcher)
    // row-line detrend completion (single-field modes): the row pass subtracted per-row lines; the remaining
    // w_y(i) * (alpha_i + gamma_i (j - jc)) transforms along x to  A_i * What(kx) + G_i * Jhat(kx):
    //   ag[b][i] = (A_i, G_i)   (real pair per row, rowline_fix_kernel) ;  wj[2 kx], wj[2 kx + 1] = What(kx), Jhat(kx)
    const cplx<T>* ag; const cplx<T>* wj;
    alignas(64) CUtensorMap tmap;  // only read when d.use_tma

    template <int LOG2L, int LOGE, int C, int V> static constexpr int hist_offset_bytes() {
        using G_ = Geometry<LOG2L, LOGE>;
        return (G_::LPAD * C * (kTwoFields ? 2 : 1)) * (int)sizeof(cplx<T>);
    }
    // cols_async_kernel: the tile is one contiguous chunk of global memory (single-field modes) -> 1-D bulk copies
    static constexpr bool kSplitEpilogue = false;
    template <int LOG2L, int C> __device__ __forceinline__ void issue_load(long tile, cplx<T>* smL, uint64_t* bar) const {
        constexpr unsigned TILE_BYTES = (unsigned)((1u << LOG2L) * C * sizeof(cplx<T>));
        constexpr unsigned PIECE = TILE_BYTES > 32768u ? 32768u : TILE_BYTES;
        static_assert(TILE_BYTES % PIECE == 0 && PIECE % 16 == 0, "bulk copies are multiples of 16 bytes");
        const char* src = reinterpret_cast<const char*>(in1 + tile * (long)(1 << LOG2L) * C);
        mbar_expect_tx(bar, TILE_BYTES);
#pragma unroll
        for (unsigned off = 0; off < TILE_BYTES; off += PIECE) bulk_load_g2s(reinterpret_cast<char*>(smL) + off, src + off, PIECE, bar);
    }
    template <int LOG2L, int LOGE, int C, int V> __device__ __forceinline__ void init(cplx<T>* smem) const {
        if constexpr (kBins) {
            HistT* hist = reinterpret_cast<HistT*>(reinterpret_cast<char*>(smem) + hist_off);
            for (int i = threadIdx.x; i < d.nbins * (kCplxStage ? 2 : 1); i += blockDim.x) hist[i] = 0;
            __syncthreads();
        }
    }

    __device__ __forceinline__ void tma_reads_done() const {
        if constexpr (MODE == EPI_POWER) {
            if (d.use_tma) {
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
            }
        }
    }
    __device__ __forceinline__ void tma_drain() const {
        if constexpr (MODE == EPI_POWER) {
            if (d.use_tma && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }

    template <int LOG2L, int C> __device__ __forceinline__ void prefetch(long tile) const {
        constexpr unsigned bytes = (unsigned)((1u << LOG2L) * C * sizeof(cplx<T>));
        if constexpr (bytes % 16 == 0) {
            prefetch_l2_bulk(in1 + tile * (long)(1 << LOG2L) * C, bytes);
            if (kTwoFields) prefetch_l2_bulk(in2 + tile * (long)(1 << LOG2L) * C, bytes);
        }
    }

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void load(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], int field) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, L = 1 << LOG2L;
        const cplx<T>* p = (field ? in2 : in1) + tile * (long)L * C + cg * V + u * C;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            if constexpr (V == 2 && sizeof(T) == 4) {
                float4 x = *reinterpret_cast<const float4*>(p + q * (NT * C));
                v[0][q] = mk<T>(x.x, x.y);
                v[V - 1][q] = mk<T>(x.z, x.w);
            } else {
#pragma unroll
                for (int vv = 0; vv < V; ++vv) v[vv][q] = p[q * (NT * C) + vv];
            }
        }
    }

    // add the row-line completion to the freshly loaded tile (rows u + q NT, columns cg V + vv); `a` holds ag of those rows
    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void fix_fetch(long tile, int u, cplx<T> (&a)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        if (ag == nullptr) return;
        const long b = tile / ntile;
        const cplx<T>* pa = ag + (b << LOG2L) + u;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) a[q] = __ldg(pa + q * NT);
    }
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void fix_apply(long tile, int, int cg, const cplx<T> (&a)[1 << LOGE], cplx<T> (&v)[V][1 << LOGE], float* = nullptr, const cplx<T>* = nullptr, float4 = make_float4(1.f, 1.f, 1.f, 1.f)) const {
        if (ag == nullptr) return;
        const long b = tile / ntile;
        const int kx0 = (int)(tile - b * ntile) * C + cg * V;
        cplx<T> W[V], J[V];
#pragma unroll
        for (int vv = 0; vv < V; ++vv) { W[vv] = __ldg(wj + 2 * (kx0 + vv)); J[vv] = __ldg(wj + 2 * (kx0 + vv) + 1); }
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q)
#pragma unroll
            for (int vv = 0; vv < V; ++vv) {
                v[vv][q].x += a[q].x * W[vv].x + a[q].y * J[vv].x;
                v[vv][q].y += a[q].x * W[vv].y + a[q].y * J[vv].y;
            }
    }

    // two-field modes: thread (u, c) loads column c of both fields
    template <int LOG2L, int LOGE, int C>
    __device__ __forceinline__ void load2f(long tile, int u, int c, cplx<T> (&v)[2][1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, L = 1 << LOG2L;
        const long off = tile * (long)L * C + c + u * C;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            v[0][q] = in1[off + q * (NT * C)];
            v[1][q] = in2[off + q * (NT * C)];
        }
    }
    template <int LOG2L, int LOGE, int C>
    __device__ __forceinline__ void store_a2f(long, int u, int c, cplx<T> (&v)[2][1 << LOGE], cplx<T>* smem) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        StageT* stage = reinterpret_cast<StageT*>(smem);
        const T sc = (T)d.scale;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int ky = final_index<LOG2L, LOGE>(u, g, t);
                if constexpr (kCplxStage) stage[ky * C + c] = cscale(cmulc(v[0][g + t * G], v[1][g + t * G]), sc);
            }
        __syncthreads();
    }

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store_a(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], cplx<T>* smem) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        constexpr int nthr = 0;
        const cplx<T>* park = nullptr;
        (void)nthr; (void)park;
        constexpr int E = G_::E;
        StageT* stage = reinterpret_cast<StageT*>(smem);
        const T sc = (T)d.scale;
        // ---- 1. epilogue value of each owned (ky, c) -> staging [ky][C]
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int ky = final_index<LOG2L, LOGE>(u, g, t);
#pragma unroll
                for (int vv = 0; vv < V; ++vv) {
                    cplx<T> f = v[vv][g + t * G];
                    StageT val;
                    if constexpr (MODE == EPI_COMPLEX) val = cscale(f, sc);
                    else if constexpr (MODE == EPI_POWER || MODE == EPI_BINS_POWER) val = (f.x * f.x + f.y * f.y) * sc;
                    else {
                        cplx<T> f1 = park[(vv * E + (g + t * G)) * nthr + threadIdx.x];
                        val = cscale(cmulc(f1, f), sc);
                    }
                    stage[ky * C + cg * V + vv] = val;
                }
            }
        if constexpr (MODE == EPI_POWER) {
            if (d.use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (TMA) reads
        }
        __syncthreads();
    }

    template <int LOG2L, int LOGE, int C, int NTHR>
    __device__ __forceinline__ void store_b(long tile, cplx<T>* smem) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int Ny = 1 << LOG2L;
        StageT* stage = reinterpret_cast<StageT*>(smem);
        if constexpr (MODE == EPI_POWER && sizeof(T) == 4) {
            if (d.use_tma) {
                // ---- 2'. TMA tensor stores: box = [box_rows][C] floats (16 B rows), one elected thread, asynchronous;
                //          the LSU never sees the 16-byte scattered row segments
                if (threadIdx.x == 0) {
                    const int Nx_ = 1 << d.logNx;
                    const long b_ = tile / ntile;
                    const int kx0_ = (int)(tile - b_ * ntile) * C;
                    const int sy_ = d.shift_y ? Ny / 2 : 0, sx_ = d.shift_x ? Nx_ / 2 : 0;
                    const int ox0_ = (kx0_ + sx_) & (Nx_ - 1);
                    const int br = d.tma_box_rows;
                    const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage);
                    for (int ky0 = 0; ky0 < Ny; ky0 += br) {
                        const int oy0 = (ky0 + sy_) & (Ny - 1);
                        const int row = (int)(b_ * Ny) + oy0;
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                     :: "l"(&tmap), "r"(ox0_), "r"(row), "r"(sbase + (unsigned)(ky0 * C * sizeof(float))) : "memory");
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                return;  // no trailing barrier: tma_reads_done() guards the buffer before it is reused
            }
        }
        // ---- 2. cooperative row-segment stores
        const int Nx = 1 << d.logNx, M = Nx >> 1;
        const int W = d.full ? Nx : M + 1;
        const long b = tile / ntile;
        const int kx0 = (int)(tile - b * ntile) * C;
        const int sy = d.shift_y ? Ny / 2 : 0, sx = d.shift_x ? Nx / 2 : 0;
        // per-column invariants
        cplx<T> rx_d[C], rx_m[C];
        bool use_rx = false;
        if constexpr (kCplxStage && !kBins) {
            use_rx = d.ramp_x != nullptr;
            if (use_rx) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int kx = kx0 + c;
                    rx_d[c] = kx <= M ? __ldg(reinterpret_cast<const cplx<T>*>(d.ramp_x) + kx) : mk<T>(1, 0);
                    rx_m[c] = (d.full && kx > 0 && kx < M) ? __ldg(reinterpret_cast<const cplx<T>*>(d.ramp_x) + (Nx - kx)) : mk<T>(1, 0);
                }
            }
        }
        T wgt[C];
#pragma unroll
        for (int c = 0; c < C; ++c) wgt[c] = (T)1;
        const bool use_w = d.weight_x != nullptr;
        if (use_w) {
#pragma unroll
            for (int c = 0; c < C; ++c) if (kx0 + c <= M) wgt[c] = __ldg(reinterpret_cast<const T*>(d.weight_x) + kx0 + c);
        }
        const int ox0 = (kx0 + sx) & (Nx - 1);  // direct cells: ox0 + c (no wrap inside an aligned tile)
        OutT* outb = kBins ? nullptr : reinterpret_cast<OutT*>(d.out) + b * (long)Ny * W;
        const bool whole = (kx0 + C - 1 <= M);
        HistT* hist = reinterpret_cast<HistT*>(reinterpret_cast<char*>(smem) + hist_off);
        constexpr int ROW_ITERS = (Ny + NTHR - 1) / NTHR;
#pragma unroll 4
        for (int it = 0; it < ROW_ITERS; ++it) {
            const int ky = threadIdx.x + it * NTHR;
            if (ky >= Ny) break;
            StageT p[C];
#pragma unroll
            for (int c = 0; c < C; ++c) p[c] = stage[ky * C + c];
            const int kym = (Ny - ky) & (Ny - 1);
            const int oy = (ky + sy) & (Ny - 1);
            const int oym = (kym + sy) & (Ny - 1);
            cplx<T> ry_d = mk<T>(1, 0), ry_m = mk<T>(1, 0);
            bool use_ry = false;
            if constexpr (kCplxStage && !kBins) {
                use_ry = d.ramp_y != nullptr;
                if (use_ry) {
                    ry_d = __ldg(reinterpret_cast<const cplx<T>*>(d.ramp_y) + ky);
                    ry_m = __ldg(reinterpret_cast<const cplx<T>*>(d.ramp_y) + kym);
                }
            }
            if constexpr (!kBins) {
                OutT* rowd = outb + (long)oy * W + ox0;
                OutT q[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    StageT x = p[c];
                    if constexpr (kCplxStage) { if (use_ry) x = cmul(x, ry_d); if (use_rx) x = cmul(x, rx_d[c]); }
                    if (use_w) x = wscale(x, wgt[c]);
                    if constexpr (MODE == EPI_PHASE) q[c] = xatan2(x.y, x.x); else q[c] = x;
                }
                if (d.full && whole && (C * sizeof(OutT)) % 16 == 0) {
                    struct alignas(16) Vec { OutT e[C]; };
                    Vec vv_;
#pragma unroll
                    for (int c = 0; c < C; ++c) vv_.e[c] = q[c];
                    *reinterpret_cast<Vec*>(rowd) = vv_;
                } else {
#pragma unroll
                    for (int c = 0; c < C; ++c) if (kx0 + c <= M) outb[(long)oy * W + ((kx0 + c + sx) & (Nx - 1))] = q[c];
                }
                if (d.full == 1) {
                    OutT* rowm = outb + (long)oym * W;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int kx = kx0 + c;
                        if (kx > 0 && kx < M) {
                            StageT x = conj_of(p[c]);
                            if constexpr (kCplxStage) { if (use_ry) x = cmul(x, ry_m); if (use_rx) x = cmul(x, rx_m[c]); }
                            OutT o;
                            if constexpr (MODE == EPI_PHASE) o = xatan2(x.y, x.x); else o = x;
                            rowm[(Nx - kx + sx) & (Nx - 1)] = o;
                        }
                    }
                }
            } else {
                // radial-bin accumulate into the CTA histogram (shared memory); LUT addressed by OUTPUT cell.
                // Consecutive columns of a row mostly fall in the same bin: run-length accumulate in registers and
                // issue one shared atomic per run; mirrored cells share the bin when the LUT is symmetric.
                const int* lutm = d.lut + (long)oym * W;
                int cur = -1;
                HistT acc_x = 0, acc_y = 0;
                auto flush = [&]() {
                    if (cur >= 0) {
                        if constexpr (kCplxStage) { atomicAdd(hist + 2 * cur, acc_x); atomicAdd(hist + 2 * cur + 1, acc_y); }
                        else atomicAdd(hist + cur, acc_x);
                    }
                };
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int kx = kx0 + c;
                    if (kx > M) continue;
                    StageT x = p[c];
                    if (use_w) x = wscale(x, wgt[c]);
                    const bool has_m = d.full && kx > 0 && kx < M;
                    const int bin = __ldg(d.lut + (long)oy * W + ((kx + sx) & (Nx - 1)));
                    HistT vx, vy = 0;
                    if constexpr (kCplxStage) { vx = (HistT)x.x; vy = (HistT)x.y; } else vx = (HistT)x;
                    if (has_m && d.lut_symmetric) {  // value + conj(value) lands in the same bin
                        vx += vx;
                        vy = 0;
                    }
                    if (bin != cur) { flush(); cur = bin; acc_x = 0; acc_y = 0; }
                    if (bin >= 0) { acc_x += vx; acc_y += vy; }
                    if (has_m && !d.lut_symmetric) {
                        const int binm = __ldg(lutm + ((Nx - kx + sx) & (Nx - 1)));
                        if (binm >= 0) {
                            if constexpr (kCplxStage) { atomicAdd(hist + 2 * binm, (HistT)x.x); atomicAdd(hist + 2 * binm + 1, -(HistT)x.y); }
                            else atomicAdd(hist + binm, (HistT)x);
                        }
                    }
                }
                flush();
            }
        }
        __syncthreads();
        if constexpr (kBins) {
            // flush the CTA histogram of this tile to the item's global bins
            const int nb = d.nbins * (kCplxStage ? 2 : 1);
            double* bb = d.bins + b * (long)nb;
            for (int i = threadIdx.x; i < nb; i += NTHR) {
                HistT h = hist[i];
                if (h != (HistT)0) { atomicAdd(bb + i, (double)h); hist[i] = 0; }
            }
            __syncthreads();
        }
    }
};

// =============================================================================================
// Column pass with the radial-bin epilogue of the isotropic power spectrum (BASELINE config 4; xrft/xrft.py:895-906,
// 948-1010), float32, symmetric LUT, full-width semantics.  Same transform as cols_kernel<ColsFused<EPI_BINS_POWER>>; what
// differs is WHERE the bin indices come from.  The grid is a multiple of the tiles per item, so a CTA meets the SAME column
// tile of every plane it processes: each thread reads the bins of its ROW_ITERS x C output cells from the host-built int32
// LUT once, packs them into bytes (nbins <= 255; 255 = masked / padding column) and keeps them in 8 registers for the
// whole kernel -- the per-plane LUT traffic (as large as the plane itself) disappears.  Per cell the epilogue is then a
// byte extract, a compare with the current run's bin and an add; runs of equal bins (radial bins change slowly along a
// row) end in one native fp32 shared-memory atomic.  Mirrored cells (-ky, -kx) share the bin of (ky, kx): columns
// 0 < kx < Nx/2 count twice, columns 0 and Nx/2 (which hold both signs of ky themselves) once.
// =============================================================================================
template <int LOG2L, int LOGE, int C>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * (C / 2), min_blocks_for((1 << (LOG2L - LOGE)) * (C / 2)))
cols_bins_kernel(const __grid_constant__ ColsFused<float, EPI_BINS_POWER> io, const float2* __restrict__ tw, long ntiles) {
    using T = float;
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, V = 2, CG = C / V, NT = G_::NT, NTHR = NT * CG, Ny = 1 << LOG2L;
    constexpr int ROW_ITERS = Ny / NTHR;               // = 32 / C rows per thread
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R;
    static_assert(Ny % NTHR == 0 && C % 4 == 0, "rows per thread and packed LUT words are whole numbers");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    float* stage = reinterpret_cast<float*>(smem_raw);
    float* hist = reinterpret_cast<float*>(smem_raw + io.hist_off);
    const int cg = threadIdx.x % CG, u = threadIdx.x / CG;
    cplx<T>* sm = smem + cg * V;
    const EpilogueDesc& d = io.d;
    const int nb = d.nbins;
    for (int i = threadIdx.x; i < nb; i += NTHR) hist[i] = 0.f;
    // ---- this CTA's column tile (the same for every plane) and the packed bins of this thread's cells
    const int t0 = (int)(blockIdx.x % (unsigned)io.ntile);
    const int kx0 = t0 * C;
    const int Nx = 1 << d.logNx, M = Nx >> 1;
    const int sy = d.shift_y ? Ny / 2 : 0, sx = d.shift_x ? M : 0;
    unsigned lutp[ROW_ITERS][C / 4];
#pragma unroll
    for (int it = 0; it < ROW_ITERS; ++it) {
        const int ky = threadIdx.x + it * NTHR;
        const int* lrow = d.lut + (long)((ky + sy) & (Ny - 1)) * Nx;
#pragma unroll
        for (int w = 0; w < C / 4; ++w) {
            unsigned packed = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kx = kx0 + 4 * w + j;
                int b = 255;
                if (kx <= M) { b = __ldg(lrow + ((kx + sx) & (Nx - 1))); if (b < 0 || b > 254) b = 255; }
                packed |= (unsigned)b << (8 * j);
            }
            lutp[it][w] = packed;
        }
    }
    // columns 0 and Nx/2 hold both signs of ky themselves; every other column stands for its mirror image too
    const float mult0 = (t0 == 0 || kx0 == M) ? 1.f : 2.f;
    const float sc = (float)d.scale;
    __syncthreads();
    cplx<T> v[V][E];
    if ((long)blockIdx.x < ntiles) io.template load<LOG2L, LOGE, C, V>((long)blockIdx.x, u, cg, v, 0);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long nxt = tile + gridDim.x;
        if (threadIdx.x == 0 && nxt + gridDim.x < ntiles) io.template prefetch<LOG2L, C>(nxt + gridDim.x);
        {
            cplx<T> ag_[E];
            io.template fix_fetch<LOG2L, LOGE>(tile, u, ag_);
            io.template fix_apply<LOG2L, LOGE, C, V>(tile, u, cg, ag_, v);
        }
        block_fft<T, LOG2L, LOGE, V, C>(v, u, sm, 1, tw);
        // ---- |F|^2 * scale of each owned (ky, c) -> staging [ky][C] (the exchange buffer is free after the last gather)
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int ky = final_index<LOG2L, LOGE>(u, g, t);
                const cplx<T> f0 = v[0][g + t * G], f1 = v[1][g + t * G];
                *reinterpret_cast<float2*>(stage + ky * C + cg * V) = make_float2((f0.x * f0.x + f0.y * f0.y) * sc, (f1.x * f1.x + f1.y * f1.y) * sc);
            }
        __syncthreads();
        if (nxt < ntiles) io.template load<LOG2L, LOGE, C, V>(nxt, u, cg, v, 0);   // in flight during the epilogue
        // ---- radial-bin accumulate: one thread per row segment of C cells, run-length over equal bins
#pragma unroll
        for (int it = 0; it < ROW_ITERS; ++it) {
            const int ky = threadIdx.x + it * NTHR;
            float p[C];
#pragma unroll
            for (int w = 0; w < C / 4; ++w) {
                const float4 q = *reinterpret_cast<const float4*>(stage + ky * C + 4 * w);
                p[4 * w] = q.x; p[4 * w + 1] = q.y; p[4 * w + 2] = q.z; p[4 * w + 3] = q.w;
            }
            unsigned cur = 255u;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned b = (lutp[it][c / 4] >> (8 * (c % 4))) & 255u;
                const float val = p[c] * (c == 0 ? mult0 : 2.f);
                if (b != cur) {
                    if (cur != 255u) atomicAdd(hist + cur, acc);
                    cur = b; acc = 0.f;
                }
                acc += val;
            }
            if (cur != 255u) atomicAdd(hist + cur, acc);
        }
        __syncthreads();
        // ---- this tile's partial sums -> the plane's fp64 bins
        {
            double* bb = d.bins + (tile / io.ntile) * (long)nb;
            for (int i = threadIdx.x; i < nb; i += NTHR) {
                const float h = hist[i];
                if (h != 0.f) { atomicAdd(bb + i, (double)h); hist[i] = 0.f; }
            }
        }
        __syncthreads();
    }
}

// =============================================================================================
// Radial bins summed on chip.  Shared pieces of the static cell-to-bin machinery (rows_bins_kernel, rowszx_bins_kernel):
// NTHR threads own 16 cells each.
// Counting sort of the CTA's cells by key (once per launch): key[i] in [0, nseg) or 0xFFFF (masked).  On return pos[i] is the
// cell's slot in the key-ordered staging array, id holds the keys of THIS thread's 16 consecutive slots and cont_in / cont_out
// tell whether the run of equal keys at its first / last slot continues in the neighbouring lane of the warp.
// sixteen 16-bit values in eight registers (the slot and key tables of a thread live across the whole tile loop)
struct Packed16 {
    unsigned w[8];
    __device__ __forceinline__ unsigned get(int i) const { return (w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu; }
};
template <int NTHR>
__device__ __forceinline__ void bins_assign_slots(const unsigned short (&key)[16], int nseg, int* cnt, unsigned short* binid, int* scan_part,
                                                  Packed16& pos, Packed16& id, bool& cont_in, bool& cont_out) {
    constexpr int NCELL = NTHR * 16;
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nseg; i += NTHR) cnt[i] = 0;
    for (int i = threadIdx.x; i < NCELL; i += NTHR) binid[i] = 0xFFFFu;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (key[i] != 0xFFFFu) atomicAdd(cnt + key[i], 1);
    __syncthreads();
    {   // exclusive prefix sum of the counters (nseg <= 4096): one thread per run of consecutive segments
        const int per = (nseg + NTHR - 1) / NTHR;
        const int lo = threadIdx.x * per, hi = lo + per < nseg ? lo + per : nseg;
        int local = 0;
        for (int i = lo; i < hi; ++i) local += cnt[i];
        int incl = local;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
        if (lane == 31) scan_part[threadIdx.x >> 5] = incl;
        __syncthreads();
        int run = incl - local;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) run += scan_part[w];
        for (int i = lo; i < hi; ++i) {   // slots [run, run + count) belong to segment i; the counter becomes the fill cursor
            const int c = cnt[i];
            for (int j = 0; j < c; ++j) binid[run + j] = (unsigned short)i;
            cnt[i] = run;
            run += c;
        }
    }
    __syncthreads();
    // the cells of a warp that share a key take consecutive slots (neighbouring cells mostly share the bin): the per-tile
    // scatter into the staging array is then nearly free of bank conflicts
#pragma unroll
    for (int i = 0; i < 8; ++i) pos.w[i] = 0u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const unsigned k = key[i];
        const unsigned grp = __match_any_sync(0xffffffffu, k);
        const int leader = __ffs(grp) - 1;
        int base = 0;
        if (lane == leader && k != 0xFFFFu) base = atomicAdd(cnt + k, __popc(grp));
        base = __shfl_sync(0xffffffffu, base, leader);
        const unsigned p = k == 0xFFFFu ? 0xFFFFu : (unsigned)(base + __popc(grp & ((1u << lane) - 1u))) & 0xFFFFu;
        pos.w[i >> 1] |= p << (16 * (i & 1));
    }
    __syncthreads();
    {
        const uint4* pid = reinterpret_cast<const uint4*>(binid + threadIdx.x * 16);
        const uint4 a = pid[0], b = pid[1];
        id.w[0] = a.x; id.w[1] = a.y; id.w[2] = a.z; id.w[3] = a.w; id.w[4] = b.x; id.w[5] = b.y; id.w[6] = b.z; id.w[7] = b.w;
    }
    const unsigned prev_last = __shfl_up_sync(0xffffffffu, id.get(15), 1);
    cont_in = lane > 0 && prev_last == id.get(0);
    const unsigned next_first = __shfl_down_sync(0xffffffffu, id.get(0), 1);
    cont_out = lane < 31 && next_first == id.get(15);
}
// Sums of the runs of equal keys in the key-ordered staging array: every thread adds up its own 16 consecutive slots
// (128-bit loads, perfectly balanced); runs that span lanes are completed by a warp-wide segmented scan of the partial sums
// still open at the end of each lane; emit(key, sum) is called once per run and warp (a run that crosses a warp boundary is
// emitted in two parts).
template <class Emit>
__device__ __forceinline__ void bins_segmented_sum(const float* slots, const Packed16& id, bool cont_in, bool cont_out, Emit&& emit) {
    const int lane = threadIdx.x & 31;
    float x[16];
    const float4* pv = reinterpret_cast<const float4*>(slots);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float4 q = pv[i]; x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w; }
    float acc = x[0], head = 0.f;
    bool split = false;     // a run boundary inside these slots
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        if (id.get(i) != id.get(i - 1)) {
            if (!split) { head = acc; split = true; } else emit(id.get(i - 1), acc);
            acc = x[i];
        } else {
            acc += x[i];
        }
    }
    float open = acc;
    bool flag = split || !cont_in;     // the run still open at the end of this lane began in this lane
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float y = __shfl_up_sync(0xffffffffu, open, off);
        const bool fy = __shfl_up_sync(0xffffffffu, (int)flag, off) != 0;
        if (lane >= off && !flag) { open += y; flag = fy; }
    }
    const float before = __shfl_up_sync(0xffffffffu, open, 1);   // sum of the run that reaches this lane's first slot
    if (split) emit(id.get(0), head + (cont_in ? before : 0.f));
    if (!cont_out) emit(id.get(15), open);
}

// =============================================================================================
// Pass 2 of the columns-first order with the radial-bin epilogue of the isotropic power spectrum (BASELINE config 4;
// xrft/xrft.py:895-906, 948-1010): rows ky in [0, Ny/2] of the half spectrum are transformed along x exactly like
// rows_kernel<RowsC2CPower>, but |F|^2 never leaves the SM -- it is summed into the plane's radial bins.
//
// No floating-point shared-memory atomics (they compile to compare-and-swap loops and serialise when neighbouring rows hit
// the same bin).  Instead the mapping from cells to bins is made STATIC per CTA: a launch covers `rows` consecutive
// half-spectrum rows of every plane starting at ky0 (rows % SEQ == 0, or rows == 1 for the Nyquist row), the grid is a
// multiple of the rows / SEQ row groups of a plane, so a CTA meets the same rows of every plane it processes.  Once per
// launch each thread looks up the bins of its E cells in the host-built LUT and the CTA counting-sorts its cells by bin
// (native integer atomics; cells of one warp that share a bin get consecutive slots, so the per-tile scatter is nearly
// conflict-free): every cell gets a fixed slot `pos` in a staging array ordered by bin, and a static table names the bin
// of every slot.  Per tile: each thread drops its E values at their slots, one barrier, then every thread adds up its own
// E CONSECUTIVE slots (perfectly balanced, 128-bit loads) -- runs that end inside its slots are complete after a warp-wide
// segmented scan of the partial sums and go to the plane's fp64 bins with one atomic each.  The mirror image (-ky, -kx) of
// a cell shares its bin (radial bins): rows 0 < ky < Ny/2 count twice, rows 0 and Ny/2 (which hold both signs of kx) once.
// rows == 1: the SEQ row slots of a CTA are the same row (ky0) of SEQ consecutive planes; bins are keyed per (slot, bin).
// =============================================================================================
struct RowsBins {
    RowsC2CPower<float> base;   // in, logNy, H, shifts, scale, column-line completion tables (out unused)
    const int* lut;             // int32 [Ny][Nx], bin of each OUTPUT cell (negative = skip)
    double* bins;               // [plane][nbins], accumulated
    int nbins;
    int ky0, rows;              // half-spectrum rows [ky0, ky0 + rows) of every plane
    long nplanes;
};
template <int LOG2L, int LOGE, int SEQ>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * SEQ, min_blocks_for((1 << (LOG2L - LOGE)) * SEQ))
rows_bins_kernel(RowsBins io, const float2* __restrict__ tw) {
    using T = float;
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, NT = G_::NT, NTHR = NT * SEQ, Nx = 1 << LOG2L, NCELL = SEQ * Nx;
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R;
    static_assert(NCELL == NTHR * E && E == 16, "every thread sums E = 16 consecutive slots");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    float* stage = reinterpret_cast<float*>(smem_raw);                               // aliases the exchange buffer
    unsigned char* tail = smem_raw + (size_t)SEQ * G_::LPAD * sizeof(cplx<T>);
    unsigned short* binid = reinterpret_cast<unsigned short*>(tail);                 // [NCELL] key of every slot (0xFFFF = unused)
    int* cnt = reinterpret_cast<int*>(tail + NCELL * sizeof(unsigned short));        // [nseg] counters -> fill cursors (setup only)
    __shared__ int scan_part[32];
    const int s = threadIdx.x / NT, u = threadIdx.x % NT;
    cplx<T>* sm = smem + s * G_::LPAD;
    const bool per_slot = io.rows == 1;              // Nyquist-row launch: one plane per row slot
    const int P = per_slot ? SEQ : 1;
    const int nseg = P * io.nbins;
    const int groups = per_slot ? 1 : io.rows / SEQ;  // row groups per plane; gridDim.x % groups == 0
    const int g0 = (int)(blockIdx.x % (unsigned)groups);
    const int ky = io.ky0 + (per_slot ? 0 : g0 * SEQ + s);
    const int Ny = 1 << io.base.logNy;
    const int sy = io.base.shift_y ? Ny / 2 : 0, sx = io.base.shift_x ? Nx / 2 : 0;
    // ---- once per launch: bins of this thread's cells, counting sort of the CTA's cells by (slot, bin)
    unsigned short key[E];
    Packed16 pos, id;
    bool cont_in, cont_out;
    {
        const int* lrow = io.lut + (long)((ky + sy) & (Ny - 1)) * Nx;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int kx = final_index<LOG2L, LOGE>(u, g, t);
                const int b = __ldg(lrow + ((kx + sx) & (Nx - 1)));
                key[g + t * G] = (b >= 0 && b < io.nbins) ? (unsigned short)((per_slot ? s * io.nbins : 0) + b) : (unsigned short)0xFFFFu;
            }
    }
    bins_assign_slots<NTHR>(key, nseg, cnt, binid, scan_part, pos, id, cont_in, cont_out);
    const float mult = ((ky == 0 || 2 * ky == Ny) ? 1.f : 2.f) * io.base.scale;
    // ---- tiles: mode A: plane = blockIdx.x / groups + k * (gridDim.x / groups); mode B: SEQ planes per tile
    const long tiles_total = per_slot ? (io.nplanes + SEQ - 1) / SEQ : io.nplanes;
    const long tstep = gridDim.x / groups;
    long tile = blockIdx.x / groups;
    auto seq_of = [&](long tl, bool& act) -> long {
        const long plane = per_slot ? tl * SEQ + s : tl;
        act = tl < tiles_total && plane < io.nplanes;
        return plane * io.base.H + ky;
    };
    // The row index ky of a thread is the same for every tile, so the completion factors W(ky), J(ky) of the column-line detrend
    // are loaded once; the per-plane completion vector ag travels one tile ahead together with the row itself.
    const bool fix = io.base.ag != nullptr;
    cplx<T> Wk = mk<T>(0, 0), Jk = mk<T>(0, 0);
    if (fix) { Wk = __ldg(io.base.wj + 2 * ky); Jk = __ldg(io.base.wj + 2 * ky + 1); }
    cplx<T> raw[E], agr[E];
    auto fetch_tile = [&](long tl) {
        bool act; const long sq = seq_of(tl, act);
        io.base.template fetch<LOG2L, LOGE>(sq, act, u, raw);
        if (fix) {
            const long plane = per_slot ? tl * SEQ + s : tl;
            const cplx<T>* pa = io.base.ag + (plane << LOG2L) + u;
#pragma unroll
            for (int q = 0; q < E; ++q) { agr[q] = mk<T>(0, 0); if (act) agr[q] = __ldg(pa + q * NT); }
        }
    };
    fetch_tile(tile);
    for (; tile < tiles_total; tile += tstep) {
        cplx<T> v[1][E];
        if (fix) {
#pragma unroll
            for (int q = 0; q < E; ++q)
                v[0][q] = mk<T>(raw[q].x + agr[q].x * Wk.x + agr[q].y * Jk.x, raw[q].y + agr[q].x * Wk.y + agr[q].y * Jk.y);
        } else {
#pragma unroll
            for (int q = 0; q < E; ++q) v[0][q] = raw[q];
        }
        block_fft<T, LOG2L, LOGE, 1, 1>(v, u, sm, 0, tw);
        __syncthreads();   // every gather of the last exchange is done: the buffer becomes the staging array
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const cplx<T> f = v[0][i];
            const unsigned p = pos.get(i);
            if (p != 0xFFFFu) stage[p] = (f.x * f.x + f.y * f.y) * mult;
        }
        fetch_tile(tile + tstep);   // the next tile's rows and completion vector travel while this one is binned (v is dead here)
        __syncthreads();
        bins_segmented_sum(stage + threadIdx.x * E, id, cont_in, cont_out, [&](unsigned k, float val) {
            if (k == 0xFFFFu) return;
            const long plane = per_slot ? tile * SEQ + (long)(k / (unsigned)io.nbins) : tile;
            if (plane < io.nplanes) atomicAdd(io.bins + plane * (long)io.nbins + (per_slot ? k % (unsigned)io.nbins : k), (double)val);
        });
        __syncthreads();   // the staging array is scattered into again by the next transform
    }
}

// =============================================================================================
// Two-field pass 2 with the radial-bin epilogue of the isotropic cross spectrum (xrft/xrft.py:1098-1187): the front end of
// rowszx_kernel (separation of the packed column spectra, row transforms of both fields) and the static cell-to-bin machinery
// of rows_bins_kernel.  Each of the 2 NT ROWS threads owns 16 cells of a row (kx = tt + 2 NT j).  Cells of rows 0 < ky < Ny/2
// stand for their mirror images too: C(-k) = conj C(k) lands in the same radial bin, so the bin receives 2 Re C and the
// imaginary parts cancel; rows 0 and Ny/2 hold both signs of kx and contribute C as it is.  bins: [plane][nbins][2].
// =============================================================================================
struct RowsZCrossBins {
    RowsZCross<float> base;     // z1, z2, logNy, H, shifts, scale, completion tables, tw2 (out / out2 unused)
    const int* lut;
    double* bins;
    int nbins;
    int ky0, rows;
    long nplanes;
};
template <int LOG2M, int LOGE, int ROWS>
__global__ void __launch_bounds__((1 << (LOG2M - LOGE)) * 2 * ROWS, min_blocks_for((1 << (LOG2M - LOGE)) * 2 * ROWS))
rowszx_bins_kernel(RowsZCrossBins io, const float2* __restrict__ tw) {
    using T = float;
    using G_ = Geometry<LOG2M, LOGE>;
    constexpr int E = G_::E, NT = G_::NT, M = 1 << LOG2M, Nx = 2 * M, NTHR = 2 * NT * ROWS, NCELL = ROWS * Nx;
    constexpr int ROW_STRIDE = 2 * G_::LPAD + 8;
    constexpr int R = 1 << G_::LOGR_LAST, G = E / R;
    static_assert(NCELL == NTHR * 16 && E == 16, "every thread owns 16 cells");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    float* stage_re = reinterpret_cast<float*>(smem_raw);          // both alias the exchange buffers (2 ROWS ROW_STRIDE complex)
    float* stage_im = stage_re + NCELL;
    cplx<T>* smw = smem + 2 * ROWS * ROW_STRIDE;                    // [M] radix-2 twiddles
    unsigned short* binid = reinterpret_cast<unsigned short*>(smw + M);
    int* cnt = reinterpret_cast<int*>(binid + NCELL);
    __shared__ int scan_part[32];
    const int r = threadIdx.x / (2 * NT), f = (threadIdx.x / NT) & 1, u = threadIdx.x % NT, tt = threadIdx.x % (2 * NT);
    cplx<T>* sm = smem + (2 * r + f) * ROW_STRIDE;
    for (int k = threadIdx.x; k < M; k += NTHR) smw[k] = __ldg(io.base.tw2 + k);
    const bool per_slot = io.rows == 1;
    const int P = per_slot ? ROWS : 1;
    const int nseg = P * io.nbins;
    const int groups = per_slot ? 1 : io.rows / ROWS;
    const int g0 = (int)(blockIdx.x % (unsigned)groups);
    const int ky = io.ky0 + (per_slot ? 0 : g0 * ROWS + r);
    const int Ny = 1 << io.base.logNy;
    const int sy = io.base.shift_y ? Ny / 2 : 0, sx = io.base.shift_x ? Nx / 2 : 0;
    const bool self = (ky == 0) || (2 * ky == Ny);
    // only tiles with a self-mirrored row have imaginary parts to add up (uniform per CTA)
    const bool has_im = per_slot ? (io.ky0 == 0 || 2 * io.ky0 == Ny) : (io.ky0 == 0 && g0 == 0) || (2 * (io.ky0 + g0 * ROWS) <= Ny && 2 * (io.ky0 + g0 * ROWS + ROWS - 1) >= Ny);
    unsigned short key[16];
    Packed16 pos, id;
    bool cont_in, cont_out;
    {
        const int* lrow = io.lut + (long)((ky + sy) & (Ny - 1)) * Nx;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int kx = tt + j * (2 * NT);
            const int b = __ldg(lrow + ((kx + sx) & (Nx - 1)));
            key[j] = (b >= 0 && b < io.nbins) ? (unsigned short)((per_slot ? r * io.nbins : 0) + b) : (unsigned short)0xFFFFu;
        }
    }
    bins_assign_slots<NTHR>(key, nseg, cnt, binid, scan_part, pos, id, cont_in, cont_out);
    const float mult = (self ? 1.f : 2.f) * io.base.scale;
    const cplx<T>* zf = f ? io.base.z2 : io.base.z1;
    const cplx<T>* agf = f ? io.base.ag2 : io.base.ag1;
    const long tiles_total = per_slot ? (io.nplanes + ROWS - 1) / ROWS : io.nplanes;
    const long tstep = gridDim.x / groups;
    for (long tile = blockIdx.x / groups; tile < tiles_total; tile += tstep) {
        const long b = per_slot ? tile * ROWS + r : tile;
        const bool act = b < io.nplanes;
        cplx<T> v[2][E];
        {
            const long bb = act ? b : 0;
            const cplx<T>* pa = zf + ((bb << io.base.logNy) + ky) * (long)M + u;
            const cplx<T>* pb = zf + ((bb << io.base.logNy) + ((Ny - ky) & (Ny - 1))) * (long)M + u;
            cplx<T> za[E], zb[E];
#pragma unroll
            for (int q = 0; q < E; ++q) { za[q] = act ? pa[q * NT] : mk<T>(0, 0); zb[q] = act ? pb[q * NT] : mk<T>(0, 0); }
#pragma unroll
            for (int q = 0; q < E; ++q) {
                v[0][q] = mk<T>(za[q].x + zb[q].x, za[q].y - zb[q].y);
                v[1][q] = mk<T>(za[q].y + zb[q].y, zb[q].x - za[q].x);
            }
        }
        if (agf != nullptr && act) {
            const cplx<T> W = __ldg(io.base.wj + 2 * ky), J = __ldg(io.base.wj + 2 * ky + 1);
            const cplx<T>* pa = agf + b * (long)Nx + 2 * u;
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(pa + 2 * q * NT));
                v[0][q].x += a.x * W.x + a.y * J.x; v[0][q].y += a.x * W.y + a.y * J.y;
                v[1][q].x += a.z * W.x + a.w * J.x; v[1][q].y += a.z * W.y + a.w * J.y;
            }
        }
        block_fft<T, LOG2M, LOGE, 2, 2>(v, u, sm, 1, tw);   // ends with a barrier: the exchange buffers are free
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int k = final_index<LOG2M, LOGE>(u, g, t);
                const cplx<T> fa = v[0][g + t * G];
                const cplx<T> wb = cmul(v[1][g + t * G], smw[k]);
                sm[k] = cadd(fa, wb);
                sm[k + M] = csub(fa, wb);
            }
        __syncthreads();
        float cre[16], cim[16];
        {
            const cplx<T>* s1 = smem + (2 * r) * ROW_STRIDE;
            const cplx<T>* s2 = s1 + ROW_STRIDE;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kx = tt + j * (2 * NT);
                const cplx<T> c = cmulc(s1[kx], s2[kx]);
                cre[j] = act ? c.x * mult : 0.f;
                cim[j] = (act && self) ? c.y * mult : 0.f;
            }
        }
        __syncthreads();   // both fields' rows have been read: the buffers become the staging arrays
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const unsigned p = pos.get(j);
            if (p != 0xFFFFu) { stage_re[p] = cre[j]; if (has_im) stage_im[p] = cim[j]; }
        }
        __syncthreads();
        auto target = [&](unsigned k) -> double* {
            const long plane = per_slot ? tile * ROWS + (long)(k / (unsigned)io.nbins) : tile;
            return plane < io.nplanes ? io.bins + (plane * (long)io.nbins + (per_slot ? k % (unsigned)io.nbins : k)) * 2 : nullptr;
        };
        bins_segmented_sum(stage_re + threadIdx.x * 16, id, cont_in, cont_out, [&](unsigned k, float val) {
            if (k == 0xFFFFu) return;
            if (double* p = target(k)) atomicAdd(p, (double)val);
        });
        if (has_im)
            bins_segmented_sum(stage_im + threadIdx.x * 16, id, cont_in, cont_out, [&](unsigned k, float val) {
                if (k == 0xFFFFu) return;
                if (double* p = target(k)) atomicAdd(p + 1, (double)val);
            });
        __syncthreads();   // the buffers are scattered into again by the next tile
    }
}

}  // namespace xrftb
