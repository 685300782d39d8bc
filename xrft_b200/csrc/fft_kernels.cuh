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
#include "fft_core.cuh"

namespace xrftb {

// register budget: keep >= 512 threads resident per SM (<= 128 registers/thread)
constexpr int min_blocks_for(int threads) { return threads >= 512 ? 1 : (512 / threads > 8 ? 8 : 512 / threads); }

enum : int { EPI_COMPLEX = 0, EPI_POWER = 1, EPI_CROSS = 2, EPI_PHASE = 3, EPI_BINS_POWER = 4, EPI_BINS_CROSS = 5 };

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
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long seq = grp * SEQ + s;
        const bool active = seq < nseq;
        cplx<T> v[1][E];
        io.template load<LOG2L, LOGE>(seq, active, u, v[0]);
        block_fft<T, LOG2L, LOGE, 1>(v, u, sm, 1, 0, tw);
        io.template store<LOG2L, LOGE, SEQ>(grp, seq, active, u, s, v[0], smem, SEQ_STRIDE, nseq);
    }
}

// ---- plain C2C (forward, or inverse through conj(FFT(conj x)) * scale) -----------------------
template <typename T> struct RowsC2C {
    static constexpr int kSeqSkew = 0;
    const cplx<T>* in; cplx<T>* out; long in_stride, out_stride; int inverse; T scale;

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void load(long seq, bool active, int u, cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        const cplx<T>* p = in + seq * in_stride;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            cplx<T> x = mk<T>(0, 0);
            if (active) x = p[u + q * NT];
            if (inverse) x.y = -x.y;
            v[q] = x;
        }
    }
    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store(long, long seq, bool active, int u, int, cplx<T> (&v)[1 << LOGE], cplx<T>*, int, long) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        if (!active) return;
        cplx<T>* p = out + seq * out_stride;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                cplx<T> x = v[g + t * G];
                if (inverse) x.y = -x.y;
                p[final_index<LOG2L, LOGE>(u, g, t)] = cscale(x, scale);
            }
    }
};

// R2C split of the packed half-length transform: Z = FFT_M(x[2n] + i x[2n+1]), N = 2M:
//   X[k] = E + w_N^k O,  E = (Z[k] + conj Z[M-k]) / 2,  O = -i (Z[k] - conj Z[M-k]) / 2,  k in [0, M]
template <typename T>
__device__ __forceinline__ cplx<T> r2c_split(cplx<T> zk, cplx<T> zm, cplx<T> w) {
    cplx<T> e = mk<T>((T)0.5 * (zk.x + zm.x), (T)0.5 * (zk.y - zm.y));
    cplx<T> d = mk<T>((T)0.5 * (zk.x - zm.x), (T)0.5 * (zk.y + zm.y));  // (Z[k] - conj Z[M-k]) / 2
    cplx<T> o = mul_mi(d);
    return cadd(e, cmul(w, o));
}

// ---- fused real-input row pass -----------------------------------------------------------------
// prologue : (x - plane) * wy[iy] * wx[ix]       (detrend in fp64 -- SURVEY F6 -- window in T)
// epilogue : R2C split, then either the natural half-spectrum [seq][M+1]   (tileC == 0)
//            or the BLOCKED intermediate [batch][tile][Ny][C] consumed by cols_kernel (tileC > 0):
//            each CTA writes SEQ*C*sizeof(cplx) contiguous bytes per tile, each column tile is
//            one contiguous Ny*C*sizeof(cplx) chunk for the column pass.
template <typename T> struct RowsR2CFused {
    static constexpr int kSeqSkew = 4;  // rows skewed by 4 points: conflict-free cross-row gathers
    const T* in; long in_row_stride;    // real input, element stride between consecutive rows
    int Ny;                              // rows per batch item (1 for 1-D)
    int detrend;                         // 0 none | 1 constant | 2 linear
    const double* moments;               // [batch][4] : S, S0(unused), Sy, Sx -- sum and centred first moments
    const T* wy; const T* wx;            // window vectors (nullable)
    cplx<T>* out; int tileC; long out_seq_stride;  // natural: stride per seq ; blocked: unused
    const cplx<T>* tw_r2c;               // exp(-2 pi i k / N), k in [0, M]

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void load(long seq, bool active, int u, cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT;
        constexpr int Nx = 2 << LOG2L;
        const long b = seq / Ny;
        const int iy = (int)(seq - b * Ny);
        double rowc = 0.0, cx = 0.0;
        if (detrend && active) {
            const double* m = moments + b * 4;
            const double npts = (double)Ny * (double)Nx;
            rowc = m[0] / npts;
            if (detrend == 2) {
                // least-squares plane on a full regular grid == centred first moments (orthogonal regressors)
                const double vy = (double)Nx * ((double)Ny * ((double)Ny * Ny - 1.0) / 12.0);
                const double vx = (double)Ny * ((double)Nx * ((double)Nx * Nx - 1.0) / 12.0);
                const double cy = Ny > 1 ? m[2] / vy : 0.0;
                cx = m[3] / vx;
                rowc += cy * ((double)iy - 0.5 * (Ny - 1)) - cx * (0.5 * (Nx - 1));
            }
        }
        const T wrow = (wy != nullptr && active) ? wy[iy] : (T)1;
        const T* p = in + seq * in_row_stride;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            const int n = u + q * NT;
            cplx<T> x = mk<T>(0, 0);
            if (active) x = *reinterpret_cast<const cplx<T>*>(p + 2 * n);
            if (detrend) {
                x.x = (T)((double)x.x - (rowc + cx * (double)(2 * n)));
                x.y = (T)((double)x.y - (rowc + cx * (double)(2 * n + 1)));
            }
            if (wx != nullptr) {
                cplx<T> w = *reinterpret_cast<const cplx<T>*>(wx + 2 * n);
                x.x *= w.x * wrow; x.y *= w.y * wrow;
            } else {
                x.x *= wrow; x.y *= wrow;
            }
            v[q] = x;
        }
    }

    template <int LOG2L, int LOGE, int SEQ>
    __device__ __forceinline__ void store(long grp, long seq, bool active, int u, int s, cplx<T> (&v)[1 << LOGE],
                                          cplx<T>* smem, int seq_stride, long nseq) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, M = G_::L, NT = G_::NT;
        cplx<T>* sm = smem + s * seq_stride;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) sm[padded<G_::LOGPAD>(final_index<LOG2L, LOGE>(u, g, t))] = v[g + t * G];
        __syncthreads();
        if (tileC == 0) {
            if (active) {
                cplx<T>* p = out + seq * out_seq_stride;
                for (int k = u; k <= M; k += NT) {
                    cplx<T> zk = sm[padded<G_::LOGPAD>(k & (M - 1))];
                    cplx<T> zm = sm[padded<G_::LOGPAD>((M - k) & (M - 1))];
                    p[k] = r2c_split<T>(zk, zm, __ldg(tw_r2c + k));
                }
            }
        } else {
            const int C = tileC;
            const int ntile = M / C + 1;
            const int total = ntile * SEQ * C;
            for (int w = threadIdx.x; w < total; w += NT * SEQ) {
                const int c = w % C;
                const int s2 = (w / C) % SEQ;
                const int t = w / (C * SEQ);
                const int k = t * C + c;
                const long seq2 = grp * SEQ + s2;
                if (seq2 >= nseq) continue;
                const long b = seq2 / Ny;
                const int iy = (int)(seq2 - b * Ny);
                cplx<T> r = mk<T>(0, 0);
                if (k <= M) {
                    const cplx<T>* smr = smem + s2 * seq_stride;
                    cplx<T> zk = smr[padded<G_::LOGPAD>(k & (M - 1))];
                    cplx<T> zm = smr[padded<G_::LOGPAD>((M - k) & (M - 1))];
                    r = r2c_split<T>(zk, zm, __ldg(tw_r2c + k));
                }
                out[((b * ntile + t) * (long)Ny + iy) * C + c] = r;
            }
        }
        __syncthreads();
    }
};

// ---- C2R: half spectrum [seq][M+1] -> real [seq][N], numpy irfft semantics (scale = 1/N folded in) --
//   Z[k] = E + i O,  E = (X[k] + conj X[M-k]) / 2,  O = w_N^{-k} (X[k] - conj X[M-k]) / 2
//   z = IFFT_M(Z) = conj(FFT(conj Z)) ;  x[2n] = Re z[n], x[2n+1] = Im z[n]
template <typename T> struct RowsC2R {
    static constexpr int kSeqSkew = 0;
    const cplx<T>* in; long in_stride; T* out; long out_stride; T scale; const cplx<T>* tw_r2c;

    template <int LOG2L, int LOGE>
    __device__ __forceinline__ void load(long seq, bool active, int u, cplx<T> (&v)[1 << LOGE]) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, M = 1 << LOG2L;
        const cplx<T>* p = in + seq * in_stride;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            const int k = u + q * NT;
            cplx<T> z = mk<T>(0, 0);
            if (active) {
                cplx<T> xk = p[k], xm = p[M - k];
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
    __device__ __forceinline__ void store(long, long seq, bool active, int u, int, cplx<T> (&v)[1 << LOGE], cplx<T>*, int, long) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R;
        if (!active) return;
        T* p = out + seq * out_stride;
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

// =============================================================================================
// K-B : columns
// =============================================================================================
template <typename T, int LOG2L, int LOGE, int C, int V, class IO>
__global__ void __launch_bounds__((1 << (LOG2L - LOGE)) * (C / V), min_blocks_for((1 << (LOG2L - LOGE)) * (C / V)))
cols_kernel(IO io, const cplx<T>* __restrict__ tw, long ntiles) {
    using G_ = Geometry<LOG2L, LOGE>;
    constexpr int E = G_::E, CG = C / V;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx<T>* smem = reinterpret_cast<cplx<T>*>(smem_raw);
    const int cg = threadIdx.x % CG, u = threadIdx.x / CG;
    cplx<T>* sm = smem + cg * V;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        cplx<T> v[V][E];
        io.template load<LOG2L, LOGE, C, V>(tile, u, cg, v, 0);
        block_fft<T, LOG2L, LOGE, V>(v, u, sm, C, 1, tw);
        if constexpr (IO::kTwoFields) {
            // park field-1 spectrum in thread-private smem slots, transform field 2, then combine
            cplx<T>* park = smem + G_::LPAD * C;
            constexpr int NTHR = G_::NT * CG;
#pragma unroll
            for (int vv = 0; vv < V; ++vv)
#pragma unroll
                for (int q = 0; q < E; ++q) park[(vv * E + q) * NTHR + threadIdx.x] = v[vv][q];
            io.template load<LOG2L, LOGE, C, V>(tile, u, cg, v, 1);
            block_fft<T, LOG2L, LOGE, V>(v, u, sm, C, 1, tw);
            io.template store2<LOG2L, LOGE, C, V>(tile, u, cg, v, park, NTHR);
        } else {
            io.template store<LOG2L, LOGE, C, V>(tile, u, cg, v);
        }
    }
}

// ---- plain strided C2C on a [A][L][B] row-major view (in-place safe) ---------------------------
template <typename T> struct ColsC2C {
    static constexpr bool kTwoFields = false;
    const cplx<T>* in; cplx<T>* out; long B; long tiles_per_row; int inverse; T scale;

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void load(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], int) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, L = 1 << LOG2L;
        const long a = tile / tiles_per_row;
        const long b0 = (tile - a * tiles_per_row) * C + cg * V;
        const cplx<T>* p = in + a * L * B + b0;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            const long l = u + q * NT;
#pragma unroll
            for (int vv = 0; vv < V; ++vv) {
                cplx<T> x = mk<T>(0, 0);
                if (b0 + vv < B) x = p[l * B + vv];
                if (inverse) x.y = -x.y;
                v[vv][q] = x;
            }
        }
    }
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE]) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, L = 1 << LOG2L;
        const long a = tile / tiles_per_row;
        const long b0 = (tile - a * tiles_per_row) * C + cg * V;
        cplx<T>* p = out + a * L * B + b0;
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
                        p[o * B + vv] = cscale(x, scale);
                    }
                }
            }
    }
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store2(long, int, int, cplx<T> (&)[V][1 << LOGE], cplx<T>*, int) const {}
};

// ---- fused column pass of the 2-D real transform -------------------------------------------------
// input  : blocked half-spectrum intermediate(s) [batch][ntile][Ny][C] written by RowsR2CFused
// output : by `mode`, at fftshift-ed positions, Hermitian-mirrored to the full Nx width when full != 0
struct EpilogueDesc {
    int mode;              // EPI_*
    int Ny, Nx;            // transform sizes (Nx real axis)
    int full;              // 1: write full Nx-wide spectrum (mirror), 0: half spectrum Nx/2+1 (real_dim semantics)
    int shift_y, shift_x;  // fftshift on output (shift_x requires full)
    double scale;
    const void* ramp_y;    // complex<T>[Ny]  (unshifted index), nullable
    const void* ramp_x;    // complex<T>[Nx or Nx/2+1] (unshifted index), nullable
    const void* weight_x;  // T[Nx/2+1] real one-sided weights (half mode), nullable
    void* out;             // complex<T> or T, [batch][Ny][Wout]
    const int* lut;        // bins: int32 [Ny][Wout] bin of each OUTPUT cell (negative = skip)
    double* bins;          // [batch][nbins] (power) or [batch][nbins][2] (cross)
    int nbins;
};

__device__ __forceinline__ float xatan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double xatan2(double y, double x) { return atan2(y, x); }

template <typename T>
__device__ __forceinline__ void emit_cell(const EpilogueDesc& d, long b, int ky, int kx, bool mirrored, cplx<T> f, cplx<T> f2) {
    // (ky, kx): unshifted output frequency indices in the FULL (or half) grid; f already conj'd if mirrored
    const int W = d.full ? d.Nx : d.Nx / 2 + 1;
    cplx<T> val;
    if (d.mode == EPI_COMPLEX) val = f;
    else if (d.mode == EPI_POWER || d.mode == EPI_BINS_POWER) val = mk<T>(f.x * f.x + f.y * f.y, 0);
    else val = cmulc(f, f2);
    if (d.ramp_y) val = cmul(val, __ldg(reinterpret_cast<const cplx<T>*>(d.ramp_y) + ky));
    if (d.ramp_x) val = cmul(val, __ldg(reinterpret_cast<const cplx<T>*>(d.ramp_x) + kx));
    T sc = (T)d.scale;
    if (d.weight_x) sc *= __ldg(reinterpret_cast<const T*>(d.weight_x) + kx);
    val = cscale(val, sc);
    const int oy = d.shift_y ? ((ky + d.Ny / 2) & (d.Ny - 1)) : ky;
    const int ox = d.shift_x ? ((kx + d.Nx / 2) & (d.Nx - 1)) : kx;
    const long cell = (long)oy * W + ox;
    switch (d.mode) {
        case EPI_COMPLEX:
        case EPI_CROSS:
            reinterpret_cast<cplx<T>*>(d.out)[(b * d.Ny) * (long)W + cell] = val;
            break;
        case EPI_POWER:
            reinterpret_cast<T*>(d.out)[(b * d.Ny) * (long)W + cell] = val.x;
            break;
        case EPI_PHASE:
            reinterpret_cast<T*>(d.out)[(b * d.Ny) * (long)W + cell] = xatan2(val.y, val.x);
            break;
        case EPI_BINS_POWER: {
            const int bin = d.lut[cell];
            if (bin >= 0) atomicAdd(d.bins + b * d.nbins + bin, (double)val.x);
        } break;
        case EPI_BINS_CROSS: {
            const int bin = d.lut[cell];
            if (bin >= 0) {
                atomicAdd(d.bins + (b * d.nbins + bin) * 2, (double)val.x);
                atomicAdd(d.bins + (b * d.nbins + bin) * 2 + 1, (double)val.y);
            }
        } break;
    }
}

template <typename T, bool TWO> struct ColsFused {
    static constexpr bool kTwoFields = TWO;
    const cplx<T>* in1; const cplx<T>* in2; int ntile; EpilogueDesc d;

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void load(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], int field) const {
        constexpr int NT = Geometry<LOG2L, LOGE>::NT, L = 1 << LOG2L;
        const cplx<T>* p = (field ? in2 : in1) + tile * (long)L * C + cg * V;
#pragma unroll
        for (int q = 0; q < (1 << LOGE); ++q) {
            const int l = u + q * NT;
            if constexpr (V == 2 && sizeof(T) == 4) {
                float4 x = *reinterpret_cast<const float4*>(p + (long)l * C);
                v[0][q] = mk<T>(x.x, x.y);
                v[V - 1][q] = mk<T>(x.z, x.w);
            } else {
#pragma unroll
                for (int vv = 0; vv < V; ++vv) v[vv][q] = p[(long)l * C + vv];
            }
        }
    }

    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void emit_all(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], const cplx<T>* park, int nthr) const {
        using G_ = Geometry<LOG2L, LOGE>;
        constexpr int R = 1 << G_::LOGR_LAST, G = G_::E / R, E = G_::E;
        const long b = tile / ntile;
        const int t0 = (int)(tile - b * ntile);
        const int M = d.Nx / 2;
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int ky = final_index<LOG2L, LOGE>(u, g, t);
#pragma unroll
                for (int vv = 0; vv < V; ++vv) {
                    const int kx = t0 * C + cg * V + vv;
                    if (kx > M) continue;
                    cplx<T> f = v[vv][g + t * G];
                    cplx<T> f1 = f, f2 = f;
                    if (TWO) { f1 = park[(vv * E + (g + t * G)) * nthr + threadIdx.x]; f2 = f; }
                    emit_cell<T>(d, b, ky, kx, false, f1, f2);
                    if (d.full && kx > 0 && kx < M) {
                        const int kym = (d.Ny - ky) & (d.Ny - 1);
                        emit_cell<T>(d, b, kym, d.Nx - kx, true, cconj(f1), cconj(f2));
                    }
                }
            }
    }
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE]) const {
        emit_all<LOG2L, LOGE, C, V>(tile, u, cg, v, nullptr, 0);
    }
    template <int LOG2L, int LOGE, int C, int V>
    __device__ __forceinline__ void store2(long tile, int u, int cg, cplx<T> (&v)[V][1 << LOGE], cplx<T>* park, int nthr) const {
        emit_all<LOG2L, LOGE, C, V>(tile, u, cg, v, park, nthr);
    }
};

}  // namespace xrftb
