// xrft_b200 -- contiguous-axis pass dispatch (length -> template instantiation).
#pragma once
#include <cstdlib>
#include "launch.cuh"

namespace xrftb {

#define XRFTB_ROWS_CASES(X) \
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)

template <typename T>
int rows_c2c(const cplx<T>* in, cplx<T>* out, int log2L, long nseq, long in_stride, long out_stride, int inverse, T scale,
             cudaStream_t st, const RowsC2C<T>* extra) {
    RowsC2C<T> io{};
    if (extra) io = *extra;
    io.in = in; io.out = out; io.in_stride = in_stride; io.out_stride = out_stride; io.inverse = inverse; io.scale = scale;
    switch (log2L) {
#define X(K) case K: return launch_rows<T, K, rows_seq_generic<K, cmin(TypeCfg<T>::LOGE, K)>()>(io, nseq, st);
        XRFTB_ROWS_CASES(X)
#undef X
        case 14:
            if constexpr (TypeCfg<T>::MAX_ROWS_LOG2 >= 14) return launch_rows<T, 14, 1>(io, nseq, st);
        default: break;
    }
    set_error("rows_c2c: unsupported length 2^%d", log2L);
    return -2;
}

template <typename T>
int rows_r2c(RowsR2CFused<T> io, int log2M, long nseq, cudaStream_t st) {
    io.tw_r2c = twiddle_r2c<T>(log2M + 1);
    if (!io.tw_r2c) return -3;
    // hot 2-D path: two rows per thread (float32, blocked output, enough rows per item)
    if constexpr (sizeof(T) == 4) {
        // PAIRS row pairs per CTA so that a CTA has 256 threads; its 2*PAIRS rows must be consecutive rows of one item
        if (rows2_eligible(log2M, io.logNy) && io.logC >= 0 && io.in_row_stride == (2L << log2M)) {
            switch (log2M) {
#define Z(K, P) case K: if (io.logNy >= ilog2c(2 * P) && nseq % (2 * P) == 0) \
                    return io.rowstats ? launch_rows2<T, K, P, true>(io, nseq, st) : launch_rows2<T, K, P, false>(io, nseq, st); break;
                Z(7, 32) Z(8, 16) Z(9, 8) Z(10, 4) Z(11, 2) Z(12, 1)
#undef Z
                default: break;
            }
        }
    }
    if (io.rowstats != nullptr) { set_error("rows_r2c: row-line detrend needs the two-rows-per-thread kernel"); return -2; }
    switch (log2M) {
#define X(K) case K: return launch_rows<T, K, rows_seq_fused<K, cmin(TypeCfg<T>::LOGE, K)>()>(io, nseq, st);
        XRFTB_ROWS_CASES(X)
#undef X
        case 14:
            if constexpr (TypeCfg<T>::MAX_ROWS_LOG2 >= 14) return launch_rows<T, 14, 1>(io, nseq, st);
        default: break;
    }
    set_error("rows_r2c: unsupported half length 2^%d", log2M);
    return -2;
}

template <typename T>
int rows_c2c_power(const RowsC2CPower<T>& io, int log2L, long nseq, cudaStream_t st) {
    // float32: two rows per thread, PAIRS row pairs per 256-thread CTA
    if constexpr (sizeof(T) == 4) {
        switch (log2L) {
#define Z(K, P) case K: return launch_rows2c_power<T, K, P>(io, nseq, st);
            Z(11, 2) Z(12, 1) Z(13, 1)   // measured: shorter rows are faster one row per thread
#undef Z
            default: break;
        }
    }
    switch (log2L) {
#define X(K) case K: return launch_rows<T, K, rows_seq_generic<K, cmin(TypeCfg<T>::LOGE, K)>()>(io, nseq, st);
        XRFTB_ROWS_CASES(X)
#undef X
        case 14:
            if constexpr (TypeCfg<T>::MAX_ROWS_LOG2 >= 14) return launch_rows<T, 14, 1>(io, nseq, st);
        default: break;
    }
    set_error("rows_c2c_power: unsupported length 2^%d", log2L);
    return -2;
}

// pass 2 of the columns-first order on packed column spectra (float32; M = Nx/2 in 2^9 .. 2^12)
template <typename T>
int rows_z_power(RowsZPower<T> io, int log2M, long nseq, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        io.tw2 = twiddle_fft<T>(log2M + 1);
        if (!io.tw2) return -3;
        switch (log2M) {
#define Z(K, P) case K: return launch_rowsz_power<T, K, P>(io, nseq, st);
            Z(9, 8) Z(10, 4) Z(11, 2) Z(12, 1)
#undef Z
            default: break;
        }
    }
    set_error("rows_z_power: unsupported half length 2^%d", log2M);
    return -2;
}

// two-field pass 2 of the z-mode chain (float32; mode = EPI_CROSS, EPI_PHASE or EPI_CROSS_AND_PHASE)
template <typename T>
int rows_z_cross(RowsZCross<T> io, int log2M, long nseq, int mode, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        io.tw2 = twiddle_fft<T>(log2M + 1);
        if (!io.tw2) return -3;
        switch (log2M) {
#define Z(K, P) case K: return mode == EPI_PHASE ? launch_rowszx<K, P, EPI_PHASE>(io, nseq, st) \
                             : mode == EPI_CROSS ? launch_rowszx<K, P, EPI_CROSS>(io, nseq, st) : launch_rowszx<K, P, EPI_CROSS_AND_PHASE>(io, nseq, st);
            Z(9, 4) Z(10, 2) Z(11, 1)
#undef Z
            default: break;
        }
    }
    set_error("rows_z_cross: unsupported half length 2^%d", log2M);
    return -2;
}

template <typename T>
int rows_c2r(const cplx<T>* in, long in_stride, T* out, long out_stride, int log2M, long nseq, T scale, cudaStream_t st, const RowsC2R<T>* extra) {
    RowsC2R<T> io{};
    if (extra) io = *extra;
    io.in = in; io.in_stride = in_stride; io.out = out; io.out_stride = out_stride; io.scale = scale; io.tw_r2c = twiddle_r2c<T>(log2M + 1);
    if (!io.tw_r2c) return -3;
    switch (log2M) {
#define X(K) case K: return launch_rows<T, K, rows_seq_generic<K, cmin(TypeCfg<T>::LOGE, K)>()>(io, nseq, st);
        XRFTB_ROWS_CASES(X)
#undef X
        case 14:
            if constexpr (TypeCfg<T>::MAX_ROWS_LOG2 >= 14) return launch_rows<T, 14, 1>(io, nseq, st);
        default: break;
    }
    set_error("rows_c2r: unsupported half length 2^%d", log2M);
    return -2;
}

}  // namespace xrftb
