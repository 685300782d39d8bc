// xrft_b200 -- internal C++ interface between the C-ABI (api.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include "fft_kernels.cuh"

namespace xrftb {

void set_error(const char* fmt, ...);
int sm_count();
int check_launch(const char* what);

// device-resident twiddle tables, computed in double on the host, cached per (device, length)
template <typename T> const cplx<T>* twiddle_fft(int log2L);  // exp(-2 pi i m / L), m in [0, L)
template <typename T> const cplx<T>* twiddle_r2c(int log2N);  // exp(-2 pi i k / N), k in [0, N/2]
template <typename T> const cplx<T>* twiddle_head(int log2N); // exp(-2 pi i m / N), m in [0, min(N, 8192))

template <typename T> struct TypeCfg;
#ifndef XRFTB_F32_LOGE
#define XRFTB_F32_LOGE 4
#endif
template <> struct TypeCfg<float> { static constexpr int LOGE = XRFTB_F32_LOGE; static constexpr int V = 2; static constexpr int TILE_POINTS = 16384; static constexpr int CMAX = 16; static constexpr int MAX_ROWS_LOG2 = 14; static constexpr int MAX_COLS_LOG2 = 13; };
template <> struct TypeCfg<double> { static constexpr int LOGE = 3; static constexpr int V = 1; static constexpr int TILE_POINTS = 8192; static constexpr int CMAX = 8; static constexpr int MAX_ROWS_LOG2 = 13; static constexpr int MAX_COLS_LOG2 = 13; };

// tile width (columns per CTA) of the strided pass for a given length
template <typename T> inline int cols_tile_width(int log2L, bool two_fields) {
    int c = TypeCfg<T>::TILE_POINTS >> log2L;
    if (c > TypeCfg<T>::CMAX) c = TypeCfg<T>::CMAX;
    if (two_fields) c >>= 1;
    return c;  // 0 => unsupported
}

// Chain-selection options (include/xrft_b200.h: xrftb_set_option / xrftb_get_option).  Each one names a kernel chain that
// also serves other shapes or dtypes as the regular path, so every setting is a tested configuration, not an experiment.
enum Option : int { OPT_COLS_FIRST = 0, OPT_ZPACK, OPT_ZTMA, OPT_COLS_ASYNC, OPT_ROWLINE, OPT_CROSS_Z, OPT_BINS_STATIC, OPT_COUNT };
int option(Option o);
inline bool bins_static_enabled() { return option(OPT_BINS_STATIC) != 0; }
// pass 1 of the columns-first order (ColsR2CPack): tile width in PACKED columns
template <typename T> inline int colsfirst_tile_width(int log2L) { return cols_tile_width<T>(log2L, false); }

// the two-rows-per-thread row kernel (float32, blocked output) handles half lengths 2^7..2^12 with 2^(12 - log2M) row pairs
// per CTA; the 2 * pairs rows of a CTA must be consecutive rows of one item
inline bool rows2_eligible(int log2M, int logNy) { return log2M >= 7 && log2M <= 12 && logNy >= (12 - log2M) + 1; }

// `extra` (nullable) carries the optional members of RowsC2C (the chirp-z hooks); its in / out / strides / inverse / scale are ignored
template <typename T> int rows_c2c(const cplx<T>* in, cplx<T>* out, int log2L, long nseq, long in_stride, long out_stride,
                                   int inverse, T scale, cudaStream_t st, const RowsC2C<T>* extra = nullptr);
template <typename T> int rows_r2c(RowsR2CFused<T> io, int log2M, long nseq, cudaStream_t st);
template <typename T> int rows_c2r(const cplx<T>* in, long in_stride, T* out, long out_stride, int log2M, long nseq, T scale,
                                   cudaStream_t st, const RowsC2R<T>* extra = nullptr);
template <typename T> int rows_c2c_power(const RowsC2CPower<T>& io, int log2L, long nseq, cudaStream_t st);
template <typename T> int rows_z_power(RowsZPower<T> io, int log2M, long nseq, cudaStream_t st);
template <typename T> int rows_z_cross(RowsZCross<T> io, int log2M, long nseq, int mode, cudaStream_t st);
// pass 2 of the columns-first order with the radial-bin epilogue (float32); returns 1 when the shape is not covered
int rows_bins(const RowsBins& io, int log2L, cudaStream_t st);
bool rows_bins_shape_ok(int log2L, int ny);
int rows_zx_bins(RowsZCrossBins io, int log2M, cudaStream_t st);
bool rows_zx_bins_shape_ok(int log2M, int ny);
// half lengths the z-mode pass 2 is dispatched for (Nx = 1024 .. 4096: the sizes covered by the GPU parity tests; the
// 2^12 instantiation exists but stays off until it has been through them)
inline bool rows_z_supported(int log2M) { return log2M >= 9 && log2M <= 11; }
template <typename T> int cols_r2c_pack(const ColsR2CPack<T>& io, int log2L, int C, long ntiles, bool use_async, cudaStream_t st);
// `extra` (nullable) carries the optional members of ColsC2C: the four-step hooks (twiddle product on the stores, transposed
// store) and the xrftb_fft2r hooks (row predicate / roll / ramps / crop); its in / out / B / inverse / scale are ignored
template <typename T> int cols_c2c(const cplx<T>* in, cplx<T>* out, int log2L, long A, long B, int inverse, T scale,
                                   cudaStream_t st, const ColsC2C<T>* extra = nullptr);
// smooth.cu: mixed-radix transform of the lengths 2^a 3^b 5^c 7^d that are not powers of two, along the middle axis of an
// [A][n][B] view (in-place safe); smooth_len_ok = the length is covered (factors and shared-memory capacity)
template <typename T> bool smooth_len_ok(long n);
template <typename T> int smooth_c2c(const cplx<T>* src, cplx<T>* dst, long A, long n, long B, int inverse, T scale, cudaStream_t st);
// real transforms of even length N (N / 2 covered by smooth_len_ok) along the contiguous axis: [nseq][N] real <-> [nseq][N/2+1]
// complex in one pass (packed half-length transform + split in the stores / merge in the loads); c2r scales by `scale`
template <typename T> bool smooth_real_ok(long N);
template <typename T> int smooth_r2c(const T* in, cplx<T>* out, long nseq, long N, cudaStream_t st);
template <typename T> int smooth_c2r(const cplx<T>* in, T* out, long nseq, long N, T scale, cudaStream_t st);
// one explicit instantiation per (T, MODE), spread over several translation units
template <typename T, int MODE> int cols_fused_mode(const cplx<T>* in1, const cplx<T>* in2, int log2L, long ntiles_total, int ntile,
                                                    const EpilogueDesc& d, const CUtensorMap* tmap, cudaStream_t st);
template <typename T> inline int cols_fused(int mode, const cplx<T>* in1, const cplx<T>* in2, int log2L, long ntiles_total, int ntile,
                                            const EpilogueDesc& d, const CUtensorMap* tmap, cudaStream_t st) {
    switch (mode) {
        case EPI_COMPLEX: return cols_fused_mode<T, EPI_COMPLEX>(in1, in2, log2L, ntiles_total, ntile, d, tmap, st);
        case EPI_POWER: return cols_fused_mode<T, EPI_POWER>(in1, in2, log2L, ntiles_total, ntile, d, tmap, st);
        case EPI_CROSS: return cols_fused_mode<T, EPI_CROSS>(in1, in2, log2L, ntiles_total, ntile, d, tmap, st);
        case EPI_PHASE: return cols_fused_mode<T, EPI_PHASE>(in1, in2, log2L, ntiles_total, ntile, d, tmap, st);
        case EPI_BINS_POWER: return cols_fused_mode<T, EPI_BINS_POWER>(in1, in2, log2L, ntiles_total, ntile, d, tmap, st);
        case EPI_BINS_CROSS: return cols_fused_mode<T, EPI_BINS_CROSS>(in1, in2, log2L, ntiles_total, ntile, d, tmap, st);
    }
    set_error("cols_fused: bad mode %d", mode);
    return -1;
}

}  // namespace xrftb
