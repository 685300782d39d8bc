// xrft_b200 -- strided-axis pass dispatch.
#pragma once
#include <cstdlib>
#include "launch.cuh"

namespace xrftb {

// tile width is a pure function of (dtype, length, #fields): see cols_tile_width()
template <typename T, int LOG2L, bool TWO> struct TileC {
    static constexpr int raw = cmin(TypeCfg<T>::CMAX, TypeCfg<T>::TILE_POINTS >> LOG2L);
    static constexpr int value = TWO ? raw / 2 : raw;
};

#define XRFTB_COLS_CASES(X) \
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)

template <typename T>
int cols_c2c(const cplx<T>* in, cplx<T>* out, int log2L, long A, long B, int inverse, T scale, cudaStream_t st, const ColsC2C<T>* extra) {
    const int C = cols_tile_width<T>(log2L, false);
    if (C < 1) { set_error("cols_c2c: unsupported length 2^%d", log2L); return -2; }
    const long tpr = (B + C - 1) / C;
    ColsC2C<T> io{};
    if (extra) io = *extra;
    io.in = in; io.out = out; io.B = B; io.tiles_per_row = tpr; io.inverse = inverse; io.scale = scale;
    switch (log2L) {
#define X(K) case K: return launch_cols<T, K, TileC<T, K, false>::value>(io, A * tpr, st);
        XRFTB_COLS_CASES(X)
#undef X
        default: break;
    }
    set_error("cols_c2c: unsupported length 2^%d", log2L);
    return -2;
}

template <typename T>
int cols_r2c_pack(const ColsR2CPack<T>& io_, int log2L, int C, long ntiles, bool use_async, cudaStream_t st) {
    ColsR2CPack<T> io = io_;
    if (io.tiles_per_item < 1 || (io.tiles_per_item & (io.tiles_per_item - 1))) { set_error("cols_r2c_pack: %d tiles per item (a power of two is required)", io.tiles_per_item); return -2; }
    io.log_tpi = 0;
    while ((1 << io.log_tpi) < io.tiles_per_item) ++io.log_tpi;
    if constexpr (sizeof(T) == 4) {
        if (use_async) {   // tensor-map fed variant (io.tmap / io.box_rows are set); float32, two packed columns per thread
            switch (log2L) {
#define X(K) case K: if constexpr (K > TypeCfg<T>::LOGE && TileC<T, K, false>::value >= 2) return launch_cols_async<T, K, TileC<T, K, false>::value>(io, ntiles, st); break;
                XRFTB_COLS_CASES(X)
#undef X
                default: break;
            }
            set_error("cols_r2c_pack: no tensor-map variant for length 2^%d", log2L);
            return -2;
        }
    }
    if (C != cols_tile_width<T>(log2L, false)) { set_error("cols_r2c_pack: tile width %d needs the tensor-map fed kernel", C); return -2; }
    switch (log2L) {
#define X(K) case K: if constexpr (TileC<T, K, false>::value >= 1) return launch_cols<T, K, TileC<T, K, false>::value>(io, ntiles, st); break;
        XRFTB_COLS_CASES(X)
#undef X
        default: break;
    }
    set_error("cols_r2c_pack: unsupported length 2^%d", log2L);
    return -2;
}

template <typename T, int K, int MODE>
static int cols_fused_k(const cplx<T>* in1, const cplx<T>* in2, long ntiles_total, int ntile, const EpilogueDesc& d, const CUtensorMap* tmap, cudaStream_t st) {
    using IO = ColsFused<T, MODE>;
    constexpr int C = TileC<T, K, IO::kTwoFields>::value;
    if constexpr (C < 1) {
        set_error("cols_fused: length 2^%d too long for %d field(s)", K, IO::kTwoFields ? 2 : 1);
        return -2;
    } else {
        IO io{in1, in2, ntile, d, 0, reinterpret_cast<const cplx<T>*>(d.fix_ag), reinterpret_cast<const cplx<T>*>(d.fix_wj), {}};
        io.hist_off = IO::template hist_offset_bytes<K, cmin(TypeCfg<T>::LOGE, K), C, 1>();
        if (tmap) io.tmap = *tmap;
        const size_t extra = IO::kBins ? (size_t)d.nbins * (IO::kCplxStage ? 2 : 1) * sizeof(double) : 0;
        // isotropic power spectrum, the usual case (radial bins: symmetric LUT, <= 255 bins, full-width counting, no one-sided
        // weights): the kernel that keeps each thread's bin indices in registers
        if constexpr (sizeof(T) == 4 && MODE == EPI_BINS_POWER && (K > TypeCfg<T>::LOGE) && C >= 4 && C % 4 == 0) {
            if (d.lut_symmetric && d.full && !d.weight_x && d.nbins <= 255 && bins_static_enabled()) {
                const int rc = launch_cols_bins<K, C>(io, ntiles_total, st);
                if (rc <= 0) return rc;
            }
        }
        // real-valued single-field epilogues of the float32 path: bulk-copy fed variant (the staging buffer fits the
        // half-size exchange buffer).  XRFTB_COLS_ASYNC=0 selects the register-prefetch kernel.
        if constexpr (sizeof(T) == 4 && (MODE == EPI_POWER || MODE == EPI_BINS_POWER) && (K > TypeCfg<T>::LOGE) && C >= 2 && C % 2 == 0) {
            // measured (profiles/README.md): wins where the epilogue stores are asynchronous too (TMA tensor stores, C >= 8);
            // at C = 4 the 16-byte row-segment stores dominate and the extra exchange barriers cost more than the loads hide
            const int async_on = option(OPT_COLS_ASYNC);
            // the binned epilogue has no stores to hide: there the register-prefetch kernel is 36 % faster (config 4: 111 vs 81 GPoints/s)
            if (async_on == 1 || (async_on == 2 && C >= 8 && MODE == EPI_POWER)) return launch_cols_async<T, K, C>(io, ntiles_total, st, extra);
        }
        return launch_cols<T, K, C>(io, ntiles_total, st, extra);
    }
}

template <typename T, int MODE>
int cols_fused_mode(const cplx<T>* in1, const cplx<T>* in2, int log2L, long ntiles_total, int ntile, const EpilogueDesc& d,
                    const CUtensorMap* tmap, cudaStream_t st) {
    switch (log2L) {
#define X(K) case K: return cols_fused_k<T, K, MODE>(in1, in2, ntiles_total, ntile, d, tmap, st);
        XRFTB_COLS_CASES(X)
#undef X
        default: break;
    }
    set_error("cols_fused: unsupported length 2^%d", log2L);
    return -2;
}

}  // namespace xrftb
