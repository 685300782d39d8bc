// xrft_b200 -- strided-axis pass dispatch.
#pragma once
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
int cols_c2c(const cplx<T>* in, cplx<T>* out, int log2L, long A, long B, int inverse, T scale, cudaStream_t st) {
    const int C = cols_tile_width<T>(log2L, false);
    if (C < 1) { set_error("cols_c2c: unsupported length 2^%d", log2L); return -2; }
    const long tpr = (B + C - 1) / C;
    ColsC2C<T> io{in, out, B, tpr, inverse, scale};
    switch (log2L) {
#define X(K) case K: return launch_cols<T, K, TileC<T, K, false>::value>(io, A * tpr, st);
        XRFTB_COLS_CASES(X)
#undef X
        default: break;
    }
    set_error("cols_c2c: unsupported length 2^%d", log2L);
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
        IO io{in1, in2, ntile, d, {}};
        if (tmap) io.tmap = *tmap;
        const size_t extra = IO::kBins ? (size_t)d.nbins * (IO::kCplxStage ? 2 : 1) * sizeof(double) : 0;
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
