#include "rows_impl.cuh"
namespace xrftb {
template int rows_c2c<float>(const float2*, float2*, int, long, long, long, int, float, cudaStream_t, const RowsC2C<float>*);
template int rows_r2c<float>(RowsR2CFused<float>, int, long, cudaStream_t);
template int rows_c2r<float>(const float2*, long, float*, long, int, long, float, cudaStream_t, const RowsC2R<float>*);
template int rows_c2c_power<float>(const RowsC2CPower<float>&, int, long, cudaStream_t);
template int rows_z_power<float>(RowsZPower<float>, int, long, cudaStream_t);
template int rows_z_cross<float>(RowsZCross<float>, int, long, int, cudaStream_t);
}

namespace xrftb {
// pass 2 of the columns-first order with the radial-bin epilogue (float32, Nx = 64 .. 8192); 1 = shape not covered
int rows_bins(const RowsBins& io, int log2L, cudaStream_t st) {
    switch (log2L) {
#define Z(K) case K: return launch_rows_bins<K, rows_seq_generic<K, TypeCfg<float>::LOGE>()>(io, st);
        Z(6) Z(7) Z(8) Z(9) Z(10) Z(11) Z(12) Z(13)
#undef Z
        default: break;
    }
    return 1;
}
bool rows_bins_shape_ok(int log2L, int ny) {
    int seq = 0;
    switch (log2L) {
#define Z(K) case K: seq = rows_seq_generic<K, 4>(); break;
        Z(6) Z(7) Z(8) Z(9) Z(10) Z(11) Z(12) Z(13)
#undef Z
        default: return false;
    }
    return ny / 2 >= seq && (ny / 2) % seq == 0;
}
// two-field pass 2 of the z-mode chain with the radial-bin epilogue (float32, M = Nx/2 in 2^9 .. 2^11); 1 = shape not covered
int rows_zx_bins(RowsZCrossBins io, int log2M, cudaStream_t st) {
    io.base.tw2 = twiddle_fft<float>(log2M + 1);
    if (!io.base.tw2) return -3;
    switch (log2M) {
#define Z(K, P) case K: return launch_rowszx_bins<K, P>(io, st);
        Z(9, 4) Z(10, 2) Z(11, 1)
#undef Z
        default: break;
    }
    return 1;
}
bool rows_zx_bins_shape_ok(int log2M, int ny) {
    const int rows = log2M == 9 ? 4 : log2M == 10 ? 2 : log2M == 11 ? 1 : 0;
    return rows > 0 && ny / 2 >= rows && (ny / 2) % rows == 0;
}
}  // namespace xrftb
