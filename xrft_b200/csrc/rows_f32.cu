#include "rows_impl.cuh"
namespace xrftb {
template int rows_c2c<float>(const float2*, float2*, int, long, long, long, int, float, cudaStream_t);
template int rows_r2c<float>(RowsR2CFused<float>, int, long, cudaStream_t);
template int rows_c2r<float>(const float2*, long, float*, long, int, long, float, cudaStream_t);
template int rows_c2c_power<float>(const RowsC2CPower<float>&, int, long, cudaStream_t);
template int rows_z_power<float>(RowsZPower<float>, int, long, cudaStream_t);
template int rows_z_cross<float>(RowsZCross<float>, int, long, int, cudaStream_t);
}
