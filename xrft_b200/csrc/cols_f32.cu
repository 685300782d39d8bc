#include "cols_impl.cuh"
namespace xrftb {
template int cols_c2c<float>(const float2*, float2*, int, long, long, int, float, cudaStream_t, const ColsC2C<float>*);
template int cols_r2c_pack<float>(const ColsR2CPack<float>&, int, int, long, bool, cudaStream_t);
}
