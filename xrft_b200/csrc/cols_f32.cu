#include "cols_impl.cuh"
namespace xrftb {
template int cols_c2c<float>(const float2*, float2*, int, long, long, int, float, cudaStream_t);
}
