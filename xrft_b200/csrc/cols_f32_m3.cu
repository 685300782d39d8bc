#include "cols_impl.cuh"
namespace xrftb {
template int cols_fused_mode<float, EPI_PHASE>(const float2*, const float2*, int, long, int, const EpilogueDesc&, const CUtensorMap*, cudaStream_t);
}
