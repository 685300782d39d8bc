#include "cols_impl.cuh"
namespace xrftb {
template int cols_fused_mode<double, EPI_PHASE>(const double2*, const double2*, int, long, int, const EpilogueDesc&, const CUtensorMap*, cudaStream_t);
}
