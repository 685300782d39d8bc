#include "cols_impl.cuh"
namespace xrftb {
template int cols_c2c<double>(const double2*, double2*, int, long, long, int, double, cudaStream_t, const ColsC2C<double>*);
template int cols_r2c_pack<double>(const ColsR2CPack<double>&, int, int, long, bool, cudaStream_t);
}
