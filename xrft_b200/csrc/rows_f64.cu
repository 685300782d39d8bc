#include "rows_impl.cuh"
namespace xrftb {
template int rows_c2c<double>(const double2*, double2*, int, long, long, long, int, double, cudaStream_t, const RowsC2C<double>*);
template int rows_r2c<double>(RowsR2CFused<double>, int, long, cudaStream_t);
template int rows_c2r<double>(const double2*, long, double*, long, int, long, double, cudaStream_t, const RowsC2R<double>*);
template int rows_c2c_power<double>(const RowsC2CPower<double>&, int, long, cudaStream_t);
template int rows_z_power<double>(RowsZPower<double>, int, long, cudaStream_t);
template int rows_z_cross<double>(RowsZCross<double>, int, long, int, cudaStream_t);
}
