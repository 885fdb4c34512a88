// instantiations of the line kernels for operator kind FIL (see x3d_ops_kernels.cuh)
#include "x3d_ops_inst.cuh"
namespace x3d {
void launch_kind_FIL(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  launch_kind_nt<FIL, 3>(ctx, op, g, T, u, t);
}
}  // namespace x3d
