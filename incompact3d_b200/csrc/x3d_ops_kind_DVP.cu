// instantiations of the line kernels for operator kind DVP (see x3d_ops_kernels.cuh)
#include "x3d_ops_inst.cuh"
namespace x3d {
void launch_kind_DVP(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  launch_kind_nt<DVP, 2>(ctx, op, g, T, u, t);
}
}  // namespace x3d
