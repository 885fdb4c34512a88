// instantiations of the line kernels for operator kind DPV (see x3d_ops_kernels.cuh)
#include "x3d_ops_inst.cuh"
namespace x3d {
void launch_kind_DPV(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  launch_kind_nt<DPV, 2>(ctx, op, g, T, u, t);
}
}  // namespace x3d
