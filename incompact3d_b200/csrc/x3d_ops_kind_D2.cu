// instantiations of the line kernels for operator kind D2 (see x3d_ops_kernels.cuh)
#include "x3d_ops_inst.cuh"
namespace x3d {
void launch_kind_D2(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  if (op.c[2] != 0.0 || op.c[3] != 0.0) launch_kind_nt<D2, 4>(ctx, op, g, T, u, t);
  else launch_kind_nt<D2, 2>(ctx, op, g, T, u, t);
}
}  // namespace x3d
