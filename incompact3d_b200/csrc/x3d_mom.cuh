// x3d_mom.cuh -- fused momentum-RHS kernels, host interface (see x3d_mom_kernels.cuh)
#pragma once
#include "x3d_ctx.cuh"

namespace x3d {

struct MomTable {        // compressed coefficient table of one periodic operator
  double *d_c = nullptr;     // [3][MOM_TABS][L] double2
  double *d_scan = nullptr;  // [10][32]
  int L = 0, nc = 0;
  double ff = 0.0;           // constant super-diagonal of the periodic operator
  bool ok = false;
  ~MomTable();
};
bool build_mom_table(Ctx &ctx, const TriTable &T, MomTable &M);
// table-free cyclic solves (pair_solve_cyclic, x3d_mom_kernels.cuh) are possible for this scheme and chunking
bool mom_cyclic_ok(double alpha, int n, int L);
// overlap-save segmentation of a periodic line of ntot rows for the fused kernels (nseg = 1: the line fits a tile)
bool mom_segments(int ntot, double alpha1, double alpha2, int &S, int &H, int &nseg);
// time integration folded into the x kernel: N = sum (+ extra) + r_x ; u <- ca N + cb old_in + u ; old_out <- N
struct MomIntt {
  const double *sum[3], *extra[3], *old_in[3];
  double *u[3], *old_out[3];
  double *u_out[3];       // = u, or a separate array when the x lines are segmented (neighbouring tiles read the old u)
  double ca, cb;
  bool use_old, store_old, has_extra;
};
// y / z lines: out[c] (+)= xnu D2(f[c]) - 1/2 (D1(f[c] f[axis]) + f[axis] D1(f[c]))
void launch_mom_pair(Ctx &ctx, int axis, const DevOp &op1, const DevOp &op2, const MomTable &M1, const MomTable &M2, double xnu,
                     const double *const f[3], double *const out[3], long long n1, int nline, long long nouter, long long sline,
                     long long souter, bool add = false, bool cyclic = false);
bool mom_pair_eligible(int n, int L);
// x lines (contiguous): f / out are (n, nlines) arrays
void launch_mom_x(Ctx &ctx, const DevOp &op1, const DevOp &op2, const MomTable &M1, const MomTable &M2, double xnu,
                  const double *const f[3], double *const out[3], int n, long long nlines, bool add = false, bool cyclic = false,
                  const MomIntt *intt = nullptr);
bool mom_x_eligible(int n, int L);

// z lines of slabs without transposes (x3d_slab_kernels.cuh): zero-carry solves of the slab's own rows + carry planes, then the
// face corrections.  f / out: the slab (nlanes, n rows); halo[c]: (nlanes, 16) with rows 4..7 = the 4 planes below, 8..11 = above;
// carry: [18][nlanes] (Yout of the 9 systems, then Z0)
bool mom_slab_eligible(double alpha1, double alpha2, long long nlanes, int n);
void launch_mom_slab(Ctx &ctx, const DevOp &op1, const DevOp &op2, double xnu, const double *const f[3], const double *const halo[3],
                     double *const out[3], long long nlanes, int n, bool add, double *carry);
struct ZFix {
  DevBuf tab;     // [4][n]: A (D1), B (D1), A (D2), B (D2)
  int n = 0, W = 0;
  double k1 = 0.0, k2 = 0.0;
};
void build_zfix(Ctx &ctx, const DevOp &op1, const DevOp &op2, double xnu, int n, ZFix &Z);
void launch_zfix(Ctx &ctx, const ZFix &Z, const double *yin, const double *z0n, const double *yout, double *const sum[3], const double *a,
                 long long nlanes);

// fused pairs of periodic staggered operators on y / z lines (x3d_stag_kernels.cuh, x3d_stag.cu)
//   mode 0: outA = opA(inA) + opB(inB), (opA, opB) = (inter?vp, der?vp);  mode 1: outA = opA(inA), outB = opB(inA), (inter?pv, der?pv)
bool stag_pair_eligible(Ctx &ctx, const DevOp &opA, const DevOp &opB, long long n1, int nline, long long sline, long long souter, long long nouter);
void launch_stag_pair(Ctx &ctx, int mode, int axis, const DevOp &opA, const DevOp &opB, const double *inA, const double *inB, double *outA,
                      double *outB, long long n1, int nline, long long nouter, long long sline, long long souter);
void stag_release(Ctx *ctx);

}  // namespace x3d
