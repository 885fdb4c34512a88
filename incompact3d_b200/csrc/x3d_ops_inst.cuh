// x3d_ops_inst.cuh -- per-kind launch dispatch; included by the x3d_ops_kind*.cu files so
// that the (fully unrolled) kernel instantiations compile in parallel.
#pragma once
#include "x3d_ctx.cuh"
#include "x3d_ops_kernels.cuh"

namespace x3d {

struct LineGeom {
  int axis;
  int n1;             // strided: extent of the coalesced (lane) direction
  long long nouter;   // strided: number of outer slabs (gridDim.y)
  long long sin, sout, oin, oout;
  long long nlines;   // contiguous: number of lines
};

template <int KIND, int NT, int L, int LX, int NCMAX, int MINB>
static void launch_strided_one(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  auto kern = k_strided<KIND, NT, L, LX, NCMAX, MINB>;
  static bool configured = false;
  if (!configured) {
    // leave most of the unified L1/shared array to L1 (coefficient rows are served from it)
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 25));
    configured = true;
  }
  dim3 block(LX, T.nc);
  dim3 grid(static_cast<unsigned>((g.n1 + LX - 1) / LX), static_cast<unsigned>(g.nouter));
  kern<<<grid, block, 0, ctx.stream>>>(op, u, t, T.d_rows, T.d_chunk, T.nc, g.n1, g.sin, g.sout, g.oin, g.oout);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

template <int KIND, int NT>
static void launch_strided(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  const int v = ctx.strided_variant;
  if (T.L == 8 && T.nc <= 16) launch_strided_one<KIND, NT, 8, 32, 16, 2>(ctx, op, g, T, u, t);
  else if (T.L == 16 && T.nc <= 16) launch_strided_one<KIND, NT, 16, 32, 16, 2>(ctx, op, g, T, u, t);
  else if (T.L == 16 && T.nc <= 32 && v == 3) launch_strided_one<KIND, NT, 16, 32, 32, 1>(ctx, op, g, T, u, t);
  else if (T.L == 16 && T.nc <= 32) launch_strided_one<KIND, NT, 16, 16, 32, 2>(ctx, op, g, T, u, t);
  else if (T.L == 32 && T.nc <= 16 && v == 0) launch_strided_one<KIND, NT, 32, 32, 16, 1>(ctx, op, g, T, u, t);
  else if (T.L == 32 && T.nc <= 16) launch_strided_one<KIND, NT, 32, 16, 16, 2>(ctx, op, g, T, u, t);
  else if (T.L == 32 && T.nc <= 32) launch_strided_one<KIND, NT, 32, 16, 32, 1>(ctx, op, g, T, u, t);
  else throw Error("no strided kernel for this line length (n <= 1024 supported)");
}

template <int KIND, int NT, int L, int WPB, int NB, int MINB, bool TMA>
static void launch_contig_one(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  const int NP = T.nc * L;
  int NBUF = (NP > op.n_in ? NP : op.n_in) + 2 * HALO;
  NBUF = (NBUF + 1) & ~1;
  const int COEF = (7 * NP + 1) & ~1;
  const size_t smem = static_cast<size_t>(COEF + WPB * (NB * NBUF + 8) + WPB * NB) * sizeof(double);
  auto kern = k_contig<KIND, NT, L, WPB, NB, MINB, TMA>;
  static size_t configured = 0;
  if (smem > configured) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  long long blocks = (g.nlines + WPB - 1) / WPB;
  const long long cap = static_cast<long long>(ctx.sm_count) * MINB;
  if (blocks > cap) blocks = cap;
  kern<<<static_cast<unsigned>(blocks), 32 * WPB, smem, ctx.stream>>>(op, u, t, T.d_rows, T.d_scan, T.nc, g.nlines, NP, NBUF, COEF);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

template <int KIND, int NT, int L>
static void launch_contig_L(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  const bool even = (op.n_in % 2 == 0) && (op.n_out % 2 == 0);
  const bool aligned = (reinterpret_cast<uintptr_t>(u) % 16 == 0) && (reinterpret_cast<uintptr_t>(t) % 16 == 0);
  const int v = (even && aligned) ? ctx.contig_variant : 0;
  if constexpr (L == 9 || L == 17) {
    if (v == 1) return launch_contig_one<KIND, NT, L, 8, 2, 2, true>(ctx, op, g, T, u, t);
    if (v == 2) return launch_contig_one<KIND, NT, L, 8, 3, 1, true>(ctx, op, g, T, u, t);
    if (v == 3) return launch_contig_one<KIND, NT, L, 4, 3, 3, true>(ctx, op, g, T, u, t);
  }
  if constexpr (L <= 17) launch_contig_one<KIND, NT, L, 8, 1, 2, false>(ctx, op, g, T, u, t);
  else launch_contig_one<KIND, NT, L, 8, 1, 1, false>(ctx, op, g, T, u, t);
}

template <int KIND, int NT>
static void launch_contig(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  switch (T.L) {
    case 5: launch_contig_L<KIND, NT, 5>(ctx, op, g, T, u, t); break;
    case 9: launch_contig_L<KIND, NT, 9>(ctx, op, g, T, u, t); break;
    case 17: launch_contig_L<KIND, NT, 17>(ctx, op, g, T, u, t); break;
    case 25: launch_contig_L<KIND, NT, 25>(ctx, op, g, T, u, t); break;
    case 33: launch_contig_L<KIND, NT, 33>(ctx, op, g, T, u, t); break;
    default: throw Error("no contiguous kernel for this line length (n <= 1056 supported)");
  }
}

template <int KIND, int NT>
static void launch_kind_nt(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  if (g.axis == 0) launch_contig<KIND, NT>(ctx, op, g, T, u, t);
  else launch_strided<KIND, NT>(ctx, op, g, T, u, t);
}

}  // namespace x3d
