// x3d_ops_inst.cuh -- per-kind launch dispatch; included by the x3d_ops_kind*.cu files so
// that the (fully unrolled) kernel instantiations compile in parallel.
#pragma once
#include "x3d_ctx.cuh"
#include <algorithm>
#include "x3d_ops_kernels.cuh"

namespace x3d {

struct LineGeom {
  int axis;
  bool pair;          // strided: use the warp-per-lane-pair kernel (k_pair); T.L is then odd
  int n1;             // strided: extent of the coalesced (lane) direction
  long long nouter;   // strided: number of outer slabs (gridDim.y)
  long long sin, sout, oin, oout;
  long long nlines;   // contiguous: number of lines
};

template <int KIND, int NT, int L, int LX, int NCMAX, int MINB>
static void launch_strided_one(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  auto kern = k_strided<KIND, NT, L, LX, NCMAX, MINB>;
  static bool configured_dev[64] = {};     // the attribute is per device
  bool &configured = configured_dev[ctx.device & 63];
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

// ---- tensor-map TMA tiles ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    X3D_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) throw Error("cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// 3-D view (lanes, line, outer) of a pencil with element strides (1, sline, souter); box = (lx, br, 1)
inline CUtensorMap make_line_map(const double *base, long long n1, int nline, long long nouter, long long sline, long long souter,
                                 int lx, int br, bool swizzle128 = false) {
  alignas(64) CUtensorMap m;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(n1), static_cast<cuuint64_t>(nline), static_cast<cuuint64_t>(nouter)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(sline) * 8u, static_cast<cuuint64_t>(nouter > 1 ? souter : sline * nline) * 8u};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(lx), static_cast<cuuint32_t>(br), 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error("cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ")");
  return m;
}
inline bool tile_eligible(const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, const double *t) {
  if (T.nc > 32) return false;
  if ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(t)) & 15u) return false;
  if ((g.sin | g.sout) & 1) return false;                       // TMA strides are multiples of 16 bytes
  if (g.nouter > 1 && ((g.oin | g.oout) & 1)) return false;
  (void)op;
  return true;
}

template <int KIND, int NT, int L, int LX, int NB, int MINB>
static void launch_tile_one(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  auto kern = k_tile<KIND, NT, L, LX, NB, MINB>;
  TileGeom tg{};
  auto boxes = [](int n, int &nbox, int &br) { nbox = (n + 255) / 256; br = (n + nbox - 1) / nbox; br = (br + 1) & ~1; if (br > 256) { ++nbox; br = ((n + nbox - 1) / nbox + 1) & ~1; } };
  boxes(op.n_in, tg.nbox_in, tg.br_in);
  boxes(op.n_out, tg.nbox_out, tg.br_out);
  int rows = std::max(std::max(tg.nbox_in * tg.br_in, tg.nbox_out * tg.br_out), T.nc * L);
  tg.rows_slot = (rows + 15) & ~15;
  tg.nbx = (g.n1 + LX - 1) / LX;
  tg.ntiles = static_cast<long long>(tg.nbx) * g.nouter;
  const size_t smem = (static_cast<size_t>(NB) * tg.rows_slot * LX + 2 * 32 * (LX + 1) + 2 * LX + NB) * sizeof(double);
  static size_t configured_dev[64] = {};   // the attribute is per device: one record per device ordinal
  size_t &configured = configured_dev[ctx.device & 63];
  static int per_sm = 1;
  const int threads = ((LX * T.nc + 31) / 32) * 32;
  if (smem > configured) {
    if (smem > 227 * 1024) throw Error("line too long for the shared-memory tile kernel");
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  X3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) throw Error("tile kernel does not fit on an SM");
  const CUtensorMap mi = make_line_map(u, g.n1, op.n_in, g.nouter, g.sin, g.oin, LX, tg.br_in);
  const CUtensorMap mo = make_line_map(t, g.n1, op.n_out, g.nouter, g.sout, g.oout, LX, tg.br_out);
  long long blocks = static_cast<long long>(ctx.sm_count) * per_sm;
  if (blocks > tg.ntiles) blocks = tg.ntiles;
  kern<<<static_cast<unsigned>(blocks), threads, smem, ctx.stream>>>(op, mi, mo, T.d_rows, T.d_scan, T.nc, tg);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

// ---- warp-per-lane-pair kernel: geometry, eligibility, launch --------------------------------------
constexpr int PAIR_NB = 3;
inline bool pair_boxes(int n, bool exact, int &nbox, int &br) {
  for (nbox = (n + 255) / 256; nbox <= 8; ++nbox) {
    br = (n + nbox - 1) / nbox;
    br = (br + 7) & ~7;
    if (br > 256) continue;
    if (!exact || nbox * br == n) return true;
  }
  return false;
}
// fills pg / smem for (op, geometry, chunk length L); false when the kernel cannot run this call
inline bool pair_plan(const DevOp &op, const LineGeom &g, int L, const double *u, const double *t, PairGeom &pg, size_t &smem) {
  if (L <= 0 || (L & 1) == 0) return false;
  const int nc = (op.n_out + L - 1) / L;
  if (nc > 32 || nc < 1) return false;
  if ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(t)) & 15u) return false;
  if ((g.sin | g.sout) & 1) return false;
  if (g.nouter > 1 && ((g.oin | g.oout) & 1)) return false;
  if (op.untouched) return false;
  pg = PairGeom{};
  pg.halo = op.periodic ? 1 : 0;
  if (op.periodic && ((op.n_in & 7) || op.n_in != op.n_out || op.n_in < 8)) return false;
  if (!pair_boxes(op.n_in, op.periodic != 0, pg.nbox_in, pg.br_in)) return false;
  if (!pair_boxes(op.n_out, false, pg.nbox_out, pg.br_out)) return false;
  const int data_rows = std::max(pg.nbox_in * pg.br_in, pg.nbox_out * pg.br_out);
  pg.slot_rows = 8 + data_rows + 8;
  pg.NP = nc * L;
  pg.nbx = (g.n1 + 15) / 16;
  pg.ntiles = static_cast<long long>(pg.nbx) * g.nouter;
  // the last chunk's window may read up to NP + 2*HALO + 8 physical rows: what follows the last slot
  // (coefficient table, closure rows) must cover the overrun
  const size_t tail = static_cast<size_t>(3) * pg.NP * 16 + PAIR_WARPS * 8 * 16 + 2 * PAIR_NB * 8;
  const long long overrun = static_cast<long long>(pg.NP + 8 + HALO - pg.slot_rows) * 128;
  if (overrun > static_cast<long long>(tail)) return false;
  smem = static_cast<size_t>(PAIR_NB) * pg.slot_rows * 128 + tail;
  return smem <= 227 * 1024;
}

template <int KIND, int NT, int L>
static void launch_pair_one(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  auto kern = k_pair<KIND, NT, L, PAIR_NB>;
  PairGeom pg;
  size_t smem = 0;
  if (!pair_plan(op, g, L, u, t, pg, smem)) throw Error("internal: k_pair launched on an ineligible call");
  static size_t configured_dev[64] = {};   // the attribute is per device: one record per device ordinal
  size_t &configured = configured_dev[ctx.device & 63];
  if (smem > configured) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  int per_sm = 1;
  X3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * (PAIR_WARPS + 1), smem));
  if (per_sm < 1) throw Error("k_pair does not fit on an SM");
  const CUtensorMap mi = make_line_map(u, g.n1, op.n_in, g.nouter, g.sin, g.oin, 16, pg.br_in, true);
  const CUtensorMap mh = make_line_map(u, g.n1, op.n_in, g.nouter, g.sin, g.oin, 16, 8, true);  // 8-row wrap-ghost boxes
  const CUtensorMap mo = make_line_map(t, g.n1, op.n_out, g.nouter, g.sout, g.oout, 16, pg.br_out, true);
  long long blocks = static_cast<long long>(ctx.sm_count) * per_sm;
  if (blocks > pg.ntiles) blocks = pg.ntiles;
  kern<<<static_cast<unsigned>(blocks), 32 * (PAIR_WARPS + 1), smem, ctx.stream>>>(op, mi, mh, mo, T.d_rows, T.d_scan, T.nc, pg);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

template <int KIND, int NT>
static void launch_pair(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  switch (T.L) {
    case 5: launch_pair_one<KIND, NT, 5>(ctx, op, g, T, u, t); break;
    case 9: launch_pair_one<KIND, NT, 9>(ctx, op, g, T, u, t); break;
    case 17: launch_pair_one<KIND, NT, 17>(ctx, op, g, T, u, t); break;
    default: throw Error("no lane-pair kernel for this chunk length");
  }
}

template <int KIND, int NT>
static void launch_strided(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  const int v = ctx.strided_variant;
  if (g.pair) return launch_pair<KIND, NT>(ctx, op, g, T, u, t);
  if (v >= 4 && tile_eligible(op, g, T, u, t)) {
    if (T.L == 8 && T.nc <= 16) return launch_tile_one<KIND, NT, 8, 32, 3, 1>(ctx, op, g, T, u, t);
    if (T.L == 8) return launch_tile_one<KIND, NT, 8, 16, 3, 2>(ctx, op, g, T, u, t);
    if (T.L == 16 && v == 5) return launch_tile_one<KIND, NT, 16, 16, 3, 1>(ctx, op, g, T, u, t);
    if (T.L == 16) return launch_tile_one<KIND, NT, 16, 8, 3, 2>(ctx, op, g, T, u, t);
    if (T.L == 32) return launch_tile_one<KIND, NT, 32, 8, 3, 1>(ctx, op, g, T, u, t);
    if (T.L == 48) return launch_tile_one<KIND, NT, 48, 8, 2, 1>(ctx, op, g, T, u, t);
  }
  if (T.L == 8 && T.nc <= 16) launch_strided_one<KIND, NT, 8, 32, 16, 2>(ctx, op, g, T, u, t);
  else if (T.L == 8 && T.nc <= 32) launch_strided_one<KIND, NT, 8, 32, 32, 1>(ctx, op, g, T, u, t);
  else if (T.L == 16 && T.nc <= 16) launch_strided_one<KIND, NT, 16, 32, 16, 2>(ctx, op, g, T, u, t);
  else if (T.L == 16 && T.nc <= 32) launch_strided_one<KIND, NT, 16, 16, 32, 2>(ctx, op, g, T, u, t);
  else if (T.L == 32 && T.nc <= 16) launch_strided_one<KIND, NT, 32, 16, 16, 2>(ctx, op, g, T, u, t);
  else if (T.L == 32 && T.nc <= 32) launch_strided_one<KIND, NT, 32, 16, 32, 1>(ctx, op, g, T, u, t);
  else if (T.L == 48 && T.nc <= 32) launch_strided_one<KIND, NT, 48, 8, 32, 1>(ctx, op, g, T, u, t);
  else if (T.L == 64 && T.nc <= 32) launch_strided_one<KIND, NT, 64, 8, 32, 1>(ctx, op, g, T, u, t);
  else throw Error("no strided kernel for this line length (n <= 2048 supported)");
}

template <int KIND, int NT, int L, int WPB, int NB, int MINB, bool TMA>
static void launch_contig_one(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t,
                              int chead = 0) {
  const int NPL = T.nc * L;                                    // padded line length
  const int NP = chead > 0 ? (2 * chead + 2) * L : NPL;        // coefficient rows held in shared memory
  int NBUF = (NPL > op.n_in ? NPL : op.n_in) + 2 * HALO;
  NBUF = (NBUF + 1) & ~1;
  const int COEF = (7 * NP + (chead > 0 ? (2 * T.c_head_rs + 2) * L : 0) + 1) & ~1;
  const size_t smem = static_cast<size_t>(COEF + WPB * (NB * NBUF + 8) + WPB * NB) * sizeof(double);
  auto kern = k_contig<KIND, NT, L, WPB, NB, MINB, TMA>;
  static size_t configured_dev[64] = {};   // the attribute is per device: one record per device ordinal
  size_t &configured = configured_dev[ctx.device & 63];
  if (smem > configured) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  long long blocks = (g.nlines + WPB - 1) / WPB;
  const long long cap = static_cast<long long>(ctx.sm_count) * MINB;
  if (blocks > cap) blocks = cap;
  kern<<<static_cast<unsigned>(blocks), 32 * WPB, smem, ctx.stream>>>(op, u, t, chead > 0 ? T.d_rows_c : T.d_rows, T.d_scan, T.nc, g.nlines, NP, NBUF,
                                                                         COEF, chead, chead > 0 ? T.c_head_rs : 0);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

// dynamic shared memory of k_contig<.., L, WPB, NB, ..> for this operator (same formula as launch_contig_one)
template <int L>
static size_t contig_smem(const DevOp &op, const TriTable &T, int wpb, int nb, int chead = 0) {
  const int NPL = T.nc * L;
  const int NP = chead > 0 ? (2 * chead + 2) * L : NPL;
  int NBUF = (NPL > op.n_in ? NPL : op.n_in) + 2 * HALO;
  NBUF = (NBUF + 1) & ~1;
  const int COEF = (7 * NP + (chead > 0 ? (2 * T.c_head_rs + 2) * L : 0) + 1) & ~1;
  return static_cast<size_t>(COEF + wpb * (nb * NBUF + 8) + wpb * nb) * sizeof(double);
}

template <int KIND, int NT, int L>
static void launch_contig_L(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  const bool even = (op.n_in % 2 == 0) && (op.n_out % 2 == 0);
  const bool aligned = (reinterpret_cast<uintptr_t>(u) % 16 == 0) && (reinterpret_cast<uintptr_t>(t) % 16 == 0);
  const int v = (even && aligned) ? ctx.contig_variant : 0;
  if constexpr (L >= 25) {
    // long lines (n > 544): the same two-slot TMA ring as the short ones, so that the load of the next line and the
    // store of the previous one run under the arithmetic; 8 warps when their line buffers fit beside the coefficient
    // table, else 4.  Without it (X3D_CONTIG_VARIANT=0, odd n) a warp loads, solves and stores one after the other.
    constexpr size_t SMEM_MAX = 227 * 1024;
    if (v != 0) {
      if (contig_smem<L>(op, T, 8, 2) <= SMEM_MAX) return launch_contig_one<KIND, NT, L, 8, 2, 1, true>(ctx, op, g, T, u, t);
      // the full coefficient table (7 columns x line length) leaves room for 4 warps only: use the compressed one
      const int chead = ctx.contig_compress ? compress_tri(T) : -1;
      if (chead > 0 && contig_smem<L>(op, T, 8, 2, chead) <= SMEM_MAX)
        return launch_contig_one<KIND, NT, L, 8, 2, 1, true>(ctx, op, g, T, u, t, chead);
      if (chead > 0 && contig_smem<L>(op, T, 4, 2, chead) <= SMEM_MAX)
        return launch_contig_one<KIND, NT, L, 4, 2, 1, true>(ctx, op, g, T, u, t, chead);
      if (contig_smem<L>(op, T, 4, 2) <= SMEM_MAX) return launch_contig_one<KIND, NT, L, 4, 2, 1, true>(ctx, op, g, T, u, t);
    }
  }
  if constexpr (L == 9 || L == 17) {
    if (v == 1) return launch_contig_one<KIND, NT, L, 8, 2, 2, true>(ctx, op, g, T, u, t);
    if (v == 2) return launch_contig_one<KIND, NT, L, 8, 3, 1, true>(ctx, op, g, T, u, t);
    if (v == 3) return launch_contig_one<KIND, NT, L, 4, 3, 3, true>(ctx, op, g, T, u, t);
  }
  if constexpr (L <= 17) launch_contig_one<KIND, NT, L, 8, 1, 2, false>(ctx, op, g, T, u, t);
  else if constexpr (L <= 33) launch_contig_one<KIND, NT, L, 8, 1, 1, false>(ctx, op, g, T, u, t);
  else launch_contig_one<KIND, NT, L, 4, 1, 1, false>(ctx, op, g, T, u, t);  // long lines: 4 warps, up to 255 registers
}

template <int KIND, int NT>
static void launch_contig(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  switch (T.L) {
    case 5: launch_contig_L<KIND, NT, 5>(ctx, op, g, T, u, t); break;
    case 9: launch_contig_L<KIND, NT, 9>(ctx, op, g, T, u, t); break;
    case 17: launch_contig_L<KIND, NT, 17>(ctx, op, g, T, u, t); break;
    case 25: launch_contig_L<KIND, NT, 25>(ctx, op, g, T, u, t); break;
    case 33: launch_contig_L<KIND, NT, 33>(ctx, op, g, T, u, t); break;
    case 49: launch_contig_L<KIND, NT, 49>(ctx, op, g, T, u, t); break;
    case 65: launch_contig_L<KIND, NT, 65>(ctx, op, g, T, u, t); break;
    default: throw Error("no contiguous kernel for this line length (n <= 2080 supported)");
  }
}

template <int KIND, int NT>
static void launch_kind_nt(Ctx &ctx, const DevOp &op, const LineGeom &g, const TriTable &T, const double *u, double *t) {
  if (g.axis == 0) launch_contig<KIND, NT>(ctx, op, g, T, u, t);
  else launch_strided<KIND, NT>(ctx, op, g, T, u, t);
}

}  // namespace x3d
