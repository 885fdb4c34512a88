// x3d_mom.cu -- host side of the fused momentum-RHS kernels (x3d_mom_kernels.cuh): compressed
// coefficient tables, tensor maps, eligibility, launch.
#include <cmath>
#include "x3d_mom.cuh"
#include "x3d_mom_kernels.cuh"
#include "x3d_slab_kernels.cuh"
#include "x3d_ops_inst.cuh"

namespace x3d {

MomTable::~MomTable() {
  if (d_c) cudaFree(d_c);
  if (d_scan) cudaFree(d_scan);
}

// Compress the per-row table of a periodic operator into MOM_TABS chunk tables.  Valid when every chunk
// 3 .. nc-4 has the rows of chunk 3 (the LU recurrence of prepare() has reached its fixed point) and a
// Sherman-Morrison vector that is zero at double precision relative to its boundary values.
bool build_mom_table(Ctx &ctx, const TriTable &T, MomTable &M) {
  const int L = T.L, nc = T.nc;
  const int H = mom_head(L), MOM_TABS = mom_tabs(L);
  M.ok = false;
  if (nc < MOM_TABS || nc > 32 || T.h_rows.empty()) return false;
  auto R = [&](int row, int col) { return T.h_rows[static_cast<size_t>(row) * TRI_W + col]; };
  double rsmax = 0.0;
  for (int i = 0; i < nc * L; ++i) rsmax = std::max(rsmax, std::fabs(R(i, T_RS)));
  const int cols[5] = {T_S, T_PF, T_W, T_FW, T_PB};
  for (int c = H + 1; c <= nc - H - 2; ++c)
    for (int m = 0; m < L; ++m) {
      for (int q = 0; q < 5; ++q) {
        const double a = R(c * L + m, cols[q]), b = R(H * L + m, cols[q]);
        if (std::fabs(a - b) > 4e-16 * std::max(std::fabs(a), std::fabs(b))) return false;
      }
    }
  for (int c = H; c <= nc - H - 2; ++c)
    for (int m = 0; m < L; ++m)
      if (std::fabs(R(c * L + m, T_RS)) > 1e-18 * rsmax) return false;
  // constant super-diagonal: fw(i) = ff w(i) for every row but the last
  const double ff = R(1, T_FW) / R(1, T_W);
  for (int i = 0; i < T.n - 1; ++i)
    if (std::fabs(R(i, T_FW) - ff * R(i, T_W)) > 4e-16 * std::fabs(R(i, T_FW))) return false;
  M.ff = ff;
  std::vector<double> h(static_cast<size_t>(3) * MOM_TABS * L * 2, 0.0);
  for (int t = 0; t < MOM_TABS; ++t) {
    const int c = t < H ? t : (t == H ? H : nc - H - 1 + (t - H - 1));
    for (int m = 0; m < L; ++m) {
      const int row = c * L + m;
      for (int pr = 0; pr < 3; ++pr) {
        const size_t o = ((static_cast<size_t>(pr) * MOM_TABS + t) * L + m) * 2;
        h[o] = R(row, 2 * pr);
        h[o + 1] = (t == H && pr == 2) ? 0.0 : R(row, 2 * pr + 1);  // generic chunk: rs = 0
      }
    }
  }
  X3D_CUDA(cudaMalloc(&M.d_c, h.size() * sizeof(double)));
  X3D_CUDA(cudaMalloc(&M.d_scan, T.h_scan.size() * sizeof(double)));
  X3D_CUDA(cudaMemcpyAsync(M.d_c, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(M.d_scan, T.h_scan.data(), T.h_scan.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  M.L = L; M.nc = nc;
  M.ok = true;
  return true;
}

// parameters of pair_solve_cyclic for the periodic system tri(alpha, 1, alpha); scale_num / c multiplies the solution
static bool make_cyc(double alpha, int n, int L, double scale_num, MomGeom::Cyc &cy, bool open = false) {
  if (!(std::fabs(alpha) < 0.5) || std::fabs(alpha) < 1e-3) return false;
  const int nc = (n + L - 1) / L, rem = n - (nc - 1) * L;
  const double rho = (-1.0 + std::sqrt(1.0 - 4.0 * alpha * alpha)) / (2.0 * alpha);
  const double c = -alpha / rho;
  int K = 1;
  while (std::pow(std::fabs(rho), static_cast<double>(K) * L) >= 1e-18 && K < 16) ++K;
  if (rem != L) ++K;     // one of the K chunks may be the short last one
  if (K > 8 || (!open && K > nc - 1)) return false;   // open (slab) lines: the look-back stops at the face
  cy.rho = rho;
  cy.rhoL = std::pow(rho, L);
  cy.rhoR = std::pow(rho, rem);
  cy.esc = std::pow(rho, -(L - rem));
  if (!std::isfinite(cy.esc) || std::fabs(cy.esc) > 1e100) return false;
  const double geo = (1.0 - std::pow(rho, 2 * (L - rem))) / (1.0 - rho * rho);
  cy.gamma = std::pow(rho, rem + 1) * geo;
  cy.delta = rho * geo;
  cy.scale = scale_num / c;
  cy.K = K;
  return true;
}
bool mom_cyclic_ok(double alpha, int n, int L) {
  MomGeom::Cyc cy{};
  return make_cyc(alpha, n, L, 1.0, cy);
}

// Overlap-save segments for a periodic line of ntot rows (MomGeom::ntot ...): tile rows T = S + 2 H <= 544.
// nseg = 1 when the whole line fits a tile.  False when no segmentation works (line extents, decay of the two operators).
bool mom_segments(int ntot, double alpha1, double alpha2, int &S, int &H, int &nseg) {
  if (ntot <= 544) { S = ntot; H = 0; nseg = 1; return true; }
  double rmax = 0.0;
  for (double a : {alpha1, alpha2}) {
    if (!(std::fabs(a) < 0.5) || std::fabs(a) < 1e-3) return false;
    rmax = std::max(rmax, std::fabs((-1.0 + std::sqrt(1.0 - 4.0 * a * a)) / (2.0 * a)));
  }
  for (int h : {48, 64, 96, 128}) {
    if (ntot % h) continue;
    if (std::pow(rmax, h - 4) > 1e-17) continue;   // the 4 stencil rows next to the artificial wrap carry O(1) errors
    for (int s = ((544 - 2 * h) / h) * h; s >= h; s -= h)
      if (ntot % s == 0) { S = s; H = h; nseg = ntot / s; return true; }
  }
  return false;
}

bool mom_pair_plan(int n, int L, MomGeom &g, size_t &smem) {
  if (L != 17 && L != 9) return false;
  if ((n & 7) || n < 64) return false;
  int nbox, br;
  if (!pair_boxes(n, true, nbox, br)) return false;
  g.nbox = nbox; g.br = br;
  g.n = n;
  g.nc = (n + L - 1) / L;
  if (g.nc > 32 || g.nc < mom_tabs(L)) return false;
  const int slot_rows = 8 + n + 8;
  g.slot_bytes = slot_rows * 128;
  const size_t tail = static_cast<size_t>(2) * 3 * mom_tabs(L) * L * 16 + 2 * 320 * 8 + 2 * 3 * 8;
  const long long overrun = static_cast<long long>(g.nc * L + 8 + HALO - slot_rows) * 128;
  if (overrun > static_cast<long long>(tail)) return false;
  smem = static_cast<size_t>(3) * g.slot_bytes + tail;
  return smem <= 227 * 1024;
}

// x lines: 16 lines of pitch n+8 doubles per slot
bool mom_x_plan(int n, int L, MomGeom &g, size_t &smem) {
  if (L != 17 && L != 9) return false;
  if ((n & 1) || n < 64) return false;
  g.n = n;
  g.nc = (n + L - 1) / L;
  if (g.nc > 32 || g.nc < mom_tabs(L)) return false;
  g.pitch = n + 2 * HALO;
  g.slot_bytes = ((16 * g.pitch * 8 + 1023) / 1024) * 1024;
  const size_t tail = static_cast<size_t>(2) * 3 * mom_tabs(L) * L * 16 + 2 * 320 * 8 + 2 * 3 * 8;
  // the last chunk's window may read up to nc*L + HALO elements of the last line of a slot
  const long long overrun = (15LL * g.pitch + g.nc * L + 2 * HALO) * 8 - g.slot_bytes;
  if (overrun > static_cast<long long>(tail)) return false;
  smem = static_cast<size_t>(3) * g.slot_bytes + tail;
  return smem <= 227 * 1024;
}
bool mom_x_eligible(int n, int L) {
  MomGeom g{};
  size_t smem;
  return mom_x_plan(n, L, g, smem);
}

bool mom_pair_eligible(int n, int L) {
  MomGeom g{};
  size_t smem;
  return mom_pair_plan(n, L, g, smem);
}

// fields: f[0..2] = ux, uy, uz pencils with the layout described by (n1, nline, nouter, sline, souter);
// out[0..2] = xnu D2(c) - 1/2 (D1(c a) + a D1(c)) along this axis, a = f[axis]
void launch_mom_pair(Ctx &ctx, int axis, const DevOp &op1, const DevOp &op2, const MomTable &M1, const MomTable &M2, double xnu,
                     const double *const f[3], double *const out[3], long long n1, int nline, long long nouter, long long sline,
                     long long souter, bool add, bool cyclic) {
  MomGeom g{};
  size_t smem = 0;
  int segS, segH, nseg;
  if (!mom_segments(nline, op1.alpha, op2.alpha, segS, segH, nseg) || (nseg > 1 && !cyclic)) throw Error("fused momentum kernel: ineligible call");
  const int T = segS + 2 * segH;                  // rows of a tile (= nline when the line is not segmented)
  const int L = pick_L_contig(T);
  if (!cyclic && (!M1.ok || !M2.ok || M1.L != M2.L || M1.L != L)) throw Error("fused momentum kernel: ineligible call");
  if (!mom_pair_plan(T, L, g, smem)) throw Error("fused momentum kernel: ineligible call");
  g.ntot = nline; g.seg_S = segS; g.seg_H = segH; g.nseg = nseg;
  if (nseg > 1) { g.br = segH; g.nbox = T / segH; }   // boxes of seg_H rows never straddle the end of the line
  g.rem = T - (g.nc - 1) * L;
  if (cyclic && !(make_cyc(op1.alpha, T, L, -0.5, g.cy1) && make_cyc(op2.alpha, T, L, xnu, g.cy2)))
    throw Error("fused momentum kernel: cyclic solves are not possible for this scheme");
  g.nbx = static_cast<int>((n1 + 15) / 16);
  g.npos = static_cast<long long>(g.nbx) * nouter * nseg;
  g.ia = axis; g.ic1 = (axis + 1) % 3; g.ic2 = (axis + 2) % 3;
  g.xnu = xnu;
  g.add = add ? 1 : 0;
  MomMaps maps;
  for (int q = 0; q < 3; ++q) {
    maps.in[q] = make_line_map(f[q], n1, nline, nouter, sline, souter, 16, g.br, true);
    maps.halo[q] = make_line_map(f[q], n1, nline, nouter, sline, souter, 16, 8, true);
    maps.out[q] = make_line_map(out[q], n1, nline, nouter, sline, souter, 16, g.br, true);
  }
  MomTabs tb{reinterpret_cast<const double2 *>(M1.d_c), reinterpret_cast<const double2 *>(M2.d_c), M1.d_scan, M2.d_scan, M1.ff, M2.ff};
  const bool nt4 = op2.c[2] != 0.0 || op2.c[3] != 0.0;
  auto launch = [&](auto kern) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    long long blocks = ctx.sm_count;
    if (blocks > g.npos) blocks = g.npos;
    kern<<<static_cast<unsigned>(blocks), MOM_THREADS, smem, ctx.stream>>>(op1, op2, maps, tb, g);
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  };
  ProfScope ps(ctx, axis == 1 ? "momentum_fused_y(k_mom_pair)" : "momentum_fused_z(k_mom_pair)");
  if (cyclic) {
    if (L == 17) { if (nt4) launch(k_mom_pair<17, 4, false, true, false>); else launch(k_mom_pair<17, 2, false, true, false>); }
    else { if (nt4) launch(k_mom_pair<9, 4, false, true, false>); else launch(k_mom_pair<9, 2, false, true, false>); }
  } else {
    if (L == 17) { if (nt4) launch(k_mom_pair<17, 4, false, false, false>); else launch(k_mom_pair<17, 2, false, false, false>); }
    else { if (nt4) launch(k_mom_pair<9, 4, false, false, false>); else launch(k_mom_pair<9, 2, false, false, false>); }
  }
}

// x lines: fields are (n, nlines) arrays with contiguous lines
void launch_mom_x(Ctx &ctx, const DevOp &op1, const DevOp &op2, const MomTable &M1, const MomTable &M2, double xnu,
                  const double *const f[3], double *const out[3], int n, long long nlines, bool add, bool cyclic, const MomIntt *intt) {
  MomGeom g{};
  size_t smem = 0;
  int segS, segH, nseg;
  if (!mom_segments(n, op1.alpha, op2.alpha, segS, segH, nseg) || (nseg > 1 && !cyclic)) throw Error("fused x momentum kernel: ineligible call");
  const int T = segS + 2 * segH;
  const int L = pick_L_contig(T);
  if (!cyclic && (!M1.ok || !M2.ok || M1.L != M2.L || M1.L != L)) throw Error("fused x momentum kernel: ineligible call");
  if (!mom_x_plan(T, L, g, smem)) throw Error("fused x momentum kernel: ineligible call");
  g.ntot = n; g.seg_S = segS; g.seg_H = segH; g.nseg = nseg;
  g.rem = T - (g.nc - 1) * L;
  if (cyclic && !(make_cyc(op1.alpha, T, L, -0.5, g.cy1) && make_cyc(op2.alpha, T, L, xnu, g.cy2)))
    throw Error("fused x momentum kernel: cyclic solves are not possible for this scheme");
  if (intt && nseg > 1 && intt->u_out[0] == intt->u[0]) throw Error("fused x momentum kernel: a segmented line needs a separate output velocity");
  if (intt && !cyclic) throw Error("fused x momentum kernel: the folded time integration needs the cyclic solves");
  for (int q = 0; q < 3; ++q) {
    if ((reinterpret_cast<uintptr_t>(f[q]) | (intt ? 0 : reinterpret_cast<uintptr_t>(out[q]))) & 15u) throw Error("fused x momentum kernel: unaligned field");
    g.fin[q] = f[q]; g.fout[q] = intt ? nullptr : out[q];
  }
  if (intt) {
    for (int q = 0; q < 3; ++q) {
      g.isum[q] = intt->sum[q]; g.iextra[q] = intt->has_extra ? intt->extra[q] : nullptr; g.iold_in[q] = intt->use_old ? intt->old_in[q] : nullptr;
      g.iu[q] = intt->u[q]; g.iu_out[q] = intt->u_out[q]; g.iold_out[q] = intt->store_old ? intt->old_out[q] : nullptr;
      const uintptr_t all = reinterpret_cast<uintptr_t>(g.isum[q]) | reinterpret_cast<uintptr_t>(g.iextra[q]) | reinterpret_cast<uintptr_t>(g.iold_in[q]) |
                            reinterpret_cast<uintptr_t>(g.iu[q]) | reinterpret_cast<uintptr_t>(g.iu_out[q]) | reinterpret_cast<uintptr_t>(g.iold_out[q]);
      if (all & 15u) throw Error("fused x momentum kernel: unaligned field");
    }
    g.ca = intt->ca; g.cb = intt->cb;
    g.use_old = intt->use_old; g.store_old = intt->store_old; g.has_extra = intt->has_extra;
  }
  g.nlines = nlines;
  g.npos = ((nlines + 15) / 16) * nseg;
  g.nbx = 1;
  g.ia = 0; g.ic1 = 1; g.ic2 = 2;
  g.xnu = xnu;
  g.add = add ? 1 : 0;
  MomMaps maps{};
  MomTabs tb{reinterpret_cast<const double2 *>(M1.d_c), reinterpret_cast<const double2 *>(M2.d_c), M1.d_scan, M2.d_scan, M1.ff, M2.ff};
  const bool nt4 = op2.c[2] != 0.0 || op2.c[3] != 0.0;
  auto launch = [&](auto kern) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    long long blocks = ctx.sm_count;
    if (blocks > g.npos) blocks = g.npos;
    kern<<<static_cast<unsigned>(blocks), MOM_THREADS, smem, ctx.stream>>>(op1, op2, maps, tb, g);
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  };
  ProfScope ps(ctx, intt ? "momentum_fused_x+intt(k_mom_pair)" : "momentum_fused_x(k_mom_pair)");
  if (intt) {
    if (L == 17) { if (nt4) launch(k_mom_pair<17, 4, true, true, true>); else launch(k_mom_pair<17, 2, true, true, true>); }
    else { if (nt4) launch(k_mom_pair<9, 4, true, true, true>); else launch(k_mom_pair<9, 2, true, true, true>); }
  } else if (cyclic) {
    if (L == 17) { if (nt4) launch(k_mom_pair<17, 4, true, true, false>); else launch(k_mom_pair<17, 2, true, true, false>); }
    else { if (nt4) launch(k_mom_pair<9, 4, true, true, false>); else launch(k_mom_pair<9, 2, true, true, false>); }
  } else {
    if (L == 17) { if (nt4) launch(k_mom_pair<17, 4, true, false, false>); else launch(k_mom_pair<17, 2, true, false, false>); }
    else { if (nt4) launch(k_mom_pair<9, 4, true, false, false>); else launch(k_mom_pair<9, 2, true, false, false>); }
  }
}


// ---- z part of the momentum right-hand side on slabs, without transposes (x3d_slab_kernels.cuh) -----------------------
namespace {
double root_of(double alpha) { return (-1.0 + std::sqrt(1.0 - 4.0 * alpha * alpha)) / (2.0 * alpha); }

// chunks of 17 rows from 128 rows on (8 .. 17 chunks per line: 2 or 4 lines per warp), of 9 rows below (64 rows: 8 chunks)
int slab_L(int n) { return n >= 128 ? 17 : 9; }

bool slab_plan(long long nlanes, int n, SlabGeom &g, size_t &smem) {
  const int L = slab_L(n);
  if ((n & 7) || n < 64 || n > 544 || (nlanes & 1)) return false;
  if (!pair_boxes(n, true, g.nbox, g.br)) return false;
  g.n = n;
  g.nc = (n + L - 1) / L;
  g.rem = n - (g.nc - 1) * L;
  g.nbx = static_cast<int>((nlanes + 15) / 16);
  g.ntiles = g.nbx;
  g.gshift = g.nc <= 8 ? 2 : (g.nc <= 16 ? 1 : 0);
  while (g.gshift > 0 && (g.ntiles % (1 << g.gshift)) != 0) --g.gshift;
  g.npos = g.ntiles >> g.gshift;
  g.sub_bytes = (8 + n + 8) * 128;
  g.slot_bytes = g.sub_bytes << g.gshift;
  if (static_cast<long long>(g.nc * L + 8 + HALO - (n + 16)) * 128 > 512 * 8) return false;
  smem = static_cast<size_t>(3) * g.slot_bytes + 512 * 8 + 2 * 3 * 8;
  return smem <= 227 * 1024;
}
}  // namespace

// rows of a slab line: both recurrences must forget a whole slab (|rho|^n < 1e-17), see x3d_slab_kernels.cuh
bool mom_slab_eligible(double alpha1, double alpha2, long long nlanes, int n) {
  SlabGeom g{};
  size_t smem;
  if (!slab_plan(nlanes, n, g, smem)) return false;
  MomGeom::Cyc cy{};
  for (double a : {alpha1, alpha2}) {
    if (!make_cyc(a, n, slab_L(n), 1.0, cy, true)) return false;
    if (std::pow(std::fabs(root_of(a)), n) > 1e-17) return false;
  }
  return true;
}

void launch_mom_slab(Ctx &ctx, const DevOp &op1, const DevOp &op2, double xnu, const double *const f[3], const double *const halo[3],
                     double *const out[3], long long nlanes, int n, bool add, double *carry) {
  SlabGeom g{};
  size_t smem = 0;
  const int L = slab_L(n);
  if (!slab_plan(nlanes, n, g, smem) || !make_cyc(op1.alpha, n, L, -0.5, g.cy1, true) || !make_cyc(op2.alpha, n, L, xnu, g.cy2, true))
    throw Error("slab momentum kernel: ineligible call");
  g.ia = 2; g.ic1 = 0; g.ic2 = 1;
  g.add = add ? 1 : 0;
  g.carry = carry;
  g.nlanes = nlanes;
  MomMaps maps;
  for (int q = 0; q < 3; ++q) {
    if ((reinterpret_cast<uintptr_t>(f[q]) | reinterpret_cast<uintptr_t>(halo[q]) | reinterpret_cast<uintptr_t>(out[q])) & 15u)
      throw Error("slab momentum kernel: unaligned field");
    maps.in[q] = make_line_map(f[q], nlanes, n, 1, nlanes, nlanes * n, 16, g.br, true);
    maps.halo[q] = make_line_map(halo[q], nlanes, 16, 1, nlanes, nlanes * 16, 16, 8, true);
    maps.out[q] = make_line_map(out[q], nlanes, n, 1, nlanes, nlanes * n, 16, g.br, true);
  }
  const bool nt4 = op2.c[2] != 0.0 || op2.c[3] != 0.0;
  auto launch = [&](auto kern) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    long long blocks = ctx.sm_count;
    if (blocks > g.npos) blocks = g.npos;
    kern<<<static_cast<unsigned>(blocks), MOM_THREADS, smem, ctx.stream>>>(op1, op2, maps, g);
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  };
  ProfScope ps(ctx, "momentum_fused_z_slab(k_mom_slab)");
  if (L == 17) { if (nt4) launch(k_mom_slab<17, 4>); else launch(k_mom_slab<17, 2>); }
  else { if (nt4) launch(k_mom_slab<9, 4>); else launch(k_mom_slab<9, 2>); }
}

// tables A(i), B(i) of the face corrections for the two operators, and the band width W
void build_zfix(Ctx &ctx, const DevOp &op1, const DevOp &op2, double xnu, int n, ZFix &Z) {
  const double r1 = root_of(op1.alpha), r2 = root_of(op2.alpha);
  std::vector<double> h(static_cast<size_t>(4) * n);
  for (int i = 0; i < n; ++i) {
    h[i] = std::pow(r1, i + 1) * (1.0 - std::pow(r1, 2.0 * (n - i))) / (1.0 - r1 * r1);
    h[n + i] = std::pow(r1, n - i);
    h[2 * n + i] = std::pow(r2, i + 1) * (1.0 - std::pow(r2, 2.0 * (n - i))) / (1.0 - r2 * r2);
    h[3 * n + i] = std::pow(r2, n - i);
  }
  const double rmax = std::max(std::fabs(r1), std::fabs(r2));
  int W = 1;
  while (W < n && std::pow(rmax, W) > 1e-19) ++W;
  Z.tab.reserve(h.size() * sizeof(double));
  X3D_CUDA(cudaMemcpyAsync(Z.tab.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  Z.n = n; Z.W = W;
  Z.k1 = -0.5 / (-op1.alpha / r1);
  Z.k2 = xnu / (-op2.alpha / r2);
}

void launch_zfix(Ctx &ctx, const ZFix &Z, const double *yin, const double *z0n, const double *yout, double *const sum[3], const double *a,
                 long long nlanes) {
  ZFixArgs z{};
  z.yin = yin; z.z0n = z0n; z.yout = yout;
  for (int q = 0; q < 3; ++q) z.sum[q] = sum[q];
  z.a = a;
  z.tab = static_cast<const double *>(Z.tab.p);
  z.nl = nlanes; z.n = Z.n; z.W = Z.W; z.k1 = Z.k1; z.k2 = Z.k2;
  ProfScope ps(ctx, "momentum_z_face_corrections(k_zfix)");
  if ((nlanes & 1) || ((reinterpret_cast<uintptr_t>(yin) | reinterpret_cast<uintptr_t>(z0n) | reinterpret_cast<uintptr_t>(yout) | reinterpret_cast<uintptr_t>(a) |
                       reinterpret_cast<uintptr_t>(sum[0]) | reinterpret_cast<uintptr_t>(sum[1]) | reinterpret_cast<uintptr_t>(sum[2])) & 15u))
    throw Error("slab face corrections: unaligned field");
  k_zfix<<<static_cast<unsigned>((nlanes / 2 + 255) / 256), 256, 0, ctx.stream>>>(z);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

}  // namespace x3d
