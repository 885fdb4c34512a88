// x3d_stag.cu -- host side of the fused staggered-operator pairs (x3d_stag_kernels.cuh): eligibility, parameters of
// the cyclic solves, tensor maps, launch.
#include <cmath>
#include "x3d_mom.cuh"
#include "x3d_ops_inst.cuh"
#include "x3d_stag_kernels.cuh"

namespace x3d {

bool mom_pair_plan(int n, int L, MomGeom &g, size_t &smem);

namespace {

// parameters of pair_solve_cyclic_ks for tri(alpha, 1, alpha) on a periodic line of n rows cut into chunks of L;
// scan: host image [10][32] of the Kogge-Stone multipliers
bool make_stag_cyc(double alpha, int n, int L, StagCyc &cy, std::vector<double> &scan) {
  if (!(std::fabs(alpha) < 0.5) || std::fabs(alpha) < 1e-3) return false;
  const int nc = (n + L - 1) / L, rem = n - (nc - 1) * L;
  if (nc > 32 || nc < 2) return false;
  const double rho = (-1.0 + std::sqrt(1.0 - 4.0 * alpha * alpha)) / (2.0 * alpha);
  // five Kogge-Stone levels reach 32 chunks back: what lies further must be below double precision
  if (std::pow(std::fabs(rho), 32.0 * L - (L - rem)) > 1e-19) return false;
  cy.rho = rho;
  cy.esc = std::pow(rho, -(L - rem));
  if (!std::isfinite(cy.esc) || std::fabs(cy.esc) > 1e100) return false;
  const double geo = (1.0 - std::pow(rho, 2 * (L - rem))) / (1.0 - rho * rho);
  cy.gamma = std::pow(rho, rem + 1) * geo;
  cy.delta = rho * geo;
  cy.scale = 1.0 / (-alpha / rho);
  auto len = [&](int c) { c %= nc; if (c < 0) c += nc; return c == nc - 1 ? rem : L; };
  scan.assign(320, 0.0);
  for (int lev = 0; lev < 5; ++lev)
    for (int c = 0; c < nc; ++c) {
      long long rows_f = 0, rows_b = 0;
      for (int j = 0; j < (1 << lev); ++j) { rows_f += len(c - j); rows_b += len(c + j); }
      scan[lev * 32 + c] = std::pow(rho, static_cast<double>(rows_f));
      scan[(5 + lev) * 32 + c] = std::pow(rho, static_cast<double>(rows_b));
    }
  return true;
}

struct ScanKey {
  double alpha;
  int n, L;
  bool operator<(const ScanKey &o) const { return std::tie(alpha, n, L) < std::tie(o.alpha, o.n, o.L); }
};

}  // namespace

struct StagCache {
  std::map<ScanKey, std::pair<StagCyc, double *>> m;
  ~StagCache() { for (auto &kv : m) cudaFree(kv.second.second); }
};
static std::map<Ctx *, std::unique_ptr<StagCache>> g_stag;   // per context (device memory belongs to its device)
void stag_release(Ctx *ctx) { g_stag.erase(ctx); }

static bool get_cyc(Ctx &ctx, double alpha, int n, int L, StagCyc &cy) {
  auto &slot = g_stag[&ctx];
  if (!slot) slot = std::make_unique<StagCache>();
  const ScanKey key{alpha, n, L};
  auto it = slot->m.find(key);
  if (it != slot->m.end()) { cy = it->second.first; return true; }
  std::vector<double> scan;
  if (!make_stag_cyc(alpha, n, L, cy, scan)) return false;
  double *d = nullptr;
  X3D_CUDA(cudaMalloc(&d, scan.size() * sizeof(double)));
  X3D_CUDA(cudaMemcpyAsync(d, scan.data(), scan.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  cy.scan = d;
  slot->m[key] = {cy, d};
  return true;
}

static bool stag_plan(int n, int L, StagGeom &g, size_t &smem) {
  if (L != 17 && L != 9) return false;
  if ((n & 7) || n < 64) return false;
  int nbox, br;
  if (!pair_boxes(n, true, nbox, br)) return false;
  g.nbox = nbox; g.br = br;
  g.n = n;
  g.nc = (n + L - 1) / L;
  if (g.nc > 32 || g.nc < 2) return false;
  const int slot_rows = 8 + n + 8;
  g.slot_bytes = slot_rows * 128;
  const long long overrun = static_cast<long long>(g.nc * L + 8 + HALO - slot_rows) * 128;
  if (overrun > 512 * 8) return false;
  smem = static_cast<size_t>(3) * g.slot_bytes + 512 * 8 + 2 * 3 * 8;
  return smem <= 227 * 1024;
}

// both operators periodic staggered operators on the same lines (n1 lanes, nline rows, nouter slabs; element strides
// 1, sline, souter), fields 16-byte aligned with even strides
bool stag_pair_eligible(Ctx &ctx, const DevOp &opA, const DevOp &opB, long long n1, int nline, long long sline, long long souter, long long nouter) {
  if (!opA.periodic || !opB.periodic || opA.has_post || opB.has_post || opA.rhs_only || opB.rhs_only) return false;
  if (opA.n_in != nline || opA.n_out != nline || opB.n_in != nline || opB.n_out != nline) return false;
  if ((n1 & 1) || (sline & 1) || (nouter > 1 && (souter & 1))) return false;
  const int L = pick_L_contig(nline);
  StagGeom g{};
  size_t smem;
  if (!stag_plan(nline, L, g, smem)) return false;
  StagCyc a{}, b{};
  return get_cyc(ctx, opA.alpha, nline, L, a) && get_cyc(ctx, opB.alpha, nline, L, b);
}

// mode 0: outA = opA(inA) + opB(inB);  mode 1: outA = opA(inA), outB = opB(inA)
void launch_stag_pair(Ctx &ctx, int mode, int axis, const DevOp &opA, const DevOp &opB, const double *inA, const double *inB, double *outA,
                      double *outB, long long n1, int nline, long long nouter, long long sline, long long souter) {
  const int L = pick_L_contig(nline);
  StagGeom g{};
  size_t smem = 0;
  if (!stag_plan(nline, L, g, smem) || !get_cyc(ctx, opA.alpha, nline, L, g.a) || !get_cyc(ctx, opB.alpha, nline, L, g.b))
    throw Error("fused staggered pair: ineligible call");
  const double *ins[2] = {inA, mode == 0 ? inB : inA};
  double *outs[2] = {outA, mode == 1 ? outB : outA};
  for (int q = 0; q < 2; ++q)
    if ((reinterpret_cast<uintptr_t>(ins[q]) | reinterpret_cast<uintptr_t>(outs[q])) & 15u) throw Error("fused staggered pair: unaligned field");
  g.nbx = static_cast<int>((n1 + 15) / 16);
  g.npos = static_cast<long long>(g.nbx) * nouter;
  StagMaps maps;
  maps.inA = make_line_map(ins[0], n1, nline, nouter, sline, souter, 16, g.br, true);
  maps.haloA = make_line_map(ins[0], n1, nline, nouter, sline, souter, 16, 8, true);
  maps.inB = make_line_map(ins[1], n1, nline, nouter, sline, souter, 16, g.br, true);
  maps.haloB = make_line_map(ins[1], n1, nline, nouter, sline, souter, 16, 8, true);
  maps.outA = make_line_map(outs[0], n1, nline, nouter, sline, souter, 16, g.br, true);
  maps.outB = make_line_map(outs[1], n1, nline, nouter, sline, souter, 16, g.br, true);
  auto launch = [&](auto kern) {
    X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    long long blocks = ctx.sm_count;
    if (blocks > g.npos) blocks = g.npos;
    kern<<<static_cast<unsigned>(blocks), MOM_THREADS, smem, ctx.stream>>>(opA, opB, maps, g);
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  };
  static const char *names[2][2] = {{"staggered_sum_y(k_stag)", "staggered_sum_z(k_stag)"}, {"staggered_pair_y(k_stag)", "staggered_pair_z(k_stag)"}};
  ProfScope ps(ctx, names[mode][axis == 2 ? 1 : 0]);
  if (mode == 0) {
    if (opA.kind != IVP || opB.kind != DVP) throw Error("fused staggered pair: mode 0 takes (inter?vp, der?vp)");
    if (L == 17) launch(k_stag<IVP, DVP, 0, 17>); else launch(k_stag<IVP, DVP, 0, 9>);
  } else {
    if (opA.kind != IPV || opB.kind != DPV) throw Error("fused staggered pair: mode 1 takes (inter?pv, der?pv)");
    if (L == 17) launch(k_stag<IPV, DPV, 1, 17>); else launch(k_stag<IPV, DPV, 1, 9>);
  }
}

}  // namespace x3d
