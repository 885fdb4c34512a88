// x3d_solver.cu -- device-resident time step: the reference's
// momentum_rhs_eq (src/transeq.f90:73-591), intt (src/time_integrators.f90:18-187), pre_correc /
// divergence / gradp / cor_vel (src/navier.f90:502,257,386,206), init_tgv and postprocess_tgv
// (src/Case-TGV.f90:25,189) chained on one GPU.  Fields never leave HBM; on a single rank every
// pencil transpose of the reference is the identity and is not executed.
#include <cmath>
#include <cstdlib>
#include "x3d_mom.cuh"
#include "x3d_schemes.cuh"
#include "x3d_state.cuh"

namespace x3d {

void poisson_init(Ctx &ctx, const x3d_poisson_params &p);
void poisson_solve_device(Ctx &ctx, double *d_rhs);
void decomp_init(Ctx &ctx, int nx, int ny, int nz, int p_row, int p_col, int rank, int nranks, const void *nccl_id);
int decomp_info_init(Ctx &ctx, int nx, int ny, int nz);
void decomp_info_get(Ctx &ctx, int id, x3d_decomp_info *out);
void transpose_device(Ctx &ctx, int which, const double *d_src, double *d_dst, int id, int elem);
void transpose_device_multi(Ctx &ctx, int which, int nf, const double *const *d_src, double *const *d_dst, int id, int elem);
void allreduce(Ctx &ctx, double *d_buf, int n, bool is_max);
void decomp_shape(Ctx &ctx, int *p_row, int *p_col, int *rank, int *nranks);
bool ring_exchange(Ctx &ctx, int n, const RingCopy *items);
bool ring_available(Ctx &ctx);

namespace {

// max / min of a strided plane (outflow celerity, Case-Cylinder-wake.f90:146-157): out[0] = max, out[1] = -min
__global__ void k_plane_minmax(const double *__restrict__ a, long long n, long long stride, double *__restrict__ out) {
  double mx = -1609.0, mn = 1609.0;
  for (long long q = threadIdx.x; q < n; q += blockDim.x) { const double v = a[q * stride]; mx = fmax(mx, v); mn = fmin(mn, v); }
  __shared__ double s1[32], s2[32];
  for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o)); mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o)); }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = mx; s2[threadIdx.x >> 5] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) { mx = fmax(mx, s1[w]); mn = fmin(mn, s2[w]); }
    out[0] = mx; out[1] = -mn;
  }
}

template <class F>
__global__ void k_map(long long n, F f) {
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n;
       q += static_cast<long long>(gridDim.x) * blockDim.x)
    f(q);
}

// block-level sum / max reduction into per-block partials, then one final block
template <int NV, class F>
__global__ void k_reduce_partial(long long n, F f, double *__restrict__ partial) {
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n;
       q += static_cast<long long>(gridDim.x) * blockDim.x)
    f(q, acc);
  __shared__ double sh[NV][32];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double x = acc[v];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[v][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double x = threadIdx.x < (blockDim.x >> 5) ? sh[v][threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (threadIdx.x == 0) partial[static_cast<long long>(v) * gridDim.x + blockIdx.x] = x;
    }
  }
}
template <int NV>
__global__ void k_reduce_final(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
  __shared__ double sh[NV][32];
  for (int v = 0; v < NV; ++v) {
    double x = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) x += partial[static_cast<long long>(v) * nblocks + b];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) sh[v][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int v = 0; v < NV; ++v) {
      double x = threadIdx.x < (blockDim.x >> 5) ? sh[v][threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (threadIdx.x == 0) out[v] = x;
    }
  }
}
__global__ void k_max_partial(long long n, const double *__restrict__ a, double *__restrict__ partial) {
  double m = -1609.0;  // navier.f90:349
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < n;
       q += static_cast<long long>(gridDim.x) * blockDim.x)
    m = fmax(m, a[q]);
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    double x = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : -1609.0;
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, o));
    if (threadIdx.x == 0) partial[blockIdx.x] = x;
  }
}
__global__ void k_max_final(int nblocks, const double *__restrict__ partial, double *__restrict__ out) {
  double m = -1609.0;
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) m = fmax(m, partial[b]);
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = sh[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) x = fmax(x, sh[w]);
    out[0] = x;
  }
}

}  // namespace

struct PreOp {
  DevOp op{};
  OpCall call{};
  bool ready = false;
};

struct SolverImpl : SolverState {
  x3d_solver_params p{};
  AxisCoeffs A[3];
  int nxm = 0, nym = 0, nzm = 0;
  double xnu = 0, adt[3]{}, bdt[3]{}, cdt[3]{}, gdt[3]{};
  int iadvance = 1, ntime = 1;
  long long itime = 0;
  size_t n = 0, n3 = 0;       // local x-pencil points, local pressure z-pencil points
  // decomposition (slab: p_row = 1, p_col = nranks)
  int nranks = 1, rank = 0;
  int id_v = 0, id_p3 = 0, id_p1 = 0;
  int nzl = 0, z0 = 0;        // local z extent / 0-based offset of the x- and y-pencils
  int nyl = 0, y0 = 0;        // local y extent / offset of the velocity z-pencil
  int nyml = 0;               // local y extent of the pressure z-pencils
  // fields
  DevBuf ux, uy, uz, px, py, pz, pp3, dux[3], duy[3], duz[3];
  DevBuf w[16];
  DevBuf red_partial, red_out;
  double *h_red = nullptr;  // pinned
  // prepared operators
  PreOp d1[3][2], d2[3][2];                 // [axis][npaire]
  PreOp dvp[3], ivp[3], dpv[3], ipv[3];
  // stretched y mesh and wall (Dirichlet) data
  StretchY st;
  DevBuf d_pp;                // device: pp2y[ny] | pp4y[ny] | ppy[ny]
  DevBuf dpd;                 // wall pressure gradients kept between gradp and pre_correc (navier.f90:439-496):
                              // x faces 4 x (ny,nzl) | y faces 4 x (nx,nzl) | z faces 4 x (nx,ny)
  // operators whose store is an accumulation (DevOp::store_mode): t += op(u) for the sums of divergence
  // (navier.f90:325,339) and t -= op(u) for cor_vel folded into the last pressure-gradient operators (:242-244,426-430)
  PreOp ivp_y_add, dvp_z_add, dpv_x_sub, ipv_x_sub;
  bool fuse_sums = true;      // X3D_FUSE_SUMS=0 restores the separate elementwise passes
  // apply_spatial_filter (src/tools.f90:600-675): filter operators per axis for the current filter parameter
  PreOp fil[3][2];            // [axis][npaire]
  LU3 fil_lu[3][2];
  double fil_af = -1.0;
  // case glue: channel forcing, cylinder inflow / outflow, immersed boundary (x3d_solver_set_case)
  x3d_case_params cs{};
  double fcpg = 0.0;
  DevBuf bwx;                 // wall velocities of the x faces: bxx1 bxy1 bxz1 bxxn bxyn bxzn, (ny, nzl) each
  DevBuf bnoise;              // bxo byo bzo (ny, nzl)
  DevBuf ep1;                 // (nx, ny, nzl)
  // fused momentum kernels (periodic directions): compressed tables of D1 / D2 per axis
  MomTable mt1[3], mt2[3];
  bool fused[3] = {false, false, false};
  bool cyclic[3] = {false, false, false};   // table-free cyclic solves in the fused kernels (X3D_MOM_CYC=0: tables)
  bool segmented[3] = {false, false, false};   // long lines (n > 544) run the fused kernels by overlap-save segments
  DevBuf unew[3];                           // segmented x lines: the fused time integration writes the new velocity here
  bool stag[3] = {false, false, false};     // fused pairs of staggered operators on periodic y / z lines (X3D_FUSE_STAG=0: off)
  bool fuse_intt = true;                    // time integration folded into the x momentum kernel (X3D_FUSE_INTT=0: k_map pass)
  // z part of the momentum terms on the slabs themselves (x3d_slab_kernels.cuh): no y <-> z transposes of the velocity and of
  // its right-hand side; neighbouring ranks exchange 4 halo planes per component and 18 carry planes.  X3D_SLABZ=0: transposes.
  // slab_nv > 1 on ONE rank (X3D_SLABZ_EMULATE=P): the rank's planes are treated as P virtual slabs with local copies for
  // the exchanges -- the same kernels and arithmetic as P ranks, testable on one GPU.
  // several ranks: the two fields that divergence / gradp transpose are ready at different times; the transpose of the first one
  // runs on `aux` beside the kernels that produce (divergence) or consume (gradp) the other one.  X3D_OVERLAP_DIV=0: one scope.
  bool overlap_div = true;
  bool slabz = false;
  int slab_nv = 1;
  DevBuf zhalo[3];                          // [slab_nv][16 planes]
  DevBuf zcarry_out, zcarry_in;             // [slab_nv][18 planes]: Yout | Z0 of this rank;  Yout of the previous | Z0 of the next rank
  ZFix zfix;
  // several ranks, X3D_OVERLAP=1: the y -> z transposes of the velocity run on `aux` while the x and y momentum
  // kernels compute.  Off by default: on 2 B200 it gave 42.4 ms per 512^3 step against 42.3 ms on one stream (the
  // hidden copies are paid back by slower kernels beside them and by the extra array intt then reads).
  int overlap = 0;   // 0 off | 1: y and x kernels beside the forward transposes | 2: the y kernel beside them, then z, then x (+ intt)
  // x3d_solver_advance_host: three velocity sets in rotation (the current one and two spares) so that the H2D copy of
  // job j+1 and the D2H copy of job j-1 run on their own streams beside the kernels of job j
  struct VelSet {
    DevBuf b[3];
    cudaEvent_t h2d_done = nullptr, d2h_done = nullptr;
    bool pending = false;           // a D2H copy out of this set has been queued (d2h_done is valid)
    const void *host_out = nullptr; // its destination (a later job reading that host array must wait for it)
  };
  VelSet spare[2];
  cudaEvent_t cur_d2h = nullptr, ev_step = nullptr;
  bool cur_pending = false;
  const void *cur_host_out = nullptr;
  int next_spare = 0;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  ~SolverImpl() override {
    if (h_red) cudaFreeHost(h_red);
    if (aux) cudaStreamDestroy(aux);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    for (auto &v : spare) { if (v.h2d_done) cudaEventDestroy(v.h2d_done); if (v.d2h_done) cudaEventDestroy(v.d2h_done); }
    if (cur_d2h) cudaEventDestroy(cur_d2h);
    if (ev_step) cudaEventDestroy(ev_step);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
};

namespace {

double *B(DevBuf &b) { return static_cast<double *>(b.p); }

void prep(Ctx &ctx, PreOp &P, Kind kind, int axis, const AxisCoeffs &A, int npaire, const LU3 &lu, const int dims_in[3],
          const double *post = nullptr) {
  OpCall &c = P.call;
  c = OpCall{};
  c.kind = kind; c.axis = axis; c.ncl1 = A.ncl1; c.ncln = A.ncln; c.periodic = A.periodic; c.npaire = npaire;
  c.n = A.n; c.nm = A.nm;
  for (int d = 0; d < 3; ++d) c.dims_in[d] = dims_in[d];
  c.f = lu.f.data(); c.s = lu.s.data(); c.w = lu.w.data();
  c.post = post; c.rhs_only = false;
  build_devop(ctx, c, P.op);
  if (P.op.untouched) throw Error("solver: operator variant not implemented by the reference");
  P.ready = true;
}
void run(Ctx &ctx, const PreOp &P, const double *u, double *t) { launch_line_op(ctx, P.op, P.call, u, t); }

int gridn(const Ctx &ctx, long long n) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(ctx.sm_count) * 16;
  return static_cast<int>(b < cap ? b : cap);
}
template <class F>
void map(Ctx &ctx, long long n, F f) {
  ProfScope ps(ctx, "elementwise(k_map)");
  k_map<<<gridn(ctx, n), 256, 0, ctx.stream>>>(n, f);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

}  // namespace

void solver_init(Ctx &ctx, const x3d_solver_params &p) {
  X3D_CUDA(cudaSetDevice(ctx.device));
  auto S = std::make_unique<SolverImpl>();
  S->p = p;
  // decomposition: single rank unless x3d_decomp_init was called before with several ranks
  if (!ctx.decomp) decomp_init(ctx, p.nx, p.ny, p.nz, 1, 1, 0, 1, nullptr);
  int pr, pc;
  decomp_shape(ctx, &pr, &pc, &S->rank, &S->nranks);
  if (S->nranks > 1 && pr != 1) throw Error("x3d_solver_init: the distributed solver uses slabs (p_row = 1, p_col = number of GPUs)");
  SchemeOpts o;
  o.ifirstder = p.ifirstder; o.isecondder = p.isecondder; o.ipinter = p.ipinter; o.nu0nu = p.nu0nu; o.cnu = p.cnu;
  const int nn[3] = {p.nx, p.ny, p.nz};
  const int ncl[3][2] = {{p.nclx1, p.nclxn}, {p.ncly1, p.nclyn}, {p.nclz1, p.nclzn}};
  const double len[3] = {p.xlx, p.yly, p.zlz};
  for (int a = 0; a < 3; ++a) {
    S->A[a] = make_axis_coeffs(nn[a], ncl[a][0], ncl[a][1], len[a], o);
    ctx.dc[a] = S->A[a].c;
    ctx.have_dc[a] = true;
    ctx.ncl[a] = S->A[a].periodic;
  }
  // Dirichlet faces (ncl = 2) are no-slip walls (zero wall velocity, Case-Channel.f90:67); inflow / outflow planes
  // would need the case arrays b?? (cylinder glue)
  if (p.istret != 0 && S->A[1].periodic) throw Error("x3d_solver_init: a stretched y mesh needs non-periodic y");
  if (p.itype != 0 && p.itype != 3 && p.itype != 5) throw Error("x3d_solver_init: itype 0 (box), 3 (channel) and 5 (cylinder) are implemented");
  if (p.itype == 5 && (p.nclx1 != 2 || p.nclxn != 2)) throw Error("x3d_solver_init: the cylinder case needs nclx1 = nclxn = 2 (inflow / outflow)");
  if (p.p_row > 0 && p.p_col > 0 && (p.p_row != pr || p.p_col != pc))
    throw Error("x3d_solver_init: p_row x p_col differs from the process grid of x3d_decomp_init");
  ctx.iibm = 0; ctx.istret = p.istret; ctx.iimplicit = 0;
  if (p.istret != 0) S->st = make_stretching(p.istret, p.beta, p.yly, p.ny, S->A[1].nm);
  S->nxm = S->A[0].nm; S->nym = S->A[1].nm; S->nzm = S->A[2].nm;
  S->xnu = 1.0 / p.re;  // parameters.f90:302
  const double dt = p.dt;
  if (p.itimescheme == 5) {  // variables.f90:1388-1399
    S->iadvance = 3; S->ntime = 2;
    S->adt[0] = (8.0 / 15.0) * dt; S->bdt[0] = 0.0; S->gdt[0] = S->adt[0];
    S->adt[1] = (5.0 / 12.0) * dt; S->bdt[1] = (-17.0 / 60.0) * dt; S->gdt[1] = S->adt[1] + S->bdt[1];
    S->adt[2] = (3.0 / 4.0) * dt; S->bdt[2] = (-5.0 / 12.0) * dt; S->gdt[2] = S->adt[2] + S->bdt[2];
  } else if (p.itimescheme == 1) {
    S->iadvance = 1; S->ntime = 1; S->adt[0] = dt; S->gdt[0] = dt;
  } else if (p.itimescheme == 2) {  // AB2, variables.f90:1355-1363
    S->iadvance = 1; S->ntime = 2; S->adt[0] = 1.5 * dt; S->bdt[0] = -0.5 * dt; S->gdt[0] = S->adt[0] + S->bdt[0];
  } else if (p.itimescheme == 3) {  // AB3, :1364-1374
    S->iadvance = 1; S->ntime = 3;
    S->adt[0] = (23.0 / 12.0) * dt; S->bdt[0] = -(16.0 / 12.0) * dt; S->cdt[0] = (5.0 / 12.0) * dt;
    S->gdt[0] = S->adt[0] + S->bdt[0] + S->cdt[0];
  } else {
    throw Error("x3d_solver_init: itimescheme 1 (Euler), 2 (AB2), 3 (AB3) and 5 (RK3) are implemented");
  }
  // local extents (xcompact3d.f90:191-201: main decomposition, ph3 = (nxm,nym,nz), ph1 = (nxm,nym,nzm))
  S->id_v = 0;
  S->id_p3 = decomp_info_init(ctx, S->nxm, S->nym, p.nz);
  S->id_p1 = decomp_info_init(ctx, S->nxm, S->nym, S->nzm);
  x3d_decomp_info iv{}, i3{}, i1{};
  decomp_info_get(ctx, S->id_v, &iv); decomp_info_get(ctx, S->id_p3, &i3); decomp_info_get(ctx, S->id_p1, &i1);
  if (iv.xsz[0] != p.nx || iv.ysz[1] != p.ny || iv.zsz[2] != p.nz) throw Error("x3d_solver_init: decomposition does not match the solver mesh");
  S->nzl = iv.xsz[2]; S->z0 = iv.xst[2] - 1;
  S->nyl = iv.zsz[1]; S->y0 = iv.zst[1] - 1;
  S->nyml = i3.zsz[1];
  if (i1.zsz[1] != S->nyml) throw Error("x3d_solver_init: inconsistent pressure decompositions");
  const int nzl = S->nzl, nyl = S->nyl, nyml = S->nyml;
  S->n = static_cast<size_t>(p.nx) * p.ny * nzl;
  S->n3 = static_cast<size_t>(S->nxm) * nyml * S->nzm;
  const size_t nmax = std::max({S->n, static_cast<size_t>(p.nx) * nyl * p.nz, static_cast<size_t>(1)});
  const size_t bytes = nmax * sizeof(double);
  for (DevBuf *b : {&S->ux, &S->uy, &S->uz, &S->px, &S->py, &S->pz, &S->pp3}) { b->reserve(bytes); X3D_CUDA(cudaMemsetAsync(b->p, 0, bytes, ctx.stream)); }
  for (int q = 0; q < S->ntime; ++q)
    for (DevBuf *b : {&S->dux[q], &S->duy[q], &S->duz[q]}) { b->reserve(bytes); X3D_CUDA(cudaMemsetAsync(b->p, 0, bytes, ctx.stream)); }
  for (auto &b : S->w) b.reserve(bytes);
  S->red_partial.reserve(sizeof(double) * 8 * 4096);
  S->red_out.reserve(sizeof(double) * 16);
  X3D_CUDA(cudaMallocHost(&S->h_red, sizeof(double) * 16));
  // operators on the local pencils
  const int dxy[3] = {p.nx, p.ny, nzl}, dzp[3] = {p.nx, nyl, p.nz};
  const bool stretched = p.istret != 0;
  const double *ppy = stretched ? S->st.ppy.data() : nullptr, *ppyi = stretched ? S->st.ppyi.data() : nullptr;
  for (int a = 0; a < 3; ++a) {
    const int *dd = (a == 2) ? dzp : dxy;
    // dery multiplies by ppy when the mesh is stretched (derive.f90:409-417)
    prep(ctx, S->d1[a][0], D1, a, S->A[a], 0, S->A[a].d1, dd, a == 1 ? ppy : nullptr);
    prep(ctx, S->d1[a][1], D1, a, S->A[a], 1, S->A[a].d1p, dd, a == 1 ? ppy : nullptr);
    prep(ctx, S->d2[a][0], D2, a, S->A[a], 0, S->A[a].d2, dd);
    prep(ctx, S->d2[a][1], D2, a, S->A[a], 1, S->A[a].d2p, dd);
  }
  // divergence chain (navier.f90:297-336): x on (nx,ny,nzl), y on (nxm,ny,nzl), z on (nxm,nyml,nz)
  const int dy[3] = {S->nxm, p.ny, nzl}, dz[3] = {S->nxm, nyml, p.nz};
  const int *dsv[3] = {dxy, dy, dz};
  for (int a = 0; a < 3; ++a) {
    prep(ctx, S->dvp[a], DVP, a, S->A[a], 0, S->A[a].vp, dsv[a], a == 1 ? ppyi : nullptr);   // deryvp * ppyi, derive.f90:4572-4580
    prep(ctx, S->ivp[a], IVP, a, S->A[a], 1, S->A[a].ivpp, dsv[a]);
  }
  // gradp chain (navier.f90:404-431): z on (nxm,nyml,nzm), y on (nxm,nym,nzl), x on (nxm,ny,nzl)
  const int gz[3] = {S->nxm, nyml, S->nzm}, gy[3] = {S->nxm, S->nym, nzl}, gx[3] = {S->nxm, p.ny, nzl};
  const int *gsv[3] = {gx, gy, gz};
  for (int a = 0; a < 3; ++a) {
    const AxisCoeffs &A = S->A[a];
    prep(ctx, S->dpv[a], DPV, a, A, 1, A.periodic ? A.vp : A.pvp, gsv[a], a == 1 ? ppy : nullptr);  // derypv * ppy, :4905-4913
    prep(ctx, S->ipv[a], IPV, a, A, 1, A.periodic ? A.ivp : A.ipvp, gsv[a]);
  }
  // fused momentum kernels where the direction is periodic and the tile kernel fits (x3d_mom.cu)
  {
    const char *e = getenv("X3D_FUSED");
    const bool want = !(e && atoi(e) == 0);
    for (int a = 0; a < 3 && want; ++a) {
      const long long lanes = (a == 1) ? p.nx : static_cast<long long>(p.nx) * nyl;
      if (!S->A[a].periodic || (p.nx & 1)) continue;
      const PreOp &o1 = S->d1[a][0], &o2 = S->d2[a][0];
      int segS, segH, nseg;
      if (!mom_segments(nn[a], o1.op.alpha, o2.op.alpha, segS, segH, nseg)) continue;
      const int n = segS + 2 * segH;           // rows of a tile: the line, or one overlap-save segment of a long line
      const int L = pick_L_contig(n);
      if (L <= 0) continue;
      if (a == 0 ? !mom_x_eligible(n, L) : (!mom_pair_eligible(n, L) || (lanes & 1))) continue;
      const char *ec = getenv("X3D_MOM_CYC");
      S->cyclic[a] = !(ec && atoi(ec) == 0) && mom_cyclic_ok(o1.op.alpha, n, L) && mom_cyclic_ok(o2.op.alpha, n, L);
      S->segmented[a] = nseg > 1;
      if (S->cyclic[a]) { S->fused[a] = true; continue; }
      if (nseg > 1) continue;                  // segments need the table-free solves
      const TriTable &T1 = get_tri(ctx, o1.call.f, o1.call.s, o1.call.w, n, L, true, o1.op.alpha, nullptr);
      const TriTable &T2 = get_tri(ctx, o2.call.f, o2.call.s, o2.call.w, n, L, true, o2.op.alpha, nullptr);
      S->fused[a] = build_mom_table(ctx, T1, S->mt1[a]) && build_mom_table(ctx, T2, S->mt2[a]);
    }
    if (const char *e2 = getenv("X3D_FUSE_INTT")) S->fuse_intt = atoi(e2) != 0;
  }
  {
    const char *e = getenv("X3D_FUSE_STAG");
    const bool want = !(e && atoi(e) == 0);
    const long long nxm = S->nxm;
    if (want) {
      // y lines of (nxm, ny, nzl) arrays; z lines of (nxm, nyml, nz) arrays
      S->stag[1] = S->A[1].periodic && stag_pair_eligible(ctx, S->ivp[1].op, S->dvp[1].op, nxm, p.ny, nxm, nxm * p.ny, nzl) &&
                   stag_pair_eligible(ctx, S->ipv[1].op, S->dpv[1].op, nxm, p.ny, nxm, nxm * p.ny, nzl);
      S->stag[2] = S->A[2].periodic && stag_pair_eligible(ctx, S->ivp[2].op, S->dvp[2].op, nxm * nyml, p.nz, nxm * nyml, 0, 1) &&
                   stag_pair_eligible(ctx, S->ipv[2].op, S->dpv[2].op, nxm * nyml, p.nz, nxm * nyml, 0, 1);
    }
  }
  S->ivp_y_add = S->ivp[1]; S->ivp_y_add.op.store_mode = 1;
  S->dvp_z_add = S->dvp[2]; S->dvp_z_add.op.store_mode = 1;
  S->dpv_x_sub = S->dpv[0]; S->dpv_x_sub.op.store_mode = 2;
  S->ipv_x_sub = S->ipv[0]; S->ipv_x_sub.op.store_mode = 2;
  if (const char *e = getenv("X3D_FUSE_SUMS")) S->fuse_sums = atoi(e) != 0;
  if (const char *e = getenv("X3D_OVERLAP")) S->overlap = atoi(e);
  if (const char *e = getenv("X3D_OVERLAP_DIV")) S->overlap_div = atoi(e) != 0;
  {
    // slab z kernels: periodic z, the fused kernels with table-free solves in all three directions, equal slabs
    const char *e = getenv("X3D_SLABZ"), *em = getenv("X3D_SLABZ_EMULATE");
    int nv = 1;
    bool want = !(e && atoi(e) == 0);
    if (S->nranks == 1) { nv = em ? atoi(em) : 1; want = want && nv > 1; }
    else want = want && ring_available(ctx);
    const long long plane = static_cast<long long>(p.nx) * p.ny;
    if (want && S->fused[1] && S->fused[2] && S->cyclic[2] && p.nz % (S->nranks * nv) == 0 && nzl * S->nranks == p.nz &&
        mom_slab_eligible(S->d1[2][0].op.alpha, S->d2[2][0].op.alpha, plane, nzl / nv)) {
      S->slabz = true;
      S->slab_nv = nv;
      for (auto &b : S->zhalo) { b.reserve(static_cast<size_t>(nv) * 16 * plane * sizeof(double)); X3D_CUDA(cudaMemsetAsync(b.p, 0, b.bytes, ctx.stream)); }
      S->zcarry_out.reserve(static_cast<size_t>(nv) * 18 * plane * sizeof(double));
      S->zcarry_in.reserve(static_cast<size_t>(nv) * 18 * plane * sizeof(double));
      X3D_CUDA(cudaMemsetAsync(S->zcarry_out.p, 0, S->zcarry_out.bytes, ctx.stream));
      X3D_CUDA(cudaMemsetAsync(S->zcarry_in.p, 0, S->zcarry_in.bytes, ctx.stream));
      build_zfix(ctx, S->d1[2][0].op, S->d2[2][0].op, S->xnu, nzl / nv, S->zfix);
    }
  }
  x3d_poisson_params pp{};
  pp.nx = p.nx; pp.ny = p.ny; pp.nz = p.nz;
  pp.bcx = S->A[0].periodic ? 0 : 1; pp.bcy = S->A[1].periodic ? 0 : 1; pp.bcz = S->A[2].periodic ? 0 : 1;
  pp.xlx = p.xlx; pp.yly = p.yly; pp.zlz = p.zlz; pp.istret = p.istret;
  pp.alpha = S->st.alpha; pp.beta = p.beta;
  poisson_init(ctx, pp);
  if (stretched) {
    std::vector<double> h(3 * static_cast<size_t>(p.ny));
    for (int j = 0; j < p.ny; ++j) { h[j] = S->st.pp2y[j]; h[p.ny + j] = S->st.pp4y[j]; h[2 * p.ny + j] = S->st.ppy[j]; }
    S->d_pp.reserve(h.size() * sizeof(double));
    X3D_CUDA(cudaMemcpyAsync(S->d_pp.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  if (p.nclx1 == 2 || p.nclxn == 2) {
    const size_t nb = 6 * std::max<size_t>(static_cast<size_t>(p.ny) * nzl, 1) * sizeof(double);
    S->bwx.reserve(nb);
    X3D_CUDA(cudaMemsetAsync(S->bwx.p, 0, S->bwx.bytes, ctx.stream));
  }
  S->cs.u1 = 1.0; S->cs.u2 = 1.0;
  {
    const size_t ndpd = 4 * (static_cast<size_t>(p.ny) * nzl + static_cast<size_t>(p.nx) * nzl + static_cast<size_t>(p.nx) * p.ny);
    S->dpd.reserve(std::max<size_t>(ndpd, 1) * sizeof(double));
    X3D_CUDA(cudaMemsetAsync(S->dpd.p, 0, S->dpd.bytes, ctx.stream));
  }
  ctx.solver = std::move(S);
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}

static SolverImpl &SOL(Ctx &ctx) {
  auto *S = dynamic_cast<SolverImpl *>(ctx.solver.get());
  if (!S) throw Error("solver: x3d_solver_init has not been called");
  return *S;
}

// pencil transposes of the reference; on one rank they are the identity and alias the source
static const double *TR(Ctx &ctx, SolverImpl &S, int which, const double *src, double *dst, int id) {
  if (S.nranks == 1) return src;
  transpose_device(ctx, which, src, dst, id, 1);
  return dst;
}

// Case-TGV.f90:90-95
void solver_init_tgv(Ctx &ctx) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const int nx = S.p.nx, ny = S.p.ny, z0 = S.z0;
  const double dx = S.A[0].d, dy = S.A[1].d, dz = S.A[2].d;
  double *ux = B(S.ux), *uy = B(S.uy), *uz = B(S.uz);
  map(ctx, S.n, [=] __device__(long long q) {
    const int i = static_cast<int>(q % nx), j = static_cast<int>((q / nx) % ny), k = static_cast<int>(q / (static_cast<long long>(nx) * ny)) + z0;
    const double x = static_cast<double>(i) * dx, y = static_cast<double>(j) * dy, z = static_cast<double>(k) * dz;
    ux[q] = sin(x) * cos(y) * cos(z);
    uy[q] = -cos(x) * sin(y) * cos(z);
    uz[q] = 0.0;
  });
  for (int q = 0; q < S.ntime; ++q)
    for (DevBuf *b : {&S.dux[q], &S.duy[q], &S.duz[q]}) X3D_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, ctx.stream));
  for (DevBuf *b : {&S.px, &S.py, &S.pz, &S.pp3}) X3D_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, ctx.stream));
  S.itime = 0;
}

// init_channel with iin = 0, Case-Channel.f90:71-94
void solver_init_channel(Ctx &ctx) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const int nx = S.p.nx, ny = S.p.ny, z0 = S.z0;
  const double dx = S.A[0].d, dy = S.A[1].d, dz = S.A[2].d, yly = S.p.yly;
  double *ux = B(S.ux), *uy = B(S.uy), *uz = B(S.uz);
  DevBuf ypd;
  ypd.reserve(sizeof(double) * ny);
  std::vector<double> yh(ny);
  for (int j = 0; j < ny; ++j) yh[j] = (S.p.istret == 0) ? static_cast<double>(j) * dy - yly * 0.5 : S.st.yp[j] - yly * 0.5;
  X3D_CUDA(cudaMemcpyAsync(ypd.p, yh.data(), sizeof(double) * ny, cudaMemcpyHostToDevice, ctx.stream));
  const double *yp = B(ypd);
  map(ctx, S.n, [=] __device__(long long q) {
    const int i = static_cast<int>(q % nx), j = static_cast<int>((q / nx) % ny), k = static_cast<int>(q / (static_cast<long long>(nx) * ny)) + z0;
    const double y = yp[j];
    ux[q] = 1.0 - y * y;
    uy[q] = 0.0;
    uz[q] = sin(static_cast<double>(i) * dx) + cos(static_cast<double>(k) * dz);
  });
  for (int q = 0; q < S.ntime; ++q)
    for (DevBuf *b : {&S.dux[q], &S.duy[q], &S.duz[q]}) X3D_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, ctx.stream));
  for (DevBuf *b : {&S.px, &S.py, &S.pz, &S.pp3, &S.dpd}) X3D_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));  // ypd goes out of scope
  S.itime = 0;
}

// boundary_conditions_channel (Case-Channel.f90:150-170, cpg = F, idir_stream = 1): channel_cfr(ux, 2/3), :220-261
static void boundary_conditions(Ctx &ctx, SolverImpl &S, int itr) {
  if (S.p.itype == 5) {  // boundary_conditions_cyl, Case-Cylinder-wake.f90:84-98: inflow (:100-133), outflow (:135-203)
    const int nx = S.p.nx, ny = S.p.ny, nzl = S.nzl;
    const long long nyz = static_cast<long long>(ny) * nzl;
    if (nyz == 0) return;
    double *bw = B(S.bwx);
    const double *bn = S.bnoise.p ? B(S.bnoise) : nullptr;
    const double u1 = S.cs.u1, u2 = S.cs.u2, noise = S.cs.inflow_noise;
    const double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
    double *mm = B(S.red_out) + 10;
    k_plane_minmax<<<1, 1024, 0, ctx.stream>>>(u + (nx - 2), nyz, nx, mm);   // uxmax / uxmin over the plane nx - 1 (1-based)
    X3D_CUDA(cudaGetLastError()); ctx.launches++;
    allreduce(ctx, mm, 2, true);
    const double g = S.gdt[itr - 1], udx = 1.0 / S.A[0].d;
    const int mode = u1 == 0.0 ? 0 : (u1 == 1.0 ? 1 : (u1 == 2.0 ? 2 : 3));
    map(ctx, nyz, [=] __device__(long long q) {
      bw[q] = u1 + (bn ? bn[q] : 0.0) * noise;
      bw[nyz + q] = 0.0 + (bn ? bn[nyz + q] : 0.0) * noise;
      bw[2 * nyz + q] = 0.0 + (bn ? bn[2 * nyz + q] : 0.0) * noise;
      const double uxmax = mm[0], uxmin = -mm[1];
      double cx;
      if (mode == 0) cx = (0.5 * (uxmax + uxmin)) * g * udx;
      else if (mode == 1) cx = uxmax * g * udx;
      else if (mode == 2) cx = u2 * g * udx;
      else cx = (0.5 * (u1 + u2)) * g * udx;
      const long long p1 = q * nx + nx - 1, p0 = p1 - 1;
      bw[3 * nyz + q] = u[p1] - cx * (u[p1] - u[p0]);
      bw[4 * nyz + q] = v[p1] - cx * (v[p1] - v[p0]);
      bw[5 * nyz + q] = w[p1] - cx * (w[p1] - w[p0]);
    });
    return;
  }
  if (S.p.itype != 3 || S.cs.cpg) return;   // Case-Channel.f90:157: channel_cfr only without a constant pressure gradient
  const long long n = static_cast<long long>(S.n);
  const int nx = S.p.nx, ny = S.p.ny;
  double *u = B(S.ux);
  const double *ippy = S.p.istret ? B(S.d_pp) + 2 * ny : nullptr;
  double *partial = B(S.red_partial), *dout = B(S.red_out);
  const int nb = gridn(ctx, n) > 4096 ? 4096 : gridn(ctx, n);
  k_reduce_partial<1><<<nb, 256, 0, ctx.stream>>>(n, [=] __device__(long long q, double *acc) {
    const int j = static_cast<int>((q / nx) % ny);
    acc[0] += ippy ? u[q] / ippy[j] : u[q];
  }, partial);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  k_reduce_final<1><<<1, 256, 0, ctx.stream>>>(nb, partial, dout + 12);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  allreduce(ctx, dout + 12, 1, false);
  const double coeff = S.A[1].d / (S.p.yly * static_cast<double>(nx) * static_cast<double>(S.p.nz));
  const double constant = 2.0 / 3.0;
  const double *ub = dout + 12;
  map(ctx, n, [=] __device__(long long q) { u[q] = u[q] - (-(constant - ub[0] * coeff)); });
}

// transeq.f90:73-591 (incompressible, explicit diffusion, uniform mesh)
static void momentum_rhs(Ctx &ctx, SolverImpl &S, double *dux1, double *duy1, double *duz1) {
  const long long n = static_cast<long long>(S.n);                                   // x / y pencil
  const long long nz3 = static_cast<long long>(S.p.nx) * S.nyl * S.p.nz;             // z pencil
  double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
  // iibm = 2 / 3: every collocated derivative first rebuilds its input inside the bodies, in place -- the velocity
  // arrays included (src/derive.f90:23-24); lind = product of the body velocities (src/transeq.f90:120-146)
  const int iibm = S.cs.iibm;
  const double bx = S.cs.ubcx, by = S.cs.ubcy, bz = S.cs.ubcz;
  auto run = [&](const PreOp &P, const double *in, double *out, double lind = 0.0) {
    if (iibm == 2) lagpol_device(ctx, P.call.axis, const_cast<double *>(in), P.call.dims_in[0], P.call.dims_in[1], P.call.dims_in[2]);
    else if (iibm == 3) cubspl_device(ctx, P.call.axis, const_cast<double *>(in), P.call.dims_in[0], P.call.dims_in[1], P.call.dims_in[2], lind);
    x3d::run(ctx, P, in, out);
  };
  double *ta = B(S.w[0]), *tb = B(S.w[1]), *tc = B(S.w[2]), *td = B(S.w[3]), *te = B(S.w[4]), *tf = B(S.w[5]);
  double *tg1 = B(S.w[6]), *th1 = B(S.w[7]), *ti1 = B(S.w[8]), *tg2 = B(S.w[9]), *th2 = B(S.w[10]), *ti2 = B(S.w[11]);
  double *tg3 = B(S.w[12]), *th3 = B(S.w[13]), *ti3 = B(S.w[14]);
  const double xnu = S.xnu, half = 0.5;
  // x, :114-146
  map(ctx, n, [=] __device__(long long q) { const double a = u[q]; ta[q] = a * a; tb[q] = a * v[q]; tc[q] = a * w[q]; });
  run(S.d1[0][1], ta, td, bx * bx); run(S.d1[0][0], tb, te, bx * by); run(S.d1[0][0], tc, tf, bx * bz);
  run(S.d1[0][0], u, ta, bx); run(S.d1[0][1], v, tb, by); run(S.d1[0][1], w, tc, bz);
  map(ctx, n, [=] __device__(long long q) { const double a = u[q]; tg1[q] = td[q] + a * ta[q]; th1[q] = te[q] + a * tb[q]; ti1[q] = tf[q] + a * tc[q]; });
  // y, :188-219
  map(ctx, n, [=] __device__(long long q) { const double a = v[q]; td[q] = u[q] * a; te[q] = a * a; tf[q] = w[q] * a; });
  run(S.d1[1][0], td, tg2, bx * by); run(S.d1[1][1], te, th2, by * by); run(S.d1[1][0], tf, ti2, bz * by);
  run(S.d1[1][1], u, td, bx); run(S.d1[1][0], v, te, by); run(S.d1[1][1], w, tf, bz);
  map(ctx, n, [=] __device__(long long q) { const double a = v[q]; tg2[q] = tg2[q] + a * td[q]; th2[q] = th2[q] + a * te[q]; ti2[q] = ti2[q] + a * tf[q]; });
  // z, :236-314 (transpose_y_to_z of the three velocity components, :236-238)
  const double *u3 = TR(ctx, S, 1, u, ta, S.id_v), *v3 = TR(ctx, S, 1, v, tb, S.id_v), *w3 = TR(ctx, S, 1, w, tc, S.id_v);
  {
    double *pa = td, *pb = te, *pc = tf;
    map(ctx, nz3, [=] __device__(long long q) { const double a = w3[q]; pa[q] = u3[q] * a; pb[q] = v3[q] * a; pc[q] = a * a; });
  }
  run(S.d1[2][0], td, tg3, bx * bz); run(S.d1[2][0], te, th3, by * bz); run(S.d1[2][1], tf, ti3, bz * bz);
  run(S.d1[2][1], u3, td, bx); run(S.d1[2][1], v3, te, by); run(S.d1[2][0], w3, tf, bz);
  map(ctx, nz3, [=] __device__(long long q) {  // convective z terms, :272-274
    const double a = w3[q];
    tg3[q] = tg3[q] + a * td[q]; th3[q] = th3[q] + a * te[q]; ti3[q] = ti3[q] + a * tf[q];
  });
  run(S.d2[2][1], u3, td, bx); run(S.d2[2][1], v3, te, by); run(S.d2[2][0], w3, tf, bz);  // :301-303
  map(ctx, nz3, [=] __device__(long long q) {  // :312-314
    tg3[q] = xnu * td[q] - half * tg3[q]; th3[q] = xnu * te[q] - half * th3[q]; ti3[q] = xnu * tf[q] - half * ti3[q];
  });
  // back to y pencils, :318-325
  const double *zx = TR(ctx, S, 2, tg3, td, S.id_v), *zy = TR(ctx, S, 2, th3, te, S.id_v), *zz = TR(ctx, S, 2, ti3, tf, S.id_v);
  map(ctx, n, [=] __device__(long long q) { tg2[q] = zx[q] - half * tg2[q]; th2[q] = zy[q] - half * th2[q]; ti2[q] = zz[q] - half * ti2[q]; });
  // y diffusion, :336-433
  run(S.d2[1][1], u, td, bx); run(S.d2[1][0], v, te, by); run(S.d2[1][1], w, tf, bz);
  if (S.p.istret != 0) {  // td = td pp2y - pp4y dery(u), :339-372 (the dery carries its ppy factor)
    const double *pp2 = B(S.d_pp), *pp4 = B(S.d_pp) + S.p.ny;
    const int nx_ = S.p.nx, ny_ = S.p.ny;
    double *tj = B(S.w[15]);
    const PreOp *ops[3] = {&S.d1[1][1], &S.d1[1][0], &S.d1[1][1]};
    const double *fld[3] = {u, v, w};
    double *dst[3] = {td, te, tf};
    const double lin[3] = {bx, by, bz};
    for (int c = 0; c < 3; ++c) {
      run(*ops[c], fld[c], tj, lin[c]);
      double *t2 = dst[c];
      map(ctx, n, [=] __device__(long long q) {
        const int j = static_cast<int>((q / nx_) % ny_);
        t2[q] = t2[q] * pp2[j] - pp4[j] * tj[q];
      });
    }
  }
  // x diffusion and final sum, :442-470
  run(S.d2[0][0], u, ta, bx); run(S.d2[0][1], v, tb, by); run(S.d2[0][1], w, tc, bz);
  map(ctx, n, [=] __device__(long long q) {
    const double ax = xnu * td[q] + tg2[q], ay = xnu * te[q] + th2[q], az = xnu * tf[q] + ti2[q];
    dux1[q] = ax - half * tg1[q] + xnu * ta[q];
    duy1[q] = ay - half * th1[q] + xnu * tb[q];
    duz1[q] = az - half * ti1[q] + xnu * tc[q];
  });
  // momentum_forcing_channel (the end of momentum_rhs_eq, src/transeq.f90:539 -> src/Case-Channel.f90:396-420, idir_stream = 1)
  if (S.p.itype == 3) {
    if (S.cs.cpg) {
      const double f = S.fcpg;
      map(ctx, n, [=] __device__(long long q) { dux1[q] = dux1[q] + f; });
    }
    if (S.itime < S.cs.spinup_time && S.cs.iin <= 2 && S.cs.wrotation != 0.0) {
      const double wr = S.cs.wrotation;
      map(ctx, n, [=] __device__(long long q) { dux1[q] = dux1[q] - wr * v[q]; duy1[q] = duy1[q] + wr * u[q]; });
    }
  }
}

static void ensure_aux(SolverImpl &S) {
  if (S.aux) return;
  int lo = 0, hi = 0;
  X3D_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  X3D_CUDA(cudaStreamCreateWithPriority(&S.aux, cudaStreamNonBlocking, hi));
  X3D_CUDA(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
  X3D_CUDA(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming));
}
// run f with ctx.stream = S.aux (the transposes take their stream from the context)
template <class F>
static void on_aux(Ctx &ctx, SolverImpl &S, F f) {
  cudaStream_t main_stream = ctx.stream;
  ctx.stream = S.aux;
  try { f(); } catch (...) { ctx.stream = main_stream; throw; }
  ctx.stream = main_stream;
}

// Fused form of momentum_rhs for periodic y and z: per direction one kernel forms
// r_c = xnu D2(c) - 1/2 (D1(c a) + a D1(c)).
//  * one rank: the z kernel stores its r_c into sum[c], the y and x kernels add theirs with TMA reduce-add stores, so
//    that intt reads one array per component: sum = (r_z + r_y) + r_x;
//  * several ranks: the y -> z transposes of u, v, w run on a second stream (copy engines / peer stores) while the
//    y kernel stores and the x kernel adds into sum[c]; the z kernel follows and its result comes back through the
//    z -> y transposes into extra[c], which intt adds: extra + (r_y + r_x).
// Same terms as transeq.f90:312-314,323-325,460-470, summed in a different order.
// itr > 0: the caller allows the time integration of sub-step itr to be folded into the x kernel; returns true when it was
static bool momentum_rhs_fused(Ctx &ctx, SolverImpl &S, double *sum[3], double *extra[3], int itr) {
  const long long n = static_cast<long long>(S.n);
  const double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
  const double xnu = S.xnu, half = 0.5;
  const int nx = S.p.nx, ny = S.p.ny, nz = S.p.nz;
  const double *f[3] = {u, v, w};
  const bool alias = S.nranks == 1;
  double *t[3] = {B(S.w[9]), B(S.w[10]), B(S.w[11])}, *o[3] = {B(S.w[12]), B(S.w[13]), B(S.w[14])};
  const long long lanes = static_cast<long long>(nx) * S.nyl;
  extra[0] = extra[1] = extra[2] = nullptr;
  bool z_first = true;
  const bool slab = S.slabz;
  const long long plane = static_cast<long long>(nx) * ny;
  const int nv = S.slab_nv, nzv = S.nzl / (nv > 0 ? nv : 1);
  if (slab) {
    // transeq.f90:236-320 without the transposes: every rank differentiates its own planes of the z lines (zero-carry solves),
    // the face corrections follow once the carries of the neighbours are there (x3d_slab_kernels.cuh)
    double *hz[3] = {B(S.zhalo[0]), B(S.zhalo[1]), B(S.zhalo[2])};
    double *co = B(S.zcarry_out);
    const size_t pb = static_cast<size_t>(plane) * sizeof(double);
    if (S.nranks > 1) {
      RingCopy rc[6];
      for (int c = 0; c < 3; ++c) {
        rc[2 * c] = RingCopy{f[c] + static_cast<long long>(S.nzl - 4) * plane, hz[c], 4 * pb, 4 * pb, +1};   // my last planes: below the next slab
        rc[2 * c + 1] = RingCopy{f[c], hz[c], 8 * pb, 4 * pb, -1};                                            // my first planes: above the previous slab
      }
      if (!ring_exchange(ctx, 6, rc)) throw Error("solver: slab z kernels without a peer-to-peer path");
    } else {
      for (int v = 0; v < nv; ++v)
        for (int c = 0; c < 3; ++c) {
          const int vp = (v + nv - 1) % nv, vn = (v + 1) % nv;
          double *h = hz[c] + static_cast<long long>(v) * 16 * plane;
          X3D_CUDA(cudaMemcpyAsync(h + 4 * plane, f[c] + (static_cast<long long>(vp) * nzv + nzv - 4) * plane, 4 * pb, cudaMemcpyDeviceToDevice, ctx.stream));
          X3D_CUDA(cudaMemcpyAsync(h + 8 * plane, f[c] + static_cast<long long>(vn) * nzv * plane, 4 * pb, cudaMemcpyDeviceToDevice, ctx.stream));
        }
    }
    for (int v = 0; v < nv; ++v) {
      const long long o = static_cast<long long>(v) * nzv * plane;
      const double *fv[3] = {f[0] + o, f[1] + o, f[2] + o};
      const double *hv[3] = {hz[0] + static_cast<long long>(v) * 16 * plane, hz[1] + static_cast<long long>(v) * 16 * plane, hz[2] + static_cast<long long>(v) * 16 * plane};
      double *ov[3] = {sum[0] + o, sum[1] + o, sum[2] + o};
      launch_mom_slab(ctx, S.d1[2][0].op, S.d2[2][0].op, xnu, fv, hv, ov, plane, nzv, false, co + static_cast<long long>(v) * 18 * plane);
    }
    if (S.nranks > 1) {
      RingCopy rc[2] = {RingCopy{co, B(S.zcarry_in), 0, 9 * pb, +1},                    // Yout -> Yin of the next slab
                        RingCopy{co + 9 * plane, B(S.zcarry_in), 9 * pb, 9 * pb, -1}};  // Z0 -> the previous slab
      if (!ring_exchange(ctx, 2, rc)) throw Error("solver: slab z kernels without a peer-to-peer path");
    }
  }
  if (!alias && !slab) {  // transpose_y_to_z of the three components, one barrier pair (transeq.f90:236-238)
    if (S.overlap) {
      if (!S.aux) {
        int lo = 0, hi = 0;
        X3D_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        X3D_CUDA(cudaStreamCreateWithPriority(&S.aux, cudaStreamNonBlocking, hi));
        X3D_CUDA(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
        X3D_CUDA(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming));
      }
      cudaStream_t main_stream = ctx.stream;
      X3D_CUDA(cudaEventRecord(S.ev_fork, main_stream));
      X3D_CUDA(cudaStreamWaitEvent(S.aux, S.ev_fork, 0));
      ctx.stream = S.aux;
      try {
        transpose_device_multi(ctx, 1, 3, f, t, S.id_v, 1);
      } catch (...) { ctx.stream = main_stream; throw; }
      ctx.stream = main_stream;
      X3D_CUDA(cudaEventRecord(S.ev_join, S.aux));
      z_first = false;
    } else {
      transpose_device_multi(ctx, 1, 3, f, t, S.id_v, 1);
    }
  }
  auto z_dir = [&](bool into_sum) {  // transeq.f90:236-314 (z pencils)
    const double *fz[3] = {alias ? u : t[0], alias ? v : t[1], alias ? w : t[2]};
    double *oz[3] = {alias ? sum[0] : o[0], alias ? sum[1] : o[1], alias ? sum[2] : o[2]};
    launch_mom_pair(ctx, 2, S.d1[2][0].op, S.d2[2][0].op, S.mt1[2], S.mt2[2], xnu, fz, oz, lanes, nz, 1, lanes, lanes * nz, false, S.cyclic[2]);
    if (!alias) {  // transeq.f90:318-320
      if (into_sum) {
        transpose_device_multi(ctx, 2, 3, o, sum, S.id_v, 1);
      } else {
        transpose_device_multi(ctx, 2, 3, o, t, S.id_v, 1);
        extra[0] = t[0]; extra[1] = t[1]; extra[2] = t[2];
      }
    }
  };
  if (z_first && !slab) z_dir(true);
  // ---- y, transeq.f90:188-219,336-338
  launch_mom_pair(ctx, 1, S.d1[1][0].op, S.d2[1][0].op, S.mt1[1], S.mt2[1], xnu, f, sum, nx, ny, S.nzl, nx, static_cast<long long>(nx) * ny, z_first, S.cyclic[1]);
  if (slab) {   // face corrections of the z part: A Yin + B Zin on the rows next to the slab faces
    const double *co = B(S.zcarry_out), *ci = B(S.zcarry_in);
    for (int v = 0; v < nv; ++v) {
      const long long o = static_cast<long long>(v) * nzv * plane;
      const double *yin = S.nranks > 1 ? ci : co + static_cast<long long>((v + nv - 1) % nv) * 18 * plane;
      const double *z0n = S.nranks > 1 ? ci + 9 * plane : co + (static_cast<long long>((v + 1) % nv) * 18 + 9) * plane;
      double *sv[3] = {sum[0] + o, sum[1] + o, sum[2] + o};
      launch_zfix(ctx, S.zfix, yin, z0n, co + static_cast<long long>(v) * 18 * plane, sv, w + o, plane);
    }
  }
  bool z_pending = !z_first;
  if (z_pending && S.overlap == 2) {   // the forward transposes ran beside the y kernel: z now, x (+ intt) last
    X3D_CUDA(cudaStreamWaitEvent(ctx.stream, S.ev_join, 0));
    z_dir(false);
    z_pending = false;
  }
  // ---- x, transeq.f90:114-146,442-444
  bool folded = false;
  if (S.fused[0] && S.cyclic[0] && S.fuse_intt && itr > 0 && !z_pending && (S.p.itimescheme == 5 || S.p.itimescheme == 1)) {
    // x kernel + intt (time_integrators.f90:71-74,151-157): u <- adt N + bdt old + u, old <- N, N = sum (+ extra) + r_x
    MomIntt I{};
    double *vel[3] = {B(S.ux), B(S.uy), B(S.uz)}, *old[3] = {B(S.dux[1]), B(S.duy[1]), B(S.duz[1])};
    const bool rk3 = S.p.itimescheme == 5;
    I.has_extra = extra[0] != nullptr;
    I.use_old = rk3 && itr > 1;
    I.store_old = rk3 && itr < S.iadvance;   // the last sub-step's right-hand side is never read (bdt(1) = 0)
    I.ca = (!rk3 || itr == 1) ? S.gdt[0] : S.adt[itr - 1];
    I.cb = I.use_old ? S.bdt[itr - 1] : 0.0;
    for (int c = 0; c < 3; ++c) { I.sum[c] = sum[c]; I.extra[c] = extra[c]; I.old_in[c] = old[c]; I.u[c] = vel[c]; I.u_out[c] = vel[c]; I.old_out[c] = old[c]; }
    DevBuf *vb[3] = {&S.ux, &S.uy, &S.uz};
    if (S.segmented[0])   // neighbouring tiles of a segmented line still read the old velocity: write the new one elsewhere, then swap
      for (int c = 0; c < 3; ++c) { S.unew[c].reserve(vb[c]->bytes); I.u_out[c] = B(S.unew[c]); }
    launch_mom_x(ctx, S.d1[0][0].op, S.d2[0][0].op, S.mt1[0], S.mt2[0], xnu, f, sum, nx, static_cast<long long>(ny) * S.nzl, true, true, &I);
    if (S.segmented[0])
      for (int c = 0; c < 3; ++c) { std::swap(vb[c]->p, S.unew[c].p); std::swap(vb[c]->bytes, S.unew[c].bytes); }
    folded = true;
  } else if (S.fused[0]) {
    launch_mom_x(ctx, S.d1[0][0].op, S.d2[0][0].op, S.mt1[0], S.mt2[0], xnu, f, sum, nx, static_cast<long long>(ny) * S.nzl, true, S.cyclic[0]);
  } else {  // operator kernels + three elementwise passes
    double *ta = B(S.w[3]), *tb = B(S.w[4]), *tc = B(S.w[5]), *td = B(S.w[6]), *te = B(S.w[7]), *tf = B(S.w[8]);
    double *rx = sum[0], *ry = sum[1], *rz = sum[2];
    map(ctx, n, [=] __device__(long long q) { const double a = u[q]; ta[q] = a * a; tb[q] = a * v[q]; tc[q] = a * w[q]; });
    run(ctx, S.d1[0][1], ta, td); run(ctx, S.d1[0][0], tb, te); run(ctx, S.d1[0][0], tc, tf);
    run(ctx, S.d1[0][0], u, ta); run(ctx, S.d1[0][1], v, tb); run(ctx, S.d1[0][1], w, tc);
    map(ctx, n, [=] __device__(long long q) { const double a = u[q]; td[q] = td[q] + a * ta[q]; te[q] = te[q] + a * tb[q]; tf[q] = tf[q] + a * tc[q]; });
    run(ctx, S.d2[0][0], u, ta); run(ctx, S.d2[0][1], v, tb); run(ctx, S.d2[0][1], w, tc);
    map(ctx, n, [=] __device__(long long q) {
      rx[q] = rx[q] + (xnu * ta[q] - half * td[q]); ry[q] = ry[q] + (xnu * tb[q] - half * te[q]); rz[q] = rz[q] + (xnu * tc[q] - half * tf[q]);
    });
  }
  if (z_pending) {
    X3D_CUDA(cudaStreamWaitEvent(ctx.stream, S.ev_join, 0));
    z_dir(false);
  }
  return folded;
}

// Adams-Bashforth 2 / 3 (time_integrators.f90:75-100) for the three components; rhs(q, c) is the current right-hand
// side of component c.  First step Euler, second step of AB3 is AB2; the stored right-hand sides shift afterwards.
template <class F>
static void adams_bashforth(Ctx &ctx, SolverImpl &S, F rhs) {
  const long long n = static_cast<long long>(S.n);
  double *fld[3] = {B(S.ux), B(S.uy), B(S.uz)};
  double *h2[3] = {B(S.dux[1]), B(S.duy[1]), B(S.duz[1])};
  double *h3[3] = {S.ntime > 2 ? B(S.dux[2]) : nullptr, S.ntime > 2 ? B(S.duy[2]) : nullptr, S.ntime > 2 ? B(S.duz[2]) : nullptr};
  const double dt = S.p.dt, a = S.adt[0], b = S.bdt[0], c3 = S.cdt[0], g = S.gdt[0];
  const bool ab3 = S.p.itimescheme == 3;
  const int mode = S.itime == 1 ? 0 : ((ab3 && S.itime == 2) ? 1 : (ab3 ? 3 : 2));   // 0 Euler, 1 AB2 start of AB3, 2 AB2, 3 AB3
  for (int c = 0; c < 3; ++c) {
    double *u = fld[c], *d2 = h2[c], *d3 = h3[c];
    map(ctx, n, [=] __device__(long long q) {
      const double x = rhs(q, c);
      if (mode == 0) u[q] = (ab3 ? dt : g) * x + u[q];
      else if (mode == 1) { u[q] = 1.5 * dt * x - 0.5 * dt * d2[q] + u[q]; d3[q] = d2[q]; }
      else if (mode == 2) u[q] = a * x + b * d2[q] + u[q];
      else { u[q] = a * x + b * d2[q] + c3 * d3[q] + u[q]; d3[q] = d2[q]; }
      d2[q] = x;
    });
  }
}

// right-hand side of component c at point q in the fused form: sum[c], plus extra[c] when the z part came back
// through a transpose (several ranks)
struct FusedRhs {
  const double *r[3], *e[3];
  __device__ __forceinline__ double operator()(long long q, int c) const { return e[0] ? e[c][q] + r[c][q] : r[c][q]; }
};

// intt for the fused form (dux1 of the reference is formed on the fly)
static void intt3_fused(Ctx &ctx, SolverImpl &S, int itr, double *sum[3], double *extra[3]) {
  const long long n = static_cast<long long>(S.n);
  double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
  const FusedRhs R{{sum[0], sum[1], sum[2]}, {extra[0], extra[1], extra[2]}};
  if (S.p.itimescheme == 1) {
    const double g = S.gdt[0];
    map(ctx, n, [=] __device__(long long q) { u[q] = g * R(q, 0) + u[q]; v[q] = g * R(q, 1) + v[q]; w[q] = g * R(q, 2) + w[q]; });
    return;
  }
  double *a2 = B(S.dux[1]), *b2 = B(S.duy[1]), *c2 = B(S.duz[1]);
  if (S.p.itimescheme == 2 || S.p.itimescheme == 3) {
    adams_bashforth(ctx, S, [=] __device__(long long q, int c) { return R(q, c); });
    return;
  }
  if (itr == 1) {
    const double g = S.gdt[0];
    map(ctx, n, [=] __device__(long long q) {
      const double x = R(q, 0), y = R(q, 1), z = R(q, 2);
      u[q] = g * x + u[q]; v[q] = g * y + v[q]; w[q] = g * z + w[q];
      a2[q] = x; b2[q] = y; c2[q] = z;
    });
  } else if (itr < S.iadvance) {
    const double a = S.adt[itr - 1], b = S.bdt[itr - 1];
    map(ctx, n, [=] __device__(long long q) {
      const double x = R(q, 0), y = R(q, 1), z = R(q, 2);
      u[q] = a * x + b * a2[q] + u[q]; v[q] = a * y + b * b2[q] + v[q]; w[q] = a * z + b * c2[q] + w[q];
      a2[q] = x; b2[q] = y; c2[q] = z;
    });
  } else {
    // last sub-step: the next one (itr = 1 of the following step) has bdt = 0 and never reads the stored
    // right-hand side (time_integrators.f90:151-154), so it is not written
    const double a = S.adt[itr - 1], b = S.bdt[itr - 1];
    map(ctx, n, [=] __device__(long long q) {
      u[q] = a * R(q, 0) + b * a2[q] + u[q]; v[q] = a * R(q, 1) + b * b2[q] + v[q]; w[q] = a * R(q, 2) + b * c2[q] + w[q];
    });
  }
}

// time_integrators.f90:71-74,151-157 for the three components at once
static void intt3(Ctx &ctx, SolverImpl &S, int itr) {
  const long long n = static_cast<long long>(S.n);
  double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
  const double *a1 = B(S.dux[0]), *b1 = B(S.duy[0]), *c1 = B(S.duz[0]);
  if (S.p.itimescheme == 1) {
    const double g = S.gdt[0];
    map(ctx, n, [=] __device__(long long q) { u[q] = g * a1[q] + u[q]; v[q] = g * b1[q] + v[q]; w[q] = g * c1[q] + w[q]; });
    return;
  }
  double *a2 = B(S.dux[1]), *b2 = B(S.duy[1]), *c2 = B(S.duz[1]);
  if (S.p.itimescheme == 2 || S.p.itimescheme == 3) {
    adams_bashforth(ctx, S, [=] __device__(long long q, int c) { return c == 0 ? a1[q] : (c == 1 ? b1[q] : c1[q]); });
    return;
  }
  if (itr == 1) {
    const double g = S.gdt[0];
    map(ctx, n, [=] __device__(long long q) {
      const double x = a1[q], y = b1[q], z = c1[q];
      u[q] = g * x + u[q]; v[q] = g * y + v[q]; w[q] = g * z + w[q];
      a2[q] = x; b2[q] = y; c2[q] = z;
    });
  } else {
    const double a = S.adt[itr - 1], b = S.bdt[itr - 1];
    map(ctx, n, [=] __device__(long long q) {
      const double x = a1[q], y = b1[q], z = c1[q];
      u[q] = a * x + b * a2[q] + u[q]; v[q] = a * y + b * b2[q] + v[q]; w[q] = a * z + b * c2[q] + w[q];
      a2[q] = x; b2[q] = y; c2[q] = z;
    });
  }
}

// navier.f90:599-613,693-711,751-769 (free-slip faces); the x/y pencil holds z planes z0 .. z0+nzl-1
static void pre_correc(Ctx &ctx, SolverImpl &S, int itr) {
  const int nx = S.p.nx, ny = S.p.ny, nzl = S.nzl;
  double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
  const bool x1 = S.p.nclx1 == 1, xn = S.p.nclxn == 1, y1 = S.p.ncly1 == 1, yn = S.p.nclyn == 1;
  const bool z1 = S.p.nclz1 == 1 && S.z0 == 0, zn = S.p.nclzn == 1 && S.z0 + nzl == S.p.nz;
  // ---- Dirichlet (no-slip) faces, navier.f90:560-589,616-689,712-745: wall velocity 0, tangential components
  //      + wall pressure gradient of the previous gradp times gdt(itr)
  {
    const double g = S.gdt[itr - 1];
    double *d = B(S.dpd);
    const long long nyz = static_cast<long long>(ny) * nzl, nxz = static_cast<long long>(nx) * nzl, nxy = static_cast<long long>(nx) * ny;
    double *dx_ = d, *dy_ = d + 4 * nyz, *dz_ = d + 4 * nyz + 4 * nxz;
    const bool dx1 = S.p.nclx1 == 2, dxn = S.p.nclxn == 2, dy1 = S.p.ncly1 == 2, dyn = S.p.nclyn == 2;
    const bool dz1 = S.p.nclz1 == 2 && S.z0 == 0, dzn = S.p.nclzn == 2 && S.z0 + nzl == S.p.nz;
    double *bw = B(S.bwx);   // bxx1 bxy1 bxz1 bxxn bxyn bxzn
    // inflow / outflow flow-rate balance, navier.f90:534-560 (channel, uniform, abl with nclx = 2 on both sides; not itype_cyl)
    if (S.p.itype == 3 && dx1 && dxn && nyz > 0) {
      double *partial = B(S.red_partial), *dout = B(S.red_out) + 13;
      const int nb = gridn(ctx, nyz) > 4096 ? 4096 : gridn(ctx, nyz);
      k_reduce_partial<2><<<nb, 256, 0, ctx.stream>>>(nyz, [=] __device__(long long q, double *acc) { acc[0] += bw[q]; acc[1] += bw[3 * nyz + q]; }, partial);
      X3D_CUDA(cudaGetLastError()); ctx.launches++;
      k_reduce_final<2><<<1, 256, 0, ctx.stream>>>(nb, partial, dout);
      X3D_CUDA(cudaGetLastError()); ctx.launches++;
      allreduce(ctx, dout, 2, false);
      const double inv = 1.0 / (static_cast<double>(ny) * static_cast<double>(S.p.nz));
      map(ctx, nyz, [=] __device__(long long q) { bw[3 * nyz + q] = bw[3 * nyz + q] - dout[1] * inv + dout[0] * inv; });
    }
    if (dx1 || dxn)
      map(ctx, nyz, [=] __device__(long long q) {  // q = j + ny k ; planes: dpdyx1, dpdzx1, dpdyxn, dpdzxn (navier.f90:564-595)
        if (dx1) { const double a = dx_[q] * g, b = dx_[nyz + q] * g; dx_[q] = a; dx_[nyz + q] = b;
                   u[q * nx] = bw[q]; v[q * nx] = bw[nyz + q] + a; w[q * nx] = bw[2 * nyz + q] + b; }
        if (dxn) { const double a = dx_[2 * nyz + q] * g, b = dx_[3 * nyz + q] * g; dx_[2 * nyz + q] = a; dx_[3 * nyz + q] = b;
                   u[q * nx + nx - 1] = bw[3 * nyz + q]; v[q * nx + nx - 1] = bw[4 * nyz + q] + a; w[q * nx + nx - 1] = bw[5 * nyz + q] + b; }
      });
    if (dy1 || dyn)
      map(ctx, nxz, [=] __device__(long long q) {  // q = i + nx k ; planes: dpdxy1, dpdzy1, dpdxyn, dpdzyn
        const long long i = q % nx, k = q / nx;
        if (dy1) { const long long p = i + static_cast<long long>(nx) * ny * k; const double a = dy_[q] * g, b = dy_[nxz + q] * g;
                   dy_[q] = a; dy_[nxz + q] = b; u[p] = a; v[p] = 0.0; w[p] = b; }
        if (dyn) { const long long p = i + static_cast<long long>(nx) * (ny - 1 + static_cast<long long>(ny) * k);
                   const double a = dy_[2 * nxz + q] * g, b = dy_[3 * nxz + q] * g; dy_[2 * nxz + q] = a; dy_[3 * nxz + q] = b;
                   u[p] = a; v[p] = 0.0; w[p] = b; }
      });
    if (dz1 || dzn)
      map(ctx, nxy, [=] __device__(long long q) {  // planes: dpdxz1, dpdyz1, dpdxzn, dpdyzn
        if (dz1) { const double a = dz_[q] * g, b = dz_[nxy + q] * g; dz_[q] = a; dz_[nxy + q] = b; u[q] = a; v[q] = b; w[q] = 0.0; }
        if (dzn) { const long long p = q + nxy * (nzl - 1); const double a = dz_[2 * nxy + q] * g, b = dz_[3 * nxy + q] * g;
                   dz_[2 * nxy + q] = a; dz_[3 * nxy + q] = b; u[p] = a; v[p] = b; w[p] = 0.0; }
      });
  }
  if (x1 || xn)
    map(ctx, static_cast<long long>(ny) * nzl, [=] __device__(long long q) {
      if (x1) u[q * nx] = 0.0;
      if (xn) u[q * nx + nx - 1] = 0.0;
    });
  if (y1 || yn)
    map(ctx, static_cast<long long>(nx) * nzl, [=] __device__(long long q) {
      const long long i = q % nx, k = q / nx;
      if (y1) v[i + static_cast<long long>(nx) * ny * k] = 0.0;
      if (yn) v[i + static_cast<long long>(nx) * (ny - 1 + static_cast<long long>(ny) * k)] = 0.0;
    });
  if (z1 || zn)
    map(ctx, static_cast<long long>(nx) * ny, [=] __device__(long long q) {
      if (z1) w[q] = 0.0;
      if (zn) w[q + static_cast<long long>(nx) * ny * (nzl - 1)] = 0.0;
    });
}

// navier.f90:257-347 ; result in out: z-pencil of ph1 (nxm, nyml, nzm)
static void divergence(Ctx &ctx, SolverImpl &S, double *out, int nlock) {
  double *pp1 = B(S.w[0]), *pgy1 = B(S.w[1]), *pgz1 = B(S.w[2]), *upi2 = B(S.w[3]), *duy = B(S.w[4]), *po3 = B(S.w[5]);
  double *t1 = B(S.w[6]), *t2 = B(S.w[7]);
  const double *ta1 = B(S.ux), *tb1 = B(S.uy), *tc1 = B(S.uz);
  if (S.cs.iibm != 0) {  // :285-293: (1 - ep1) u + ep1 ubc
    if (!S.ep1.p) throw Error("solver: iibm /= 0 without x3d_solver_set_ibm_mask");
    double *a = B(S.w[11]), *b = B(S.w[12]), *c = B(S.w[13]);
    const double *ep = B(S.ep1), *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
    const double ubx = S.cs.ubcx, uby = S.cs.ubcy, ubz = S.cs.ubcz, one = 1.0;
    map(ctx, static_cast<long long>(S.n), [=] __device__(long long q) {
      const double e = ep[q];
      a[q] = (one - e) * u[q] + e * ubx; b[q] = (one - e) * v[q] + e * uby; c[q] = (one - e) * w[q] + e * ubz;
    });
    ta1 = a; tb1 = b; tc1 = c;
  }
  const bool ovl = S.nranks > 1 && S.overlap_div && S.stag[1];
  run(ctx, S.dvp[0], ta1, pp1);    // :297
  run(ctx, S.ivp[0], tb1, pgy1);   // :313
  if (!ovl) run(ctx, S.ivp[0], tc1, pgz1);   // :314   (transpose_x_to_y :316-318 is local: p_row = 1)
  const long long nxm_ = S.nxm;
  if (ovl) {
    // duy is complete before the x and y interpolations of uz have started: its transpose (:329) runs on the second stream
    // beside them, the transpose of their result (:330) follows on the same stream (one order of the group barriers on all ranks)
    ensure_aux(S);
    launch_stag_pair(ctx, 0, 1, S.ivp[1].op, S.dvp[1].op, pp1, pgy1, duy, nullptr, nxm_, S.p.ny, S.nzl, nxm_, nxm_ * S.p.ny);
    X3D_CUDA(cudaEventRecord(S.ev_fork, ctx.stream));
    X3D_CUDA(cudaStreamWaitEvent(S.aux, S.ev_fork, 0));
    on_aux(ctx, S, [&] { transpose_device(ctx, 1, duy, t1, S.id_p3, 1); });
    run(ctx, S.ivp[0], tc1, pgz1);       // :314
    run(ctx, S.ivp[1], pgz1, upi2);      // :327
    X3D_CUDA(cudaEventRecord(S.ev_fork, ctx.stream));
    X3D_CUDA(cudaStreamWaitEvent(S.aux, S.ev_fork, 0));
    on_aux(ctx, S, [&] { transpose_device(ctx, 1, upi2, t2, S.id_p3, 1); });
    X3D_CUDA(cudaEventRecord(S.ev_join, S.aux));
    X3D_CUDA(cudaStreamWaitEvent(ctx.stream, S.ev_join, 0));
    const long long n3o = static_cast<long long>(S.n3);
    if (nlock != 2 && S.stag[2] && n3o > 0) {
      launch_stag_pair(ctx, 0, 2, S.ivp[2].op, S.dvp[2].op, t1, t2, out, nullptr, nxm_ * S.nyml, S.p.nz, 1, nxm_ * S.nyml, 0);
      return;
    }
  }
  const double *duy3 = duy, *uzp3 = upi2;
  if (ovl) { duy3 = t1; uzp3 = t2; }
  else {
  if (S.stag[1]) {                     // :321-325 in one kernel: duy = interyvp(pp1) + deryvp(pgy1)
    launch_stag_pair(ctx, 0, 1, S.ivp[1].op, S.dvp[1].op, pp1, pgy1, duy, nullptr, nxm_, S.p.ny, S.nzl, nxm_, nxm_ * S.p.ny);
  } else {
  run(ctx, S.dvp[1], pgy1, duy);       // :322
  if (S.fuse_sums) {
    run(ctx, S.ivp_y_add, pp1, duy);   // :321 + :325: duy += interyvp(pp1), accumulated by the operator's store
  } else {
    run(ctx, S.ivp[1], pp1, upi2);     // :321
    const long long n2 = static_cast<long long>(S.nxm) * S.nym * S.nzl;
    map(ctx, n2, [=] __device__(long long q) { duy[q] = duy[q] + upi2[q]; });  // :325
  }
  }
  run(ctx, S.ivp[1], pgz1, upi2);      // :327
  if (S.nranks > 1) {  // :329-330, both fields between one pair of barriers
    const double *src[2] = {duy, upi2};
    double *dst[2] = {t1, t2};
    transpose_device_multi(ctx, 1, 2, src, dst, S.id_p3, 1);
    duy3 = t1; uzp3 = t2;
  }
  }
  const long long n3 = static_cast<long long>(S.n3);
  if (nlock != 2 && S.stag[2] && n3 > 0) {   // :333-339 in one kernel: out = interzvp(duy3) + derzvp(uzp3)
    launch_stag_pair(ctx, 0, 2, S.ivp[2].op, S.dvp[2].op, duy3, uzp3, out, nullptr, nxm_ * S.nyml, S.p.nz, 1, nxm_ * S.nyml, 0);
    return;
  }
  run(ctx, S.ivp[2], duy3, out);       // :333
  if (nlock != 2 && S.fuse_sums) {
    run(ctx, S.dvp_z_add, uzp3, out);  // :335 + :339: out += derzvp(uzp3)
    return;
  }
  run(ctx, S.dvp[2], uzp3, po3);       // :335
  if (nlock == 2) {                    // :339-347 (each rank subtracts its own corner value)
    const long long ref = static_cast<long long>(S.nxm) * S.nyml * (S.nzm - 1);
    double *tmp = B(S.red_out) + 8;
    if (n3 > 0) {
      map(ctx, 1, [=] __device__(long long) { tmp[0] = out[ref] + po3[ref]; });
      map(ctx, n3, [=] __device__(long long q) { out[q] = (out[q] + po3[q]) - tmp[0]; });
    }
  } else {
    map(ctx, n3, [=] __device__(long long q) { out[q] = out[q] + po3[q]; });
  }
}

// the wall-gradient capture of gradp needs px, py, pz: cor_vel is folded into the operators only without Dirichlet faces
static bool fused_cor_vel(const SolverImpl &S) {
  const auto &p = S.p;
  return S.fuse_sums && p.nclx1 != 2 && p.nclxn != 2 && p.ncly1 != 2 && p.nclyn != 2 && p.nclz1 != 2 && p.nclzn != 2;
}

// navier.f90:386-431
static void gradp(Ctx &ctx, SolverImpl &S, const double *pp3, int itr) {
  double *ppi3 = B(S.w[0]), *pgz3 = B(S.w[1]), *ppi2 = B(S.w[2]), *pgy2 = B(S.w[3]), *pgzi2 = B(S.w[4]);
  double *t1 = B(S.w[5]), *t2 = B(S.w[6]);
  const long long nxm_ = S.nxm;
  if (S.stag[2] && S.n3 > 0) {     // :404-406 in one kernel: one read of pp3
    launch_stag_pair(ctx, 1, 2, S.ipv[2].op, S.dpv[2].op, pp3, nullptr, ppi3, pgz3, nxm_ * S.nyml, S.p.nz, 1, nxm_ * S.nyml, 0);
  } else {
    run(ctx, S.ipv[2], pp3, ppi3);   // :404
    run(ctx, S.dpv[2], pp3, pgz3);   // :406
  }
  const double *pgz2 = pgz3, *pp2 = ppi3;
  const bool ovl = S.nranks > 1 && S.overlap_div && S.stag[1];
  if (ovl) {
    // ppi3 first (:411), then pgz3 (:410) on the second stream beside the y operators of ppi3
    ensure_aux(S);
    transpose_device(ctx, 2, ppi3, t2, S.id_p3, 1);
    X3D_CUDA(cudaEventRecord(S.ev_fork, ctx.stream));
    X3D_CUDA(cudaStreamWaitEvent(S.aux, S.ev_fork, 0));
    on_aux(ctx, S, [&] { transpose_device(ctx, 2, pgz3, t1, S.id_p3, 1); });
    X3D_CUDA(cudaEventRecord(S.ev_join, S.aux));
    pgz2 = t1; pp2 = t2;
  } else if (S.nranks > 1) {  // :410-411
    const double *src[2] = {pgz3, ppi3};
    double *dst[2] = {t1, t2};
    transpose_device_multi(ctx, 2, 2, src, dst, S.id_p3, 1);
    pgz2 = t1; pp2 = t2;
  }
  if (S.stag[1]) {                 // :413-415 in one kernel: one read of pp2
    launch_stag_pair(ctx, 1, 1, S.ipv[1].op, S.dpv[1].op, pp2, nullptr, ppi2, pgy2, nxm_, S.p.ny, S.nzl, nxm_, nxm_ * S.p.ny);
    if (ovl) X3D_CUDA(cudaStreamWaitEvent(ctx.stream, S.ev_join, 0));
  } else {
    run(ctx, S.ipv[1], pp2, ppi2);   // :413
    run(ctx, S.dpv[1], pp2, pgy2);   // :415
  }
  run(ctx, S.ipv[1], pgz2, pgzi2); // :417  (transpose_y_to_x :422-424 is local)
  if (fused_cor_vel(S)) {  // cor_vel (:242-244) folded in: u -= derxpv(ppi2), v -= interxpv(pgy2), w -= interxpv(pgzi2)
    run(ctx, S.dpv_x_sub, ppi2, B(S.ux));
    run(ctx, S.ipv_x_sub, pgy2, B(S.uy));
    run(ctx, S.ipv_x_sub, pgzi2, B(S.uz));
    return;
  }
  run(ctx, S.dpv[0], ppi2, B(S.px));   // :426
  run(ctx, S.ipv[0], pgy2, B(S.py));   // :428
  run(ctx, S.ipv[0], pgzi2, B(S.pz));  // :430
  // wall pressure gradients for the next pre_correc, :439-496 (the z faces keep py / pz, as the reference does)
  {
    const int nx = S.p.nx, ny = S.p.ny, nzl = S.nzl;
    const double g = S.gdt[itr - 1];
    const double *px = B(S.px), *py = B(S.py), *pz = B(S.pz);
    double *d = B(S.dpd);
    const long long nyz = static_cast<long long>(ny) * nzl, nxz = static_cast<long long>(nx) * nzl, nxy = static_cast<long long>(nx) * ny;
    double *dx_ = d, *dy_ = d + 4 * nyz, *dz_ = d + 4 * nyz + 4 * nxz;
    const bool dx1 = S.p.nclx1 == 2, dxn = S.p.nclxn == 2, dy1 = S.p.ncly1 == 2, dyn = S.p.nclyn == 2;
    const bool dz1 = S.p.nclz1 == 2 && S.z0 == 0, dzn = S.p.nclzn == 2 && S.z0 + nzl == S.p.nz;
    if (dx1 || dxn)
      map(ctx, nyz, [=] __device__(long long q) {
        if (dx1) { dx_[q] = py[q * nx] / g; dx_[nyz + q] = pz[q * nx] / g; }
        if (dxn) { dx_[2 * nyz + q] = py[q * nx + nx - 1] / g; dx_[3 * nyz + q] = pz[q * nx + nx - 1] / g; }
      });
    if (dy1 || dyn)
      map(ctx, nxz, [=] __device__(long long q) {
        const long long i = q % nx, k = q / nx;
        if (dy1) { const long long p = i + static_cast<long long>(nx) * ny * k; dy_[q] = px[p] / g; dy_[nxz + q] = pz[p] / g; }
        if (dyn) { const long long p = i + static_cast<long long>(nx) * (ny - 1 + static_cast<long long>(ny) * k); dy_[2 * nxz + q] = px[p] / g; dy_[3 * nxz + q] = pz[p] / g; }
      });
    if (dz1 || dzn)
      map(ctx, nxy, [=] __device__(long long q) {
        if (dz1) { dz_[q] = py[q] / g; dz_[nxy + q] = pz[q] / g; }
        if (dzn) { const long long p = q + nxy * (nzl - 1); dz_[2 * nxy + q] = py[p] / g; dz_[3 * nxy + q] = pz[p] / g; }
      });
  }
}

void solver_step(Ctx &ctx, int nsteps) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const long long n = static_cast<long long>(S.n);
  for (int st = 0; st < nsteps; ++st) {
    S.itime += 1;
    for (int itr = 1; itr <= S.iadvance; ++itr) {  // xcompact3d.f90:46-88
      boundary_conditions(ctx, S, itr);
      if (S.fused[1] && S.fused[2] && S.cs.iibm == 0) {
        double *rhs[3] = {B(S.w[0]), B(S.w[1]), B(S.w[2])}, *extra[3];
        if (!momentum_rhs_fused(ctx, S, rhs, extra, itr)) intt3_fused(ctx, S, itr, rhs, extra);
      } else {
        momentum_rhs(ctx, S, B(S.dux[0]), B(S.duy[0]), B(S.duz[0]));
        intt3(ctx, S, itr);
      }
      pre_correc(ctx, S, itr);
      divergence(ctx, S, B(S.pp3), 1);
      poisson_solve_device(ctx, B(S.pp3));
      gradp(ctx, S, B(S.pp3), itr);
      if (!fused_cor_vel(S)) {
        double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
        const double *px = B(S.px), *py = B(S.py), *pz = B(S.pz);
        map(ctx, n, [=] __device__(long long q) { u[q] = u[q] - px[q]; v[q] = v[q] - py[q]; w[q] = w[q] - pz[q]; });  // cor_vel
      }
    }
  }
}

// Case-TGV.f90:246-380 ; out5 = eek, eps, eps2, enst, DIV U max
void solver_diagnostics_tgv(Ctx &ctx, double *out5) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const int nx = S.p.nx, ny = S.p.ny, nz = S.p.nz, nzl = S.nzl, z0 = S.z0;
  const double *u = B(S.ux), *v = B(S.uy), *w = B(S.uz);
  double *ta = B(S.w[0]), *tb = B(S.w[1]), *tc = B(S.w[2]), *td = B(S.w[3]), *te = B(S.w[4]), *tf = B(S.w[5]);
  double *tg = B(S.w[6]), *th = B(S.w[7]), *ti = B(S.w[8]);
  double *s1 = B(S.w[9]), *s2 = B(S.w[10]), *s3 = B(S.w[11]), *s4 = B(S.w[12]), *s5 = B(S.w[13]), *s6 = B(S.w[14]);
  const int xs1 = S.p.nclx1 == 1 ? nx - 1 : nx, xs2 = S.p.ncly1 == 1 ? ny - 1 : ny, xs3 = S.p.nclz1 == 1 ? nz - 1 : nz;
  const double ncell = static_cast<double>(S.p.nclx1 == 1 ? S.nxm : nx) * (S.p.ncly1 == 1 ? S.nym : ny) * (S.p.nclz1 == 1 ? S.nzm : nz);
  const double xnu = S.xnu;
  const long long n = static_cast<long long>(S.n);
  const int nb = gridn(ctx, n) > 4096 ? 4096 : gridn(ctx, n);
  double *partial = B(S.red_partial), *dout = B(S.red_out);
  auto inside = [=] __device__(long long q) {
    const int i = static_cast<int>(q % nx), j = static_cast<int>((q / nx) % ny), k = static_cast<int>(q / (static_cast<long long>(nx) * ny)) + z0;
    return i < xs1 && j < xs2 && k < xs3;
  };
  auto zder = [&](const PreOp &op, const double *f, double *scratch_in, double *scratch_out, double *back) -> const double * {
    // derivative along z of an x-pencil field: y->z, operator, z->y (Case-TGV.f90:250-281)
    const double *f3 = TR(ctx, S, 1, f, scratch_in, S.id_v);
    double *o3 = (S.nranks == 1) ? back : scratch_out;
    run(ctx, op, f3, o3);
    return TR(ctx, S, 2, o3, back, S.id_v);
  };
  run(ctx, S.d1[0][0], u, ta); run(ctx, S.d1[0][1], v, tb); run(ctx, S.d1[0][1], w, tc);  // :261-263
  run(ctx, S.d1[1][1], u, td); run(ctx, S.d1[1][0], v, te); run(ctx, S.d1[1][1], w, tf);  // :265-267
  zder(S.d1[2][1], u, s1, s2, tg); zder(S.d1[2][1], v, s1, s2, th); zder(S.d1[2][0], w, s1, s2, ti);  // :269-271
  (void)s3; (void)s4; (void)s5; (void)s6; (void)nzl;
  k_reduce_partial<3><<<nb, 256, 0, ctx.stream>>>(n, [=] __device__(long long q, double *acc) {
    if (!inside(q)) return;
    const double a = tf[q] - th[q], b = tg[q] - tc[q], c = tb[q] - td[q];
    acc[0] += 0.5 * (a * a + b * b + c * c);                                                      // enstrophy :291-293
    const double e1 = 2.0 * ta[q], e2 = 2.0 * te[q], e3 = 2.0 * ti[q], e4 = td[q] + tb[q], e5 = tg[q] + tc[q], e6 = th[q] + tf[q];
    acc[1] += 0.5 * xnu * (e1 * e1 + e2 * e2 + e3 * e3 + 2.0 * e4 * e4 + 2.0 * e5 * e5 + 2.0 * e6 * e6);  // eps :305-308
    acc[2] += 0.5 * (u[q] * u[q] + v[q] * v[q] + w[q] * w[q]);                                   // eek :323
  }, partial);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  k_reduce_final<3><<<1, 256, 0, ctx.stream>>>(nb, partial, dout);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  run(ctx, S.d2[0][0], u, ta); run(ctx, S.d2[0][1], v, tb); run(ctx, S.d2[0][1], w, tc);  // :332-334
  run(ctx, S.d2[1][1], u, td); run(ctx, S.d2[1][0], v, te); run(ctx, S.d2[1][1], w, tf);  // :336-338
  zder(S.d2[2][1], u, s1, s2, tg); zder(S.d2[2][1], v, s1, s2, th); zder(S.d2[2][0], w, s1, s2, ti);  // :340-342
  k_reduce_partial<1><<<nb, 256, 0, ctx.stream>>>(n, [=] __device__(long long q, double *acc) {
    if (!inside(q)) return;
    acc[0] += (-xnu) * (u[q] * (ta[q] + td[q] + tg[q]) + v[q] * (tb[q] + te[q] + th[q]) + w[q] * (tc[q] + tf[q] + ti[q]));  // :362-365
  }, partial);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  k_reduce_final<1><<<1, 256, 0, ctx.stream>>>(nb, partial, dout + 3);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  allreduce(ctx, dout, 4, false);  // MPI_ALLREDUCE(SUM), Case-TGV.f90:297,315,327,369
  // DIV U max of the current field (divergence nlock=2, navier.f90:341-361)
  double *dv = B(S.w[9]);
  divergence(ctx, S, dv, 2);
  const long long n3 = static_cast<long long>(S.n3);
  const int nb3 = std::max(1, gridn(ctx, n3) > 4096 ? 4096 : gridn(ctx, n3));
  k_max_partial<<<nb3, 256, 0, ctx.stream>>>(n3, dv, partial);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  k_max_final<<<1, 256, 0, ctx.stream>>>(nb3, partial, dout + 4);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
  allreduce(ctx, dout + 4, 1, true);
  X3D_CUDA(cudaMemcpyAsync(S.h_red, dout, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  out5[0] = S.h_red[2] / ncell;  // eek
  out5[1] = S.h_red[1] / ncell;  // eps
  out5[2] = S.h_red[3] / ncell;  // eps2
  out5[3] = S.h_red[0] / ncell;  // enstrophy
  out5[4] = S.h_red[4];
}

void solver_divergence(Ctx &ctx, double *divmax, double *divmean) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  double *dv = B(S.w[9]);
  divergence(ctx, S, dv, 2);
  const long long n3 = static_cast<long long>(S.n3);
  const int nb3 = std::max(1, gridn(ctx, n3) > 4096 ? 4096 : gridn(ctx, n3));
  double *partial = B(S.red_partial), *dout = B(S.red_out);
  k_max_partial<<<nb3, 256, 0, ctx.stream>>>(n3, dv, partial);
  k_max_final<<<1, 256, 0, ctx.stream>>>(nb3, partial, dout);
  k_reduce_partial<1><<<nb3, 256, 0, ctx.stream>>>(n3, [=] __device__(long long q, double *acc) { acc[0] += fabs(dv[q]); }, partial);
  k_reduce_final<1><<<1, 256, 0, ctx.stream>>>(nb3, partial, dout + 1);
  X3D_CUDA(cudaGetLastError()); ctx.launches += 4;
  allreduce(ctx, dout, 1, true);
  allreduce(ctx, dout + 1, 1, false);
  X3D_CUDA(cudaMemcpyAsync(S.h_red, dout, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  if (divmax) *divmax = S.h_red[0];
  // navier.f90:359-367: local mean, summed over ranks, divided by nproc when printed
  const double ntot = static_cast<double>(S.nxm) * S.nym * S.nzm;
  if (divmean) *divmean = S.h_red[1] / ntot;
}

// x-pencil velocity of this rank: (nx, ny, nzl) -- host or device pointers
void solver_set_velocity(Ctx &ctx, const double *ux, const double *uy, const double *uz) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t bytes = S.n * sizeof(double);
  X3D_CUDA(cudaMemcpyAsync(S.ux.p, ux, bytes, cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(S.uy.p, uy, bytes, cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(S.uz.p, uz, bytes, cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}
void solver_get_velocity(Ctx &ctx, double *ux, double *uy, double *uz) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t bytes = S.n * sizeof(double);
  X3D_CUDA(cudaMemcpyAsync(ux, S.ux.p, bytes, cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(uy, S.uy.p, bytes, cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(uz, S.uz.p, bytes, cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}
// apply_spatial_filter, src/tools.f90:600-675 (called from the time loop when ifilter /= 0, src/xcompact3d.f90), on the
// solver's velocity; af = the filter parameter of set_filter_coefficients (src/filters.f90:62-219, `filter(C_filter)`).
// ifilter: 1 all directions, 2 x and z, 3 y only.
void solver_apply_spatial_filter(Ctx &ctx, int ifilter, double af) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (ifilter < 1 || ifilter > 3) throw Error("x3d_solver_apply_spatial_filter: ifilter 1, 2 or 3");
  const auto &p = S.p;
  if (S.fil_af != af) {
    const int nn[3] = {p.nx, p.ny, p.nz};
    const int dxy[3] = {p.nx, p.ny, S.nzl}, dzp[3] = {p.nx, S.nyl, p.nz};
    for (int a = 0; a < 3; ++a) {
      make_filter_axis(nn[a], S.A[a].ncl1, S.A[a].ncln, af, ctx.fc[a], S.fil_lu[a][0], S.fil_lu[a][1]);
      ctx.have_fc[a] = true;
      for (int np = 0; np < 2; ++np) prep(ctx, S.fil[a][np], FIL, a, S.A[a], np, S.fil_lu[a][np], a == 2 ? dzp : dxy);
    }
    S.fil_af = af;
  }
  const long long n = static_cast<long long>(S.n);
  double *vel[3] = {B(S.ux), B(S.uy), B(S.uz)};
  double *f1[3] = {B(S.w[0]), B(S.w[1]), B(S.w[2])}, *f2[3] = {B(S.w[3]), B(S.w[4]), B(S.w[5])};
  const int iibm = S.cs.iibm;
  const double ubc[3] = {S.cs.ubcx, S.cs.ubcy, S.cs.ubcz};
  auto fil = [&](int axis, int c, double *in, double *out) {   // the component along the axis is odd: npaire = 0 (:624-626,641-643,658-660)
    const PreOp &P = S.fil[axis][c == axis ? 0 : 1];
    if (iibm == 2) lagpol_device(ctx, axis, in, P.call.dims_in[0], P.call.dims_in[1], P.call.dims_in[2]);
    else if (iibm == 3) cubspl_device(ctx, axis, in, P.call.dims_in[0], P.call.dims_in[1], P.call.dims_in[2], ubc[c]);
    run(ctx, P, in, out);
  };
  double *cur[3] = {vel[0], vel[1], vel[2]};
  if (ifilter == 1 || ifilter == 2) { for (int c = 0; c < 3; ++c) { fil(0, c, cur[c], f1[c]); cur[c] = f1[c]; } }
  if (ifilter == 1 || ifilter == 3) { for (int c = 0; c < 3; ++c) { fil(1, c, cur[c], f2[c]); cur[c] = f2[c]; } }
  if (ifilter == 1 || ifilter == 2) {
    double *other[3];
    for (int c = 0; c < 3; ++c) other[c] = (cur[c] == f2[c]) ? f1[c] : f2[c];
    if (S.nranks == 1) {
      for (int c = 0; c < 3; ++c) { fil(2, c, cur[c], other[c]); cur[c] = other[c]; }
    } else {  // transpose_y_to_z, filz, transpose_z_to_y (:649-670)
      double *z1[3] = {B(S.w[6]), B(S.w[7]), B(S.w[8])}, *z2[3] = {B(S.w[9]), B(S.w[10]), B(S.w[11])};
      transpose_device_multi(ctx, 1, 3, cur, z1, S.id_v, 1);
      for (int c = 0; c < 3; ++c) fil(2, c, z1[c], z2[c]);
      transpose_device_multi(ctx, 2, 3, z2, other, S.id_v, 1);
      for (int c = 0; c < 3; ++c) cur[c] = other[c];
    }
  }
  for (int c = 0; c < 3; ++c)
    if (cur[c] != vel[c]) X3D_CUDA(cudaMemcpyAsync(vel[c], cur[c], n * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
}

// x3d_solver_set_case: channel forcing, cylinder inflow / outflow, immersed boundary
void solver_set_case(Ctx &ctx, const x3d_case_params &c) {
  SolverImpl &S = SOL(ctx);
  if (c.iibm != 0 && c.iibm != 2 && c.iibm != 3) throw Error("x3d_solver_set_case: iibm 0, 2 (lagpol) and 3 (cubspl) are implemented");
  if (c.iibm != 0 && S.nranks > 1) throw Error("x3d_solver_set_case: the immersed-boundary step runs on one rank");
  if (c.cpg && S.p.itype != 3) throw Error("x3d_solver_set_case: cpg is a channel option");
  S.cs = c;
  S.xnu = 1.0 / S.p.re;
  S.fcpg = 0.0;
  if (c.cpg) {  // src/parameters.f90:303-311
    const double re_cent = pow(S.p.re / 0.116, 1.0 / 0.88);
    S.xnu = 1.0 / re_cent;
    S.fcpg = 2.0 / S.p.yly * ((S.p.re / re_cent) * (S.p.re / re_cent));
  }
  ctx.iibm = 0;   // the solver runs the pre-pass itself (the operator entry points would run it a second time)
}
void solver_set_ibm_mask(Ctx &ctx, const double *ep1) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  S.ep1.reserve(std::max<size_t>(S.n, 1) * sizeof(double));
  X3D_CUDA(cudaMemcpyAsync(S.ep1.p, ep1, S.n * sizeof(double), cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}
void solver_set_inflow_noise(Ctx &ctx, const double *bxo, const double *byo, const double *bzo) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t nyz = static_cast<size_t>(S.p.ny) * S.nzl;
  S.bnoise.reserve(3 * std::max<size_t>(nyz, 1) * sizeof(double));
  X3D_CUDA(cudaMemsetAsync(S.bnoise.p, 0, S.bnoise.bytes, ctx.stream));
  const double *src[3] = {bxo, byo, bzo};
  for (int q = 0; q < 3; ++q)
    if (src[q]) X3D_CUDA(cudaMemcpyAsync(B(S.bnoise) + q * nyz, src[q], nyz * sizeof(double), cudaMemcpyDefault, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}
void solver_wall_velocity_x(Ctx &ctx, const double *const in6[6], double *const out6[6]) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (!S.bwx.p) throw Error("solver: the x faces are not Dirichlet faces");
  const size_t nyz = static_cast<size_t>(S.p.ny) * S.nzl;
  for (int q = 0; q < 6; ++q) {
    if (in6 && in6[q]) X3D_CUDA(cudaMemcpyAsync(B(S.bwx) + q * nyz, in6[q], nyz * sizeof(double), cudaMemcpyDefault, ctx.stream));
    if (out6 && out6[q]) X3D_CUDA(cudaMemcpyAsync(out6[q], B(S.bwx) + q * nyz, nyz * sizeof(double), cudaMemcpyDefault, ctx.stream));
  }
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}
// init_cyl with iin = 0, Case-Cylinder-wake.f90:205-279: uniform stream u1
void solver_init_cyl(Ctx &ctx) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  double *ux = B(S.ux), *uy = B(S.uy), *uz = B(S.uz);
  const double u1 = S.cs.u1;
  map(ctx, static_cast<long long>(S.n), [=] __device__(long long q) { ux[q] = 0.0 + u1; uy[q] = 0.0; uz[q] = 0.0; });
  for (int q = 0; q < S.ntime; ++q)
    for (DevBuf *b : {&S.dux[q], &S.duy[q], &S.duz[q]}) X3D_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, ctx.stream));
  for (DevBuf *b : {&S.px, &S.py, &S.pz, &S.pp3, &S.dpd}) X3D_CUDA(cudaMemsetAsync(b->p, 0, b->bytes, ctx.stream));
  S.itime = 0;
}

// One job = copy a host velocity field in, advance it nsteps, copy the result out; asynchronous with respect to the
// host.  Consecutive jobs are independent (each starts from its own host input), so the copies of neighbouring jobs
// overlap the kernels of the current one: H2D on s_in, kernels on ctx.stream, D2H on s_out, three device velocity
// sets in rotation.  Host arrays should be page-locked, otherwise the runtime stages the copies and they serialise.
// Intended for self-starting time schemes (RK3, Euler): the Adams-Bashforth history is not part of a job.
void solver_advance_host(Ctx &ctx, const double *const in[3], double *const out[3], int nsteps) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t bytes = S.n * sizeof(double);
  if (!S.s_in) {
    X3D_CUDA(cudaStreamCreateWithFlags(&S.s_in, cudaStreamNonBlocking));
    X3D_CUDA(cudaStreamCreateWithFlags(&S.s_out, cudaStreamNonBlocking));
    for (auto &v : S.spare) {
      for (auto &b : v.b) b.reserve(S.ux.bytes);
      X3D_CUDA(cudaEventCreateWithFlags(&v.h2d_done, cudaEventDisableTiming));
      X3D_CUDA(cudaEventCreateWithFlags(&v.d2h_done, cudaEventDisableTiming));
    }
    X3D_CUDA(cudaEventCreateWithFlags(&S.cur_d2h, cudaEventDisableTiming));
    X3D_CUDA(cudaEventCreateWithFlags(&S.ev_step, cudaEventDisableTiming));
  }
  SolverImpl::VelSet &F = S.spare[S.next_spare];
  S.next_spare ^= 1;
  if (F.pending) X3D_CUDA(cudaStreamWaitEvent(S.s_in, F.d2h_done, 0));     // the set's previous result has left it
  // a job that reads a host array an earlier job still writes waits for that copy
  for (auto &v : S.spare)
    if (v.pending && v.host_out == in[0]) X3D_CUDA(cudaStreamWaitEvent(S.s_in, v.d2h_done, 0));
  if (S.cur_pending && S.cur_host_out == in[0]) X3D_CUDA(cudaStreamWaitEvent(S.s_in, S.cur_d2h, 0));
  for (int c = 0; c < 3; ++c) X3D_CUDA(cudaMemcpyAsync(F.b[c].p, in[c], bytes, cudaMemcpyDefault, S.s_in));
  X3D_CUDA(cudaEventRecord(F.h2d_done, S.s_in));
  X3D_CUDA(cudaStreamWaitEvent(ctx.stream, F.h2d_done, 0));
  // rotate: the freshly filled set becomes the velocity, the old velocity (its D2H possibly in flight) becomes a spare
  DevBuf *cur[3] = {&S.ux, &S.uy, &S.uz};
  for (int c = 0; c < 3; ++c) { std::swap(cur[c]->p, F.b[c].p); std::swap(cur[c]->bytes, F.b[c].bytes); }
  std::swap(S.cur_d2h, F.d2h_done);
  std::swap(S.cur_pending, F.pending);
  std::swap(S.cur_host_out, F.host_out);
  solver_step(ctx, nsteps);
  X3D_CUDA(cudaEventRecord(S.ev_step, ctx.stream));
  X3D_CUDA(cudaStreamWaitEvent(S.s_out, S.ev_step, 0));
  for (int c = 0; c < 3; ++c) X3D_CUDA(cudaMemcpyAsync(out[c], cur[c]->p, bytes, cudaMemcpyDefault, S.s_out));
  X3D_CUDA(cudaEventRecord(S.cur_d2h, S.s_out));
  S.cur_pending = true;
  S.cur_host_out = out[0];
}
void solver_host_sync(Ctx &ctx) {
  SolverImpl &S = SOL(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (S.s_in) X3D_CUDA(cudaStreamSynchronize(S.s_in));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  if (S.s_out) X3D_CUDA(cudaStreamSynchronize(S.s_out));
}
void solver_local_shape(Ctx &ctx, int *d3, int *z0) {
  SolverImpl &S = SOL(ctx);
  d3[0] = S.p.nx; d3[1] = S.p.ny; d3[2] = S.nzl;
  *z0 = S.z0;
}

}  // namespace x3d
