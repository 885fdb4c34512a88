// x3d_capi.cu -- extern "C" entry points declared in include/x3d_b200.h.
#include <cstdlib>
#include <cstring>
#include "x3d_ctx.cuh"

#include "x3d_schemes.cuh"

using namespace x3d;

namespace x3d {
void poisson_init(Ctx &ctx, const x3d_poisson_params &p);
void poisson_solve_device(Ctx &ctx, double *d_rhs);
void poisson_dims(Ctx &ctx, int d[3]);
void solver_init(Ctx &ctx, const x3d_solver_params &p);
void set_ibm_geometry(Ctx &ctx, int axis, int nobjmax, int npif, int izap, int na, int nb, const int *nobj, const double *xi,
                      const double *xf, const int *nipif, const int *nfpif, const double *coords, int ncoords, double d, double len);
void lagpol(Ctx &ctx, int axis, double *u, int nx, int ny, int nz);
void cubspl(Ctx &ctx, int axis, double *u, int nx, int ny, int nz, double lind);
void set_ibm_analytic(Ctx &ctx, int axis, const double *ana_i, const double *ana_f);
void solver_init_tgv(Ctx &ctx);
void solver_init_channel(Ctx &ctx);
void solver_step(Ctx &ctx, int nsteps);
void solver_diagnostics_tgv(Ctx &ctx, double *out5);
void solver_divergence(Ctx &ctx, double *divmax, double *divmean);
void solver_set_velocity(Ctx &ctx, const double *ux, const double *uy, const double *uz);
void solver_get_velocity(Ctx &ctx, double *ux, double *uy, double *uz);
void solver_local_shape(Ctx &ctx, int *d3, int *z0);
void solver_advance_host(Ctx &ctx, const double *const in[3], double *const out[3], int nsteps);
void solver_host_sync(Ctx &ctx);
void decomp_check(Ctx &ctx);
void solver_apply_spatial_filter(Ctx &ctx, int ifilter, double af);
void solver_set_case(Ctx &ctx, const x3d_case_params &c);
void solver_set_ibm_mask(Ctx &ctx, const double *ep1);
void solver_set_inflow_noise(Ctx &ctx, const double *bxo, const double *byo, const double *bzo);
void solver_wall_velocity_x(Ctx &ctx, const double *const in6[6], double *const out6[6]);
void solver_init_cyl(Ctx &ctx);
long long transpose_selftest(Ctx &ctx, int which, int id, int elem, int mode);
void decomp_stats(Ctx &ctx, unsigned long long *remote_bytes, unsigned long long *fields);
void decomp_init(Ctx &ctx, int nx, int ny, int nz, int p_row, int p_col, int rank, int nranks, const void *nccl_id);
int decomp_info_init(Ctx &ctx, int nx, int ny, int nz);
void decomp_info_get(Ctx &ctx, int id, x3d_decomp_info *out);
void transpose(Ctx &ctx, int which, const double *src, double *dst, int id, int elem);
void transpose_pack(Ctx &ctx, int which, const double *d_src, double *d_packed, int id, int elem);
void transpose_unpack(Ctx &ctx, int which, const double *d_packed, double *d_dst, int id, int elem);
void nccl_unique_id(void *out128);
void decomp_compute(int nx, int ny, int nz, int p_row, int p_col, int rank, x3d_decomp_info *out);
void transpose_plan_host(int nx, int ny, int nz, int p_row, int p_col, int rank, int which, int *npeers, int *peer_ranks,
                         long long *scount, long long *sdispl, long long *rcount, long long *rdispl, int *send_dims,
                         int *recv_dims);
}

struct x3d_ctx {
  Ctx c;
};

static thread_local std::string g_last_error;

// every entry point selects the context's device; the caller's current device is put back on the way out
struct DeviceRestore {
  int dev = -1;
  DeviceRestore() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
  ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

template <class F>
static int guard(F &&f) {
  DeviceRestore keep;
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    return 1;
  } catch (...) {
    g_last_error = "unknown error";
    return 2;
  }
}

extern "C" {

const char *x3d_last_error(void) { return g_last_error.c_str(); }
int x3d_version(void) { return 100; }

int x3d_create(x3d_ctx **out, int device) {
  return guard([&] {
    if (!out) throw Error("x3d_create: null output");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(std::string("x3d_create: no CUDA device (") + cudaGetErrorString(e) +
                  "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) throw Error("x3d_create: bad device index");
    X3D_CUDA(cudaSetDevice(device));
    auto *h = new x3d_ctx();
    h->c.device = device;
    cudaDeviceProp prop{};
    X3D_CUDA(cudaGetDeviceProperties(&prop, device));
    h->c.sm_count = prop.multiProcessorCount;
    X3D_CUDA(cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking));
    if (const char *e = getenv("X3D_STRIDED_VARIANT")) h->c.strided_variant = atoi(e);
    if (const char *e = getenv("X3D_CONTIG_VARIANT")) h->c.contig_variant = atoi(e);
    if (const char *e = getenv("X3D_CONTIG_COMPRESS")) h->c.contig_compress = atoi(e) != 0;
    *out = h;
  });
}
int x3d_destroy(x3d_ctx *ctx) {
  return guard([&] {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    delete ctx;
  });
}
long long x3d_launch_count(const x3d_ctx *ctx) { return ctx ? ctx->c.launches : -1; }
int x3d_sync(x3d_ctx *ctx) {
  return guard([&] {
    X3D_CUDA(cudaSetDevice(ctx->c.device));
    X3D_CUDA(cudaStreamSynchronize(ctx->c.stream));
    decomp_check(ctx->c);
  });
}
unsigned long long x3d_stream(x3d_ctx *ctx) { return reinterpret_cast<unsigned long long>(ctx->c.stream); }

int x3d_set_deriv_coeffs(x3d_ctx *ctx, int axis, const x3d_deriv_coeffs *c) {
  return guard([&] {
    if (axis < 0 || axis > 2 || !c) throw Error("x3d_set_deriv_coeffs: bad argument");
    ctx->c.dc[axis] = *c;
    ctx->c.have_dc[axis] = true;
  });
}
int x3d_set_filter_coeffs(x3d_ctx *ctx, int axis, const x3d_filter_coeffs *c) {
  return guard([&] {
    if (axis < 0 || axis > 2 || !c) throw Error("x3d_set_filter_coeffs: bad argument");
    ctx->c.fc[axis] = *c;
    ctx->c.have_fc[axis] = true;
  });
}
int x3d_set_stretching(x3d_ctx *ctx, int ny, const double *yp, const double *ypi, const double *ppy, const double *pp2y,
                       const double *pp4y, const double *ppyi, const double *pp2yi, const double *pp4yi) {
  return guard([&] {
    if (ny < 1 || !yp || !ypi || !ppy || !pp2y || !pp4y || !ppyi || !pp2yi || !pp4yi) throw Error("x3d_set_stretching: bad argument");
    Ctx &c = ctx->c;
    c.st_yp.assign(yp, yp + ny); c.st_ypi.assign(ypi, ypi + ny);
    c.st_ppy.assign(ppy, ppy + ny); c.st_pp2y.assign(pp2y, pp2y + ny); c.st_pp4y.assign(pp4y, pp4y + ny);
    c.st_ppyi.assign(ppyi, ppyi + ny); c.st_pp2yi.assign(pp2yi, pp2yi + ny); c.st_pp4yi.assign(pp4yi, pp4yi + ny);
  });
}
int x3d_set_ibm_geometry(x3d_ctx *ctx, int axis, int nobjmax, int npif, int izap, int na, int nb, const int *nobj, const double *xi,
                         const double *xf, const int *nipif, const int *nfpif, const double *coords, int ncoords, double d, double len) {
  return guard([&] { set_ibm_geometry(ctx->c, axis, nobjmax, npif, izap, na, nb, nobj, xi, xf, nipif, nfpif, coords, ncoords, d, len); });
}
int x3d_lagpolx(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz) { return guard([&] { lagpol(ctx->c, 0, u, *nx, *ny, *nz); }); }
int x3d_lagpoly(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz) { return guard([&] { lagpol(ctx->c, 1, u, *nx, *ny, *nz); }); }
int x3d_lagpolz(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz) { return guard([&] { lagpol(ctx->c, 2, u, *nx, *ny, *nz); }); }
int x3d_set_ibm_analytic(x3d_ctx *ctx, int axis, const double *ana_i, const double *ana_f) {
  return guard([&] { set_ibm_analytic(ctx->c, axis, ana_i, ana_f); });
}
int x3d_cubsplx(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz, const double *lind) { return guard([&] { cubspl(ctx->c, 0, u, *nx, *ny, *nz, *lind); }); }
int x3d_cubsply(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz, const double *lind) { return guard([&] { cubspl(ctx->c, 1, u, *nx, *ny, *nz, *lind); }); }
int x3d_cubsplz(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz, const double *lind) { return guard([&] { cubspl(ctx->c, 2, u, *nx, *ny, *nz, *lind); }); }
int x3d_set_flags(x3d_ctx *ctx, int iibm, int istret, int iimplicit, int nclx, int ncly, int nclz) {
  return guard([&] {
    ctx->c.iibm = iibm; ctx->c.istret = istret; ctx->c.iimplicit = iimplicit;
    ctx->c.ncl[0] = nclx != 0; ctx->c.ncl[1] = ncly != 0; ctx->c.ncl[2] = nclz != 0;
  });
}

// ---- collocated operators --------------------------------------------------------
static int colloc(x3d_ctx *ctx, Kind kind, int axis, int ncl1, int ncln, double *t, const double *u, const double *f,
                  const double *s, const double *w, const double *pp, int nx, int ny, int nz, int npaire, const double *lind) {
  return guard([&] {
    if (!ctx) throw Error("null context");
    OpCall call{};
    call.lind = lind ? *lind : 0.0;
    call.kind = kind; call.axis = axis; call.ncl1 = ncl1; call.ncln = ncln;
    call.periodic = (ncl1 == 0 && ncln == 0);
    call.npaire = npaire;
    call.dims_in[0] = nx; call.dims_in[1] = ny; call.dims_in[2] = nz;
    call.n = call.nm = call.dims_in[axis];
    call.f = f; call.s = s; call.w = w;
    // dery_*: ty = ty*ppy when istret /= 0 (src/derive.f90:409-417)
    call.post = (kind == D1 && axis == 1 && ctx->c.istret != 0) ? pp : nullptr;
    // deryy_*: return after the RHS when iimplicit >= 1 (src/derive.f90:2166)
    call.rhs_only = (kind == D2 && axis == 1 && ctx->c.iimplicit >= 1);
    run_op(ctx->c, call, u, t);
  });
}

#define X3D_DEF_X(name, kind, axis, a, b)                                                                     \
  int x3d_##name(x3d_ctx *ctx, double *tx, const double *ux, double *rx, double *sx, const double *ffx,       \
                 const double *fsx, const double *fwx, const int *nx, const int *ny, const int *nz,           \
                 const int *npaire, const double *lind) {                                                     \
    (void)rx; (void)sx;                                                                                       \
    return colloc(ctx, kind, axis, a, b, tx, ux, ffx, fsx, fwx, nullptr, *nx, *ny, *nz, *npaire, lind);       \
  }
#define X3D_DEF_Y(name, kind, axis, a, b)                                                                     \
  int x3d_##name(x3d_ctx *ctx, double *ty, const double *uy, double *ry, double *sy, const double *ffy,       \
                 const double *fsy, const double *fwy, const double *ppy, const int *nx, const int *ny,       \
                 const int *nz, const int *npaire, const double *lind) {                                      \
    (void)ry; (void)sy;                                                                                       \
    return colloc(ctx, kind, axis, a, b, ty, uy, ffy, fsy, fwy, ppy, *nx, *ny, *nz, *npaire, lind);           \
  }
#define X3D_FIVE(M, stem, kind, axis) \
  M(stem##_00, kind, axis, 0, 0) M(stem##_11, kind, axis, 1, 1) M(stem##_12, kind, axis, 1, 2) \
  M(stem##_21, kind, axis, 2, 1) M(stem##_22, kind, axis, 2, 2)

X3D_FIVE(X3D_DEF_X, derx, D1, 0)
X3D_FIVE(X3D_DEF_Y, dery, D1, 1)
X3D_FIVE(X3D_DEF_X, derz, D1, 2)
X3D_FIVE(X3D_DEF_X, derxx, D2, 0)
X3D_FIVE(X3D_DEF_X, deryy, D2, 1)
X3D_FIVE(X3D_DEF_X, derzz, D2, 2)
X3D_FIVE(X3D_DEF_X, filx, FIL, 0)
X3D_FIVE(X3D_DEF_X, fily, FIL, 1)
X3D_FIVE(X3D_DEF_X, filz, FIL, 2)

// ---- staggered operators ------------------------------------------------------------
static int stag(x3d_ctx *ctx, Kind kind, int axis, double *t, const double *u, const double *f, const double *s,
                const double *w, const double *pp, int d0, int d1, int d2, int n, int nm, int npaire) {
  return guard([&] {
    if (!ctx) throw Error("null context");
    OpCall call{};
    call.kind = kind; call.axis = axis; call.ncl1 = call.ncln = -1;
    call.periodic = ctx->c.ncl[axis];
    call.npaire = npaire;
    call.n = n; call.nm = nm;
    call.dims_in[0] = d0; call.dims_in[1] = d1; call.dims_in[2] = d2;
    call.f = f; call.s = s; call.w = w;
    call.post = (ctx->c.istret != 0) ? pp : nullptr;  // deryvp*ppyi, derypv*ppy (derive.f90:4572,4905)
    run_op(ctx->c, call, u, t);
  });
}

int x3d_derxvp(x3d_ctx *ctx, double *tx, const double *ux, double *, double *, const double *cfx6, const double *csx6,
               const double *cwx6, const int *nx, const int *nxm, const int *ny, const int *nz, const int *npaire) {
  return stag(ctx, DVP, 0, tx, ux, cfx6, csx6, cwx6, nullptr, *nx, *ny, *nz, *nx, *nxm, *npaire);
}
int x3d_interxvp(x3d_ctx *ctx, double *tx, const double *ux, double *, double *, const double *cifx6,
                 const double *cisx6, const double *ciwx6, const int *nx, const int *nxm, const int *ny, const int *nz,
                 const int *npaire) {
  return stag(ctx, IVP, 0, tx, ux, cifx6, cisx6, ciwx6, nullptr, *nx, *ny, *nz, *nx, *nxm, *npaire);
}
// pv operators: the periodic branch sweeps with the second LU triple (cfx6..), the other with the first
int x3d_derxpv(x3d_ctx *ctx, double *tx, const double *ux, double *, double *, const double *cfi6, const double *csi6,
               const double *cwi6, const double *cfx6, const double *csx6, const double *cwx6, const int *nxm,
               const int *nx, const int *ny, const int *nz, const int *npaire) {
  const bool per = ctx && ctx->c.ncl[0];
  return stag(ctx, DPV, 0, tx, ux, per ? cfx6 : cfi6, per ? csx6 : csi6, per ? cwx6 : cwi6, nullptr, *nxm, *ny, *nz, *nx,
              *nxm, *npaire);
}
int x3d_interxpv(x3d_ctx *ctx, double *tx, const double *ux, double *, double *, const double *cifi6,
                 const double *cisi6, const double *ciwi6, const double *cifx6, const double *cisx6,
                 const double *ciwx6, const int *nxm, const int *nx, const int *ny, const int *nz, const int *npaire) {
  const bool per = ctx && ctx->c.ncl[0];
  return stag(ctx, IPV, 0, tx, ux, per ? cifx6 : cifi6, per ? cisx6 : cisi6, per ? ciwx6 : ciwi6, nullptr, *nxm, *ny, *nz,
              *nx, *nxm, *npaire);
}
int x3d_interyvp(x3d_ctx *ctx, double *ty, const double *uy, double *, double *, const double *cify6,
                 const double *cisy6, const double *ciwy6, const int *nx, const int *ny, const int *nym, const int *nz,
                 const int *npaire) {
  return stag(ctx, IVP, 1, ty, uy, cify6, cisy6, ciwy6, nullptr, *nx, *ny, *nz, *ny, *nym, *npaire);
}
int x3d_deryvp(x3d_ctx *ctx, double *ty, const double *uy, double *, double *, const double *cfy6, const double *csy6,
               const double *cwy6, const double *ppyi, const int *nx, const int *ny, const int *nym, const int *nz,
               const int *npaire) {
  return stag(ctx, DVP, 1, ty, uy, cfy6, csy6, cwy6, ppyi, *nx, *ny, *nz, *ny, *nym, *npaire);
}
int x3d_interypv(x3d_ctx *ctx, double *ty, const double *uy, double *, double *, const double *cifi6y,
                 const double *cisi6y, const double *ciwi6y, const double *cify6, const double *cisy6,
                 const double *ciwy6, const int *nx, const int *nym, const int *ny, const int *nz, const int *npaire) {
  const bool per = ctx && ctx->c.ncl[1];
  return stag(ctx, IPV, 1, ty, uy, per ? cify6 : cifi6y, per ? cisy6 : cisi6y, per ? ciwy6 : ciwi6y, nullptr, *nx, *nym,
              *nz, *ny, *nym, *npaire);
}
int x3d_derypv(x3d_ctx *ctx, double *ty, const double *uy, double *, double *, const double *cfi6y,
               const double *csi6y, const double *cwi6y, const double *cfy6, const double *csy6, const double *cwy6,
               const double *ppy, const int *nx, const int *nym, const int *ny, const int *nz, const int *npaire) {
  const bool per = ctx && ctx->c.ncl[1];
  return stag(ctx, DPV, 1, ty, uy, per ? cfy6 : cfi6y, per ? csy6 : csi6y, per ? cwy6 : cwi6y, ppy, *nx, *nym, *nz, *ny,
              *nym, *npaire);
}
int x3d_derzvp(x3d_ctx *ctx, double *tz, const double *uz, double *, double *, const double *cfz6, const double *csz6,
               const double *cwz6, const int *nx, const int *ny, const int *nz, const int *nzm, const int *npaire) {
  return stag(ctx, DVP, 2, tz, uz, cfz6, csz6, cwz6, nullptr, *nx, *ny, *nz, *nz, *nzm, *npaire);
}
int x3d_interzvp(x3d_ctx *ctx, double *tz, const double *uz, double *, double *, const double *cifz6,
                 const double *cisz6, const double *ciwz6, const int *nx, const int *ny, const int *nz, const int *nzm,
                 const int *npaire) {
  return stag(ctx, IVP, 2, tz, uz, cifz6, cisz6, ciwz6, nullptr, *nx, *ny, *nz, *nz, *nzm, *npaire);
}
int x3d_derzpv(x3d_ctx *ctx, double *tz, const double *uz, double *, double *, const double *cfiz6,
               const double *csiz6, const double *cwiz6, const double *cfz6, const double *csz6, const double *cwz6,
               const int *nx, const int *ny, const int *nzm, const int *nz, const int *npaire) {
  const bool per = ctx && ctx->c.ncl[2];
  return stag(ctx, DPV, 2, tz, uz, per ? cfz6 : cfiz6, per ? csz6 : csiz6, per ? cwz6 : cwiz6, nullptr, *nx, *ny, *nzm,
              *nz, *nzm, *npaire);
}
int x3d_interzpv(x3d_ctx *ctx, double *tz, const double *uz, double *, double *, const double *cifiz6,
                 const double *cisiz6, const double *ciwiz6, const double *cifz6, const double *cisz6,
                 const double *ciwz6, const int *nx, const int *ny, const int *nzm, const int *nz, const int *npaire) {
  const bool per = ctx && ctx->c.ncl[2];
  return stag(ctx, IPV, 2, tz, uz, per ? cifz6 : cifiz6, per ? cisz6 : cisiz6, per ? ciwz6 : ciwiz6, nullptr, *nx, *ny,
              *nzm, *nz, *nzm, *npaire);
}


// ---- per-launch timing ------------------------------------------------------------------------
int x3d_profile_begin(x3d_ctx *ctx) {
  return guard([&] {
    X3D_CUDA(cudaStreamSynchronize(ctx->c.stream));
    for (auto &r : ctx->c.prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ctx->c.prof.clear();
    ctx->c.profiling = true;
  });
}
// writes a JSON array [{"name":..,"count":..,"total_ms":..,"avg_ms":..},..] into buf
int x3d_profile_end(x3d_ctx *ctx, char *buf, int cap) {
  return guard([&] {
    Ctx &c = ctx->c;
    c.profiling = false;
    X3D_CUDA(cudaStreamSynchronize(c.stream));
    std::map<std::string, std::pair<int, double>> acc;
    for (auto &r : c.prof) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.a, r.b);
      auto &e = acc[r.name];
      e.first += 1; e.second += ms;
      cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    c.prof.clear();
    std::string out = "[";
    bool first = true;
    for (auto &kv : acc) {
      char line[256];
      snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"count\":%d,\"total_ms\":%.6f,\"avg_ms\":%.6f}", first ? "" : ",",
               kv.first.c_str(), kv.second.first, kv.second.second, kv.second.second / kv.second.first);
      out += line;
      first = false;
    }
    out += "]";
    if (static_cast<int>(out.size()) + 1 > cap) throw Error("x3d_profile_end: buffer too small");
    std::memcpy(buf, out.c_str(), out.size() + 1);
  });
}

// ---- decomposition / transposes ----------------------------------------------------------
int x3d_decomp_init(x3d_ctx *ctx, int nx, int ny, int nz, int p_row, int p_col, int rank, int nranks, const void *id) {
  return guard([&] { decomp_init(ctx->c, nx, ny, nz, p_row, p_col, rank, nranks, id); });
}
int x3d_nccl_unique_id(void *out128) {
  return guard([&] { nccl_unique_id(out128); });
}
int x3d_decomp_info_init(x3d_ctx *ctx, int nx, int ny, int nz, int *decomp_id) {
  return guard([&] { *decomp_id = decomp_info_init(ctx->c, nx, ny, nz); });
}
int x3d_decomp_info_get(x3d_ctx *ctx, int decomp_id, x3d_decomp_info *out) {
  return guard([&] { decomp_info_get(ctx->c, decomp_id, out); });
}
int x3d_decomp_compute(int nx, int ny, int nz, int p_row, int p_col, int rank, x3d_decomp_info *out) {
  return guard([&] { decomp_compute(nx, ny, nz, p_row, p_col, rank, out); });
}
int x3d_transpose_plan(int nx, int ny, int nz, int p_row, int p_col, int rank, int which, int *npeers, int *peer_ranks,
                       long long *scount, long long *sdispl, long long *rcount, long long *rdispl, int *send_dims,
                       int *recv_dims) {
  return guard([&] {
    transpose_plan_host(nx, ny, nz, p_row, p_col, rank, which, npeers, peer_ranks, scount, sdispl, rcount, rdispl, send_dims, recv_dims);
  });
}
int x3d_transpose_pack(x3d_ctx *ctx, int which, const double *src, double *packed, int decomp_id, int complex_) {
  return guard([&] { transpose_pack(ctx->c, which, src, packed, decomp_id, complex_ ? 2 : 1); });
}
int x3d_transpose_unpack(x3d_ctx *ctx, int which, const double *packed, double *dst, int decomp_id, int complex_) {
  return guard([&] { transpose_unpack(ctx->c, which, packed, dst, decomp_id, complex_ ? 2 : 1); });
}
#define X3D_DEF_TR(name, which)                                                                            \
  int x3d_transpose_##name(x3d_ctx *ctx, const double *src, double *dst, int id) {                         \
    return guard([&] { transpose(ctx->c, which, src, dst, id, 1); });                                      \
  }                                                                                                        \
  int x3d_transpose_##name##_complex(x3d_ctx *ctx, const double *src, double *dst, int id) {               \
    return guard([&] { transpose(ctx->c, which, src, dst, id, 2); });                                      \
  }
X3D_DEF_TR(x_to_y, 0)
X3D_DEF_TR(y_to_z, 1)
X3D_DEF_TR(z_to_y, 2)
X3D_DEF_TR(y_to_x, 3)

// ---- Poisson ---------------------------------------------------------------------------------
int x3d_poisson_init(x3d_ctx *ctx, const x3d_poisson_params *p) {
  return guard([&] { if (!p) throw Error("null params"); poisson_init(ctx->c, *p); });
}
int x3d_poisson(x3d_ctx *ctx, double *rhs) {
  return guard([&] {
    Ctx &c = ctx->c;
    X3D_CUDA(cudaSetDevice(c.device));
    if (is_device_ptr(rhs)) { poisson_solve_device(c, rhs); return; }
    int d[3];
    poisson_dims(c, d);
    const size_t bytes = static_cast<size_t>(d[0]) * d[1] * d[2] * sizeof(double);
    c.stage_in.reserve(bytes);
    X3D_CUDA(cudaMemcpyAsync(c.stage_in.p, rhs, bytes, cudaMemcpyHostToDevice, c.stream));
    poisson_solve_device(c, static_cast<double *>(c.stage_in.p));
    X3D_CUDA(cudaMemcpyAsync(rhs, c.stage_in.p, bytes, cudaMemcpyDeviceToHost, c.stream));
    X3D_CUDA(cudaStreamSynchronize(c.stream));
  });
}

// ---- device-resident solver -----------------------------------------------------------------------
int x3d_solver_init(x3d_ctx *ctx, const x3d_solver_params *p) {
  return guard([&] { if (!p) throw Error("null params"); solver_init(ctx->c, *p); });
}
int x3d_solver_init_channel(x3d_ctx *ctx) {
  return guard([&] { solver_init_channel(ctx->c); });
}
int x3d_solver_init_tgv(x3d_ctx *ctx) { return guard([&] { solver_init_tgv(ctx->c); }); }
int x3d_solver_set_velocity(x3d_ctx *ctx, const double *ux, const double *uy, const double *uz) {
  return guard([&] { solver_set_velocity(ctx->c, ux, uy, uz); });
}
int x3d_solver_get_velocity(x3d_ctx *ctx, double *ux, double *uy, double *uz) {
  return guard([&] { solver_get_velocity(ctx->c, ux, uy, uz); decomp_check(ctx->c); });
}
int x3d_solver_local_shape(x3d_ctx *ctx, int *dims3, int *zstart0) {
  return guard([&] { solver_local_shape(ctx->c, dims3, zstart0); });
}
int x3d_solver_advance_host(x3d_ctx *ctx, const double *ux_in, const double *uy_in, const double *uz_in, double *ux_out,
                            double *uy_out, double *uz_out, int nsteps) {
  return guard([&] {
    const double *in[3] = {ux_in, uy_in, uz_in};
    double *out[3] = {ux_out, uy_out, uz_out};
    solver_advance_host(ctx->c, in, out, nsteps);
  });
}
int x3d_transpose_selftest(x3d_ctx *ctx, int which, int decomp_id, int complex_, int mode, long long *mismatches) {
  return guard([&] { *mismatches = transpose_selftest(ctx->c, which, decomp_id, complex_ ? 2 : 1, mode); });
}
int x3d_decomp_stats(x3d_ctx *ctx, unsigned long long *remote_bytes, unsigned long long *fields) {
  return guard([&] { decomp_stats(ctx->c, remote_bytes, fields); });
}
int x3d_solver_set_case(x3d_ctx *ctx, const x3d_case_params *c) {
  return guard([&] { if (!c) throw Error("null params"); solver_set_case(ctx->c, *c); });
}
int x3d_solver_set_ibm_mask(x3d_ctx *ctx, const double *ep1) { return guard([&] { solver_set_ibm_mask(ctx->c, ep1); }); }
int x3d_solver_set_inflow_noise(x3d_ctx *ctx, const double *bxo, const double *byo, const double *bzo) {
  return guard([&] { solver_set_inflow_noise(ctx->c, bxo, byo, bzo); });
}
int x3d_solver_set_wall_velocity_x(x3d_ctx *ctx, const double *const planes6[6]) {
  return guard([&] { solver_wall_velocity_x(ctx->c, planes6, nullptr); });
}
int x3d_solver_get_wall_velocity_x(x3d_ctx *ctx, double *const planes6[6]) {
  return guard([&] { solver_wall_velocity_x(ctx->c, nullptr, planes6); });
}
int x3d_solver_apply_spatial_filter(x3d_ctx *ctx, int ifilter, double af) {
  return guard([&] { solver_apply_spatial_filter(ctx->c, ifilter, af); });
}
int x3d_solver_init_cyl(x3d_ctx *ctx) { return guard([&] { solver_init_cyl(ctx->c); }); }
int x3d_solver_host_sync(x3d_ctx *ctx) { return guard([&] { solver_host_sync(ctx->c); }); }
int x3d_solver_step(x3d_ctx *ctx, int nsteps) { return guard([&] { solver_step(ctx->c, nsteps); }); }
int x3d_solver_diagnostics_tgv(x3d_ctx *ctx, double *out5) { return guard([&] { solver_diagnostics_tgv(ctx->c, out5); decomp_check(ctx->c); }); }
int x3d_solver_divergence(x3d_ctx *ctx, double *divmax, double *divmean) {
  return guard([&] { solver_divergence(ctx->c, divmax, divmean); });
}

// ---- schemes() for non-Fortran hosts (src/schemes.f90); CPU only, no context needed -----------------
// which: 0 first derivative (ff,fs,fw), 1 its p-variant, 2 second derivative, 3 p-variant, 4 cfx6.., 5 cfxp6..,
//        6 cifx6.., 7 cifxp6.., 8 cfi6.., 9 cfip6.., 10 cifi6.., 11 cifip6..
int x3d_stretching(int istret, double beta, double yly, int ny, int nym, double *out8, double *alpha) {
  return guard([&] {
    if (istret < 1 || istret > 3 || ny < 3 || !out8) throw Error("x3d_stretching: bad argument");
    const StretchY S = make_stretching(istret, beta, yly, ny, nym);
    const std::vector<double> *v[8] = {&S.yp, &S.ypi, &S.ppy, &S.pp2y, &S.pp4y, &S.ppyi, &S.pp2yi, &S.pp4yi};
    for (int q = 0; q < 8; ++q) std::copy(v[q]->begin(), v[q]->begin() + ny, out8 + static_cast<size_t>(q) * ny);
    if (alpha) *alpha = S.alpha;
  });
}

int x3d_filter_axis(int n, int ncl1, int ncln, double af, x3d_filter_coeffs *coeffs, int which, double *f, double *s, double *w) {
  return guard([&] {
    if (which < 0 || which > 1) throw Error("x3d_filter_axis: bad selector");
    x3d_filter_coeffs c;
    LU3 plain, p;
    make_filter_axis(n, ncl1, ncln, af, c, plain, p);
    if (coeffs) *coeffs = c;
    const LU3 &L = which == 1 ? p : plain;
    if (f) std::copy(L.f.begin(), L.f.end(), f);
    if (s) std::copy(L.s.begin(), L.s.end(), s);
    if (w) std::copy(L.w.begin(), L.w.end(), w);
  });
}

int x3d_schemes_axis(int n, int ncl1, int ncln, double len, int ifirstder, int isecondder, int ipinter, double nu0nu,
                     double cnu, x3d_deriv_coeffs *coeffs, int which, double *f, double *s, double *w) {
  return guard([&] {
    SchemeOpts o; o.ifirstder = ifirstder; o.isecondder = isecondder; o.ipinter = ipinter; o.nu0nu = nu0nu; o.cnu = cnu;
    AxisCoeffs A = make_axis_coeffs(n, ncl1, ncln, len, o);
    if (coeffs) *coeffs = A.c;
    const LU3 *t[12] = {&A.d1, &A.d1p, &A.d2, &A.d2p, &A.vp, &A.vpp, &A.ivp, &A.ivpp, &A.pv, &A.pvp, &A.ipv, &A.ipvp};
    if (which < 0 || which > 11) throw Error("x3d_schemes_axis: bad selector");
    const LU3 &L = *t[which];
    if (f) std::copy(L.f.begin(), L.f.end(), f);
    if (s) std::copy(L.s.begin(), L.s.end(), s);
    if (w) std::copy(L.w.begin(), L.w.end(), w);
  });
}

}  // extern "C"
