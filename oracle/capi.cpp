// capi.cpp -- C entry points of the oracle for ctypes (TEST INFRASTRUCTURE).
#include <cmath>
#include <cstring>
#include <string>
#include <stdexcept>
#include <omp.h>
#include "x3d_oracle.hpp"

using namespace x3do;

static thread_local std::string g_err;

static vec *axis_array(AxisScheme *a, const std::string &nm) {
#define A(x) if (nm == #x) return &a->x;
  A(ff) A(fs) A(fw) A(ffp) A(fsp) A(fwp) A(sf) A(ss) A(sw) A(sfp) A(ssp) A(swp)
  A(cfx6) A(ccx6) A(cbx6) A(cfxp6) A(csxp6) A(cwxp6) A(csx6) A(cwx6)
  A(cifx6) A(cicx6) A(cibx6) A(cifxp6) A(cisxp6) A(ciwxp6) A(cisx6) A(ciwx6)
  A(cfi6) A(cci6) A(cbi6) A(cfip6) A(csip6) A(cwip6) A(csi6) A(cwi6)
  A(cifi6) A(cici6) A(cibi6) A(cifip6) A(cisip6) A(ciwip6) A(cisi6) A(ciwi6)
  A(fiff) A(fifs) A(fifw) A(fiffp) A(fifsp) A(fifwp)
#undef A
  return nullptr;
}

// parse "derx_11", "deryy_00", "filz_22", "derxvp", "interzpv" ...
static bool parse_op(const char *name, OpKind &kind, int &axis, int &ncl1, int &ncln) {
  std::string s(name);
  auto ax = [](char ch) { return ch == 'x' ? 0 : ch == 'y' ? 1 : ch == 'z' ? 2 : -1; };
  ncl1 = ncln = -1;
  if (s.size() >= 6 && (s.substr(s.size() - 2) == "vp" || s.substr(s.size() - 2) == "pv")) {
    const bool vp = s.substr(s.size() - 2) == "vp";
    const bool inter = s.rfind("inter", 0) == 0;
    axis = ax(s[inter ? 5 : 3]);
    kind = inter ? (vp ? IVP : IPV) : (vp ? DVP : DPV);
    return axis >= 0;
  }
  auto us = s.find('_');
  if (us == std::string::npos || s.size() != us + 3) return false;
  ncl1 = s[us + 1] - '0';
  ncln = s[us + 2] - '0';
  std::string head = s.substr(0, us);
  if (head.rfind("fil", 0) == 0 && head.size() == 4) { kind = FIL; axis = ax(head[3]); return axis >= 0; }
  if (head.rfind("der", 0) == 0 && head.size() == 4) { kind = D1; axis = ax(head[3]); return axis >= 0; }
  if (head.rfind("der", 0) == 0 && head.size() == 5 && head[3] == head[4]) { kind = D2; axis = ax(head[3]); return axis >= 0; }
  return false;
}

extern "C" {

const char *x3do_last_error() { return g_err.c_str(); }

void *x3do_axis_create(int n, int ncl1, int ncln, double len, int ifirstder, int isecondder, int ipinter,
                       double nu0nu, double cnu) {
  try {
    SchemeOptions o;
    o.ifirstder = ifirstder; o.isecondder = isecondder; o.ipinter = ipinter; o.nu0nu = nu0nu; o.cnu = cnu;
    return new AxisScheme(make_axis(n, ncl1, ncln, len, o));
  } catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
void x3do_axis_destroy(void *a) { delete static_cast<AxisScheme *>(a); }
void x3do_axis_set_filter(void *a, double af) { set_filter_coefficients(*static_cast<AxisScheme *>(a), af); }
int x3do_axis_nm(void *a) { return static_cast<AxisScheme *>(a)->nm; }
double x3do_axis_d(void *a) { return static_cast<AxisScheme *>(a)->d; }
int x3do_axis_get_array(void *a, const char *name, double *out, int cap) {
  vec *v = axis_array(static_cast<AxisScheme *>(a), name);
  if (!v) return -1;
  const int n = static_cast<int>(v->size());
  if (out && cap >= n) std::memcpy(out, v->data(), n * sizeof(double));
  return n;
}
void x3do_axis_get_coeffs(void *a, x3d_deriv_coeffs *c, x3d_filter_coeffs *fc) {
  auto *s = static_cast<AxisScheme *>(a);
  if (c) *c = s->c;
  if (fc) *fc = s->fc;
}

// the reference operator interface: caller passes LU arrays and module scalars
int x3do_op(const char *name, const int *dims_in, int npaire, const double *u, double *t, const double *f,
            const double *s, const double *w, const double *post, const x3d_deriv_coeffs *c,
            const x3d_filter_coeffs *fc, int periodic, int rhs_only) {
  try {
    OpDesc op{};
    int axis;
    if (!parse_op(name, op.kind, axis, op.ncl1, op.ncln)) { g_err = std::string("unknown operator ") + name; return 1; }
    op.periodic = periodic != 0;
    op.npaire = npaire;
    op.f = f; op.s = s; op.w = w; op.c = c; op.fc = fc; op.post = post; op.rhs_only = rhs_only != 0;
    const int ext = dims_in[axis];
    if (op.kind == DVP || op.kind == IVP) { op.n = ext; op.nm = op.periodic ? ext : ext - 1; }
    else if (op.kind == DPV || op.kind == IPV) { op.nm = ext; op.n = op.periodic ? ext : ext + 1; }
    else { op.n = ext; op.nm = ext; }
    apply_op(op, axis, dims_in, u, t);
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 2; }
}


// ---- Poisson -----------------------------------------------------------------------
struct PoissonBox {
  AxisScheme X, Y, Z;
  Stretch st;
  Poisson po;
};
// stretching_full (stretching.f90:96-318): out8 = yp, ypi, ppy, pp2y, pp4y, ppyi, pp2yi, pp4yi (ny each)
int x3do_stretching(int istret, double beta, double yly, int ny, int nym, double *out8, double *alpha) {
  try {
    Stretch s;
    s.istret = istret; s.beta = beta; s.yly = yly;
    stretching(s, ny, nym, 0, 0, false);
    const vec *v[8] = {&s.yp, &s.ypi, &s.ppy, &s.pp2y, &s.pp4y, &s.ppyi, &s.pp2yi, &s.pp4yi};
    for (int q = 0; q < 8; ++q) std::memcpy(out8 + static_cast<size_t>(q) * ny, v[q]->data(), ny * sizeof(double));
    if (alpha) *alpha = s.alpha;
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
void *x3do_poisson_create_stretched(int nx, int ny, int nz, const int *ncl6, double xlx, double yly, double zlz, int ifirstder,
                                    int ipinter, int istret, double beta) {
  try {
    SchemeOptions o; o.ifirstder = ifirstder; o.ipinter = ipinter;
    auto *b = new PoissonBox();
    b->X = make_axis(nx, ncl6[0], ncl6[1], xlx, o);
    b->Y = make_axis(ny, ncl6[2], ncl6[3], yly, o);
    b->Z = make_axis(nz, ncl6[4], ncl6[5], zlz, o);
    if (istret != 0) {
      b->st.istret = istret; b->st.beta = beta; b->st.yly = yly;
      stretching(b->st, ny, b->Y.nm, ncl6[2], ncl6[3], b->Y.periodic);
    }
    b->po.init(b->X, b->Y, b->Z, istret ? &b->st : nullptr);
    return b;
  } catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
// tables of the solver by name; complex arrays are returned as (re,im) pairs; returns the number of doubles
long x3do_poisson_get(void *p, const char *name, double *out, long cap) {
  auto &po = static_cast<PoissonBox *>(p)->po;
  const std::string n = name;
  const vec *rv = nullptr;
  const std::vector<cplx> *cv = nullptr;
  if (n == "ax") rv = &po.ax; else if (n == "bx") rv = &po.bx; else if (n == "ay") rv = &po.ay; else if (n == "by") rv = &po.by;
  else if (n == "az") rv = &po.az; else if (n == "bz") rv = &po.bz;
  else if (n == "xkx") cv = &po.xkx; else if (n == "xk2") cv = &po.xk2; else if (n == "exs") cv = &po.exs;
  else if (n == "yky") cv = &po.yky; else if (n == "yk2") cv = &po.yk2; else if (n == "eys") cv = &po.eys;
  else if (n == "zkz") cv = &po.zkz; else if (n == "zk2") cv = &po.zk2; else if (n == "ezs") cv = &po.ezs;
  else if (n == "kxyz") cv = &po.kxyz; else if (n == "a") cv = &po.a; else if (n == "a2") cv = &po.a2; else if (n == "a3") cv = &po.a3;
  else return -1;
  const long cnt = rv ? static_cast<long>(rv->size()) : 2 * static_cast<long>(cv->size());
  if (out && cap >= cnt) std::memcpy(out, rv ? static_cast<const void *>(rv->data()) : static_cast<const void *>(cv->data()), cnt * sizeof(double));
  return cnt;
}
// inversion5_v1 (version 1) / inversion5_v2 (version 2) on caller arrays: aaa (nx,n,nz,5) complex, eee (nx,n,nz) complex
int x3do_inversion5(int version, double *aaa, double *eee, int nx, int n, int nz) {
  try {
    const size_t cnt = static_cast<size_t>(nx) * n * nz * 5;
    std::vector<cplx> a(reinterpret_cast<cplx *>(aaa), reinterpret_cast<cplx *>(aaa) + cnt);
    if (version == 1) inversion5_v1(a, reinterpret_cast<cplx *>(eee), nx, n, nz);
    else { inversion5_v2(a, reinterpret_cast<cplx *>(eee), nx, n, nz); std::memcpy(aaa, a.data(), cnt * sizeof(cplx)); }
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
void *x3do_poisson_create(int nx, int ny, int nz, const int *ncl6, double xlx, double yly, double zlz, int ifirstder,
                          int ipinter) {
  try {
    SchemeOptions o; o.ifirstder = ifirstder; o.ipinter = ipinter;
    auto *b = new PoissonBox();
    b->X = make_axis(nx, ncl6[0], ncl6[1], xlx, o);
    b->Y = make_axis(ny, ncl6[2], ncl6[3], yly, o);
    b->Z = make_axis(nz, ncl6[4], ncl6[5], zlz, o);
    b->po.init(b->X, b->Y, b->Z, nullptr);
    return b;
  } catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
void x3do_poisson_destroy(void *p) { delete static_cast<PoissonBox *>(p); }
int x3do_poisson_solve(void *p, double *rhs) {
  try { static_cast<PoissonBox *>(p)->po.solve(rhs); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// kxyz as (nx,ny,nz/2+1) complex; returns element count
long x3do_poisson_kxyz(void *p, double *out_re_im) {
  auto &po = static_cast<PoissonBox *>(p)->po;
  if (out_re_im) std::memcpy(out_re_im, po.kxyz.data(), po.kxyz.size() * sizeof(cplx));
  return static_cast<long>(po.kxyz.size());
}
void x3do_poisson_dims(void *p, int *d3) { auto &po = static_cast<PoissonBox *>(p)->po; d3[0] = po.nx; d3[1] = po.ny; d3[2] = po.nz; }

// ---- solver ---------------------------------------------------------------------------
void *x3do_solver_create(int nx, int ny, int nz, const int *ncl6, double xlx, double yly, double zlz, double re, double dt,
                         int itimescheme, int ifirstder, int isecondder, int ipinter, int istret, double beta) {
  try {
    auto *s = new Solver();
    s->p.nx = nx; s->p.ny = ny; s->p.nz = nz;
    for (int a = 0; a < 3; ++a) { s->p.ncl[a][0] = ncl6[2 * a]; s->p.ncl[a][1] = ncl6[2 * a + 1]; }
    s->p.xlx = xlx; s->p.yly = yly; s->p.zlz = zlz; s->p.re = re; s->p.dt = dt; s->p.itimescheme = itimescheme;
    s->p.opt.ifirstder = ifirstder; s->p.opt.isecondder = isecondder; s->p.opt.ipinter = ipinter;
    s->p.istret = istret; s->p.beta = beta;
    s->init();
    return s;
  } catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
// channel flow (BASELINE config #3): Dirichlet walls in y, constant flow rate, optional stretched mesh
void *x3do_solver_create_case(int nx, int ny, int nz, const int *ncl6, double xlx, double yly, double zlz, double re, double dt,
                              int itimescheme, int ifirstder, int isecondder, int ipinter, int istret, double beta, int itype,
                              double nu0nu, double cnu) {
  try {
    auto *s = new Solver();
    s->p.nx = nx; s->p.ny = ny; s->p.nz = nz;
    for (int a = 0; a < 3; ++a) { s->p.ncl[a][0] = ncl6[2 * a]; s->p.ncl[a][1] = ncl6[2 * a + 1]; }
    s->p.xlx = xlx; s->p.yly = yly; s->p.zlz = zlz; s->p.re = re; s->p.dt = dt; s->p.itimescheme = itimescheme;
    s->p.opt.ifirstder = ifirstder; s->p.opt.isecondder = isecondder; s->p.opt.ipinter = ipinter;
    s->p.opt.nu0nu = nu0nu; s->p.opt.cnu = cnu;
    s->p.istret = istret; s->p.beta = beta; s->p.itype = itype;
    s->init();
    return s;
  } catch (std::exception &e) { g_err = e.what(); return nullptr; }
}
void x3do_solver_init_channel(void *s) { static_cast<Solver *>(s)->init_channel(); }
// pieces of the step on caller data (tests/test_oracle_step_golden.py): wall pressure gradients in the order
// dpdyx1 dpdzx1 dpdyxn dpdzxn | dpdxy1 dpdzy1 dpdxyn dpdzyn | dpdxz1 dpdyz1 dpdxzn dpdyzn
static std::vector<double> *dpd_array(Solver *s, int q) {
  std::vector<double> *v[12] = {&s->dpdyx1, &s->dpdzx1, &s->dpdyxn, &s->dpdzxn, &s->dpdxy1, &s->dpdzy1, &s->dpdxyn, &s->dpdzyn,
                                &s->dpdxz1, &s->dpdyz1, &s->dpdxzn, &s->dpdyzn};
  return v[q];
}
void x3do_solver_set_wall_gradient(void *sv, int q, const double *in) { auto *v = dpd_array(static_cast<Solver *>(sv), q); std::memcpy(v->data(), in, v->size() * 8); }
void x3do_solver_get_wall_gradient(void *sv, int q, double *out) { auto *v = dpd_array(static_cast<Solver *>(sv), q); std::memcpy(out, v->data(), v->size() * 8); }
// wall velocity plane q in the order bxx1 bxy1 bxz1 bxxn bxyn bxzn | byx1 .. byzn | bzx1 .. bzzn
void x3do_solver_set_wall_velocity(void *sv, int q, const double *in) { auto &v = static_cast<Solver *>(sv)->bw[q]; std::memcpy(v.data(), in, v.size() * 8); }
void x3do_solver_get_wall_velocity(void *sv, int q, double *out) { auto &v = static_cast<Solver *>(sv)->bw[q]; std::memcpy(out, v.data(), v.size() * 8); }
// inflow / outflow of the cylinder case on the solver's velocity; noise planes bxo, byo, bzo may be NULL (zero)
int x3do_solver_inflow_outflow(void *sv, int itr, const double *gdt3, double u1, double u2, double inflow_noise, const double *bxo,
                               const double *byo, const double *bzo) {
  try {
    auto *s = static_cast<Solver *>(sv);
    for (int q = 0; q < 3; ++q) s->gdt[q] = gdt3[q];
    s->itr = itr; s->p.u1 = u1; s->p.u2 = u2; s->p.inflow_noise = inflow_noise;
    if (bxo) std::memcpy(s->bxo.data(), bxo, s->bxo.size() * 8);
    if (byo) std::memcpy(s->byo.data(), byo, s->byo.size() * 8);
    if (bzo) std::memcpy(s->bzo.data(), bzo, s->bzo.size() * 8);
    s->inflow();
    s->outflow();
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// immersed boundary of the solver: iibm, ep1 (nx,ny,nz), wall velocity; geometry of one direction (copied)
int x3do_solver_set_ibm(void *sv, int iibm, const double *ep1, const double *ubc3) {
  try {
    auto *s = static_cast<Solver *>(sv);
    s->iibm = iibm;
    s->ep1.assign(ep1, ep1 + s->ux.size());
    for (int q = 0; q < 3; ++q) s->ubc[q] = ubc3 ? ubc3[q] : 0.0;
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
int x3do_solver_set_ibm_geometry(void *sv, int axis, int nobjmax, int npif, int izap, const int *nobj, const double *xi, const double *xf,
                                 const int *nipif, const int *nfpif) {
  try {
    auto *s = static_cast<Solver *>(sv);
    const int n[3] = {s->p.nx, s->p.ny, s->p.nz};
    const size_t nl = static_cast<size_t>(n[axis == 0 ? 1 : 0]) * n[axis == 2 ? 1 : 2];
    auto &I = s->ibm[axis];
    I.nobj.assign(nobj, nobj + nl);
    I.xi.assign(xi, xi + nl * nobjmax); I.xf.assign(xf, xf + nl * nobjmax);
    I.nipif.assign(nipif, nipif + nl * (nobjmax + 1)); I.nfpif.assign(nfpif, nfpif + nl * (nobjmax + 1));
    I.g.nobjmax = nobjmax; I.g.npif = npif; I.g.izap = izap;
    I.g.nobj = I.nobj.data(); I.g.xi = I.xi.data(); I.g.xf = I.xf.data(); I.g.nipif = I.nipif.data(); I.g.nfpif = I.nfpif.data();
    I.set = true;
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// channel forcing of the assembled step (momentum_forcing_channel, Case-Channel.f90:396-420); with cpg the viscosity and
// the pressure gradient follow parameters.f90:303-311 as in Solver::init
void x3do_solver_set_channel_forcing(void *sv, int cpg, double wrotation, int spinup_time, int iin) {
  auto *s = static_cast<Solver *>(sv);
  s->p.cpg = cpg != 0; s->p.wrotation = wrotation; s->p.spinup_time = spinup_time; s->p.iin = iin;
  s->xnu = 1.0 / s->p.re;
  s->fcpg = 0.0;
  if (s->p.cpg) {
    const double re_cent = std::pow(s->p.re / 0.116, 1.0 / 0.88);
    s->xnu = 1.0 / re_cent;
    s->fcpg = 2.0 / s->p.yly * ((s->p.re / re_cent) * (s->p.re / re_cent));
  }
}
// inflow noise of the assembled cylinder step: the random planes bxo, byo, bzo (ny, nz) and their amplitude
void x3do_solver_set_inflow_noise(void *sv, double inflow_noise, const double *bxo, const double *byo, const double *bzo) {
  auto *s = static_cast<Solver *>(sv);
  s->p.inflow_noise = inflow_noise;
  if (bxo) std::memcpy(s->bxo.data(), bxo, s->bxo.size() * 8);
  if (byo) std::memcpy(s->byo.data(), byo, s->byo.size() * 8);
  if (bzo) std::memcpy(s->bzo.data(), bzo, s->bzo.size() * 8);
}
void x3do_solver_init_cyl(void *sv, double u1, double u2) {
  auto *s = static_cast<Solver *>(sv);
  s->p.u1 = u1; s->p.u2 = u2;
  s->init_cyl();
}
// momentum_forcing (case.f90:538 -> Case-Channel.f90:396-420) on caller arrays, with the solver's velocity
int x3do_solver_momentum_forcing(void *sv, long long itime, int cpg, double fcpg, double wrotation, int spinup_time, int iin, double *dux,
                                 double *duy, double *duz) {
  try {
    auto *s = static_cast<Solver *>(sv);
    s->itime = static_cast<int>(itime);
    s->p.cpg = cpg != 0; s->fcpg = fcpg; s->p.wrotation = wrotation; s->p.spinup_time = spinup_time; s->p.iin = iin;
    s->momentum_forcing(dux, duy, duz);
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
void x3do_ibm_body(double *ux, double *uy, double *uz, const double *ep, long long n) { ibm_body(ux, uy, uz, ep, static_cast<size_t>(n)); }
void x3do_ibm_corgp(double *ux, double *uy, double *uz, const double *px, const double *py, const double *pz, long long n, int nlock) {
  ibm_corgp(ux, uy, uz, px, py, pz, static_cast<size_t>(n), nlock);
}
int x3do_solver_pre_correc(void *sv, int itr, const double *gdt3) {
  try {
    auto *s = static_cast<Solver *>(sv);
    for (int q = 0; q < 3; ++q) s->gdt[q] = gdt3[q];
    s->itr = itr;
    s->pre_correc();
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// one call of intt (src/time_integrators.f90:20-190) on caller data: var[n], dvar[n * ntime] (Fortran dvar1(:,:,:,q) blocks)
int x3do_solver_intt(void *sv, long long itime, int itr, long long n, double *var, double *dvar) {
  try {
    auto *s = static_cast<Solver *>(sv);
    s->itime = itime; s->itr = itr;
    std::vector<double> v(var, var + n), d[3];
    for (int q = 0; q < s->ntime; ++q) d[q].assign(dvar + q * n, dvar + (q + 1) * n);
    s->intt(v, d);
    std::memcpy(var, v.data(), n * 8);
    for (int q = 0; q < s->ntime; ++q) std::memcpy(dvar + q * n, d[q].data(), n * 8);
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// adt, bdt, cdt, gdt (3 each), then ntime and iadvance_time (src/variables.f90:1340-1423)
void x3do_solver_time_coefficients(void *sv, double *out14) {
  auto *s = static_cast<Solver *>(sv);
  for (int q = 0; q < 3; ++q) { out14[q] = s->adt[q]; out14[3 + q] = s->bdt[q]; out14[6 + q] = s->cdt[q]; out14[9 + q] = s->gdt[q]; }
  out14[12] = s->ntime; out14[13] = s->iadvance_time;
}
int x3do_solver_capture_wall_gradients(void *sv, int itr, const double *gdt3, const double *px, const double *py, const double *pz) {
  try {
    auto *s = static_cast<Solver *>(sv);
    for (int q = 0; q < 3; ++q) s->gdt[q] = gdt3[q];
    s->itr = itr;
    s->capture_wall_gradients(px, py, pz);
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// lagpolx / lagpoly / lagpolz (src/ibm.f90:83-343) on caller data; coords may be NULL for uniform directions
void x3do_lagpol(double *u, int nx, int ny, int nz, int axis, int nobjmax, int npif, int izap, const int *nobj, const double *xi,
                 const double *xf, const int *nipif, const int *nfpif, const double *coords, double d, double len) {
  IbmGeom g;
  g.nobjmax = nobjmax; g.npif = npif; g.izap = izap; g.nobj = nobj; g.xi = xi; g.xf = xf; g.nipif = nipif; g.nfpif = nfpif;
  lagpol(u, nx, ny, nz, axis, g, coords, d, len);
}
// cubsplx / cubsply / cubsplz (src/ibm.f90:399-874); ana_i / ana_f may be NULL (ianal = 0)
int x3do_cubspl(double *u, int nx, int ny, int nz, int axis, int nobjmax, int npif, int izap, const int *nobj, const double *xi,
                const double *xf, const int *nipif, const int *nfpif, const double *coords, double d, double len, double lind,
                const double *ana_i, const double *ana_f) {
  try {
    IbmGeom g;
    g.nobjmax = nobjmax; g.npif = npif; g.izap = izap; g.nobj = nobj; g.xi = xi; g.xf = xf; g.nipif = nipif; g.nfpif = nfpif;
    cubspl(u, nx, ny, nz, axis, g, coords, d, len, lind, ana_i, ana_f);
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
void x3do_channel_cfr(double *u, int nx, int ny, int nz, const double *ppy, double dy, double yly, double constant) {
  channel_cfr_apply(u, nx, ny, nz, ppy, dy, yly, constant);
}
void x3do_solver_destroy(void *s) { delete static_cast<Solver *>(s); }
void x3do_solver_init_tgv(void *s) { static_cast<Solver *>(s)->init_tgv(); }
int x3do_solver_step(void *s, int nsteps) {
  try { for (int i = 0; i < nsteps; ++i) static_cast<Solver *>(s)->step(); return 0; }
  catch (std::exception &e) { g_err = e.what(); return 1; }
}
int x3do_solver_postprocess_tgv(void *s, double *out4) {
  try { static_cast<Solver *>(s)->postprocess_tgv(out4); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
int x3do_solver_divergence(void *sv, double *tmax, double *tmoy) {
  try {
    auto *s = static_cast<Solver *>(sv);
    std::vector<double> dv(s->pp3.size());
    s->divergence(dv.data(), 2, tmax, tmoy);
    return 0;
  } catch (std::exception &e) { g_err = e.what(); return 1; }
}
void x3do_solver_get_velocity(void *sv, double *ux, double *uy, double *uz) {
  auto *s = static_cast<Solver *>(sv);
  std::memcpy(ux, s->ux.data(), s->ux.size() * 8); std::memcpy(uy, s->uy.data(), s->uy.size() * 8); std::memcpy(uz, s->uz.data(), s->uz.size() * 8);
}
void x3do_solver_set_velocity(void *sv, const double *ux, const double *uy, const double *uz) {
  auto *s = static_cast<Solver *>(sv);
  std::memcpy(s->ux.data(), ux, s->ux.size() * 8); std::memcpy(s->uy.data(), uy, s->uy.size() * 8); std::memcpy(s->uz.data(), uz, s->uz.size() * 8);
}
// pieces of one sub-step, for stage-by-stage parity tests
int x3do_solver_momentum_rhs(void *sv, double *dux, double *duy, double *duz) {
  try { static_cast<Solver *>(sv)->momentum_rhs_eq(dux, duy, duz); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
int x3do_solver_divergence_of(void *sv, double *pp3, int nlock) {
  try { static_cast<Solver *>(sv)->divergence(pp3, nlock, nullptr, nullptr); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
int x3do_solver_gradp(void *sv, double *px, double *py, double *pz, const double *pp3) {
  try { static_cast<Solver *>(sv)->gradp(px, py, pz, pp3); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
int x3do_solver_poisson(void *sv, double *pp3) {
  try { static_cast<Solver *>(sv)->po.solve(pp3); return 0; } catch (std::exception &e) { g_err = e.what(); return 1; }
}
// OpenMP threads of the oracle: torchrun exports OMP_NUM_THREADS=1, which would time the CPU baseline on one core.
// n > 0 sets the count; returns the count in force.
int x3do_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}
void x3do_solver_pdims(void *sv, int *d3) { auto *s = static_cast<Solver *>(sv); d3[0] = s->nxm; d3[1] = s->nym; d3[2] = s->nzm; }

}  // extern "C"
