// x3d_oracle.hpp -- CPU restatement of the Xcompact3d hot path (TEST INFRASTRUCTURE).
//
// This is the parity oracle: a plain C++17/OpenMP restatement of the reference
// algorithms (xcompact3d/Incompact3d v5.0).  Every function cites the
// reference file:line it follows.  It is pinned by
//   * tests/golden/operators.npz, schemes.npz, poisson.npz -- outputs of the
//     reference's own statements (executed from the Fortran text by
//     tests/golden/f90mini.py in the build container), and
//   * tests/TGV-Taylor-Green-vortex/reference_time_evol.dat of the reference
//     (real Fortran/MPI output), rows 1-10, copied as tests/golden/tgv_reference_time_evol.dat.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (incompact3d_b200) never
// links or calls it.
#pragma once
#include <complex>
#include <cstddef>
#include <vector>
#include "../include/x3d_b200.h"

namespace x3do {

using cplx = std::complex<double>;
using vec = std::vector<double>;

// ---- schemes.f90 ---------------------------------------------------------
struct AxisScheme {
  int n = 0, nm = 0, ncl1 = 0, ncln = 0;
  bool periodic = false;  // nclx / ncly / nclz logical (parameters.f90:273-294)
  double d = 0, len = 0;
  x3d_deriv_coeffs c{};
  x3d_filter_coeffs fc{};
  // first derivative (schemes.f90:443-599)
  vec ff, fs, fw, ffp, fsp, fwp;
  // second derivative (schemes.f90:602-856)
  vec sf, ss, sw, sfp, ssp, swp;
  // staggered, pressure-mesh sized (nm) (schemes.f90:935-1063)
  vec cfx6, ccx6, cbx6, cfxp6, csxp6, cwxp6, csx6, cwx6;
  vec cifx6, cicx6, cibx6, cifxp6, cisxp6, ciwxp6, cisx6, ciwx6;
  // staggered, velocity-mesh sized (n)
  vec cfi6, cci6, cbi6, cfip6, csip6, cwip6, csi6, cwi6;
  vec cifi6, cici6, cibi6, cifip6, cisip6, ciwip6, cisi6, ciwi6;
  // filter (filters.f90:62-219)
  vec fiff, fifs, fifw, fiffp, fifsp, fifwp;
};

struct SchemeOptions {
  int ifirstder = 4, isecondder = 4, ipinter = 3;
  double nu0nu = 4.0, cnu = 0.44;
};

void prepare(const vec &b, const vec &c, const vec &f, vec &s, vec &w, int n);
void first_derivative(AxisScheme &a, const SchemeOptions &o);
void second_derivative(AxisScheme &a, const SchemeOptions &o);
void interpolation(AxisScheme &a, const SchemeOptions &o);
void set_filter_coefficients(AxisScheme &a, double af);
// parameters.f90:273-304 + schemes(): n nodes, (ncl1,ncln), domain length
AxisScheme make_axis(int n, int ncl1, int ncln, double len, const SchemeOptions &o);

// ---- derive.f90 / filters.f90 -----------------------------------------------
enum OpKind { D1 = 0, D2 = 1, FIL = 2, DVP = 3, IVP = 4, DPV = 5, IPV = 6 };
struct OpDesc {
  OpKind kind;
  int ncl1, ncln;   // boundary pair of the routine variant (00,11,12,21,22)
  bool periodic;    // staggered ops branch on the logical only (derive.f90:3816)
  int npaire;
  int n, nm;        // velocity nodes / pressure points along the line
  const double *f, *s, *w;  // LU arrays used by the Thomas sweep
  const x3d_deriv_coeffs *c;
  const x3d_filter_coeffs *fc;
  const double *post;  // ppy / ppyi multiplier (istret /= 0) or nullptr
  bool rhs_only;       // deryy with iimplicit >= 1 (derive.f90:2166)
};
// u: (d0,d1,d2) with the line axis of extent n_in; t same with extent n_out
void apply_op(const OpDesc &op, int axis, const int dims_in[3], const double *u, double *t);
int op_n_in(const OpDesc &op);
int op_n_out(const OpDesc &op);

// ---- FFT (2DECOMP&FFT v2.0.4 decomp_2d_fft_3d semantics) ---------------------
// unnormalised DFT, forward sign -1; in-place on a strided complex line
void fft_line(cplx *x, int n, std::ptrdiff_t stride, int sign);

// ---- stretching.f90 -----------------------------------------------------------
struct Stretch {
  int istret = 0;
  double beta = 0, alpha = 0, yly = 0;
  vec yp, ypi, ppy, pp2y, pp4y, ppyi, pp2yi, pp4yi;
};
void stretching(Stretch &s, int ny, int nym, int ncly1, int nclyn, bool ncly);

// ---- poisson.f90 -----------------------------------------------------------
struct Poisson {
  int nx, ny, nz;   // pressure mesh (nxm,nym,nzm)
  int bcx, bcy, bcz;
  int nzh;          // nz/2+1
  vec ax, bx, ay, by, az, bz;
  std::vector<cplx> kxyz;  // (nx,ny,nzh), x-pencil order, or y-pencil order for 010 (same on 1 rank)
  std::vector<cplx> xk2, yk2, zk2, xkx, yky, zkz, exs, eys, ezs;
  int istret = 0;
  double alpha = 0, beta = 0;
  std::vector<cplx> a, a2, a3;  // pentadiagonal matrices (matrice_refinement)
  const AxisScheme *sx = nullptr, *sy = nullptr, *sz = nullptr;
  const Stretch *st = nullptr;
  void init(const AxisScheme &x, const AxisScheme &y, const AxisScheme &z, const Stretch *st);
  void solve(double *rhs);  // z-pencil (nx,ny,nz) in place
  void abxyz();
  void waves();
  void matrice_refinement();
  void solve_000(double *rhs);
  void solve_100(double *rhs);
  void solve_010(double *rhs);
  void solve_11x(double *rhs);
};
void inversion5_v1(const std::vector<cplx> &aaa_in, cplx *eee, int nx, int nyh, int nz);
void inversion5_v2(std::vector<cplx> &aaa, cplx *eee, int nx, int nym, int nz);

// ---- ibm.f90: Lagrange reconstruction inside the bodies (iibm = 2) ---------------------------------------
struct IbmGeom {           // module complex_geometry for one direction (src/module_param.f90:546-556)
  int nobjmax = 0, npif = 2, izap = 1;
  const int *nobj = nullptr;     // (na, nb)
  const double *xi = nullptr, *xf = nullptr;   // (nobjmax, na, nb)
  const int *nipif = nullptr, *nfpif = nullptr;  // (0:nobjmax, na, nb)
};
void lagpol(double *u, int nx, int ny, int nz, int axis, const IbmGeom &g, const double *coords, double d, double len);
void cubspl(double *u, int nx, int ny, int nz, int axis, const IbmGeom &g, const double *coords, double d, double len, double lind,
            const double *ana_i, const double *ana_f);

// ---- solver (transeq.f90, time_integrators.f90, navier.f90, Case-TGV.f90) ------
struct SolverParams {
  int nx = 65, ny = 65, nz = 65;
  int ncl[3][2] = {{1, 1}, {1, 1}, {1, 1}};
  double xlx = 3.14159265358979, yly = 3.14159265358979, zlz = 3.14159265358979;
  double re = 1600.0, dt = 0.005;
  int itimescheme = 5;
  SchemeOptions opt;
  int istret = 0;
  double beta = 0.259065151;
  int itype = 0;  // 0: TGV-type box; 3: channel (itype_channel): constant flow rate channel_cfr, Case-Channel.f90:150-170;
                  // 5: cylinder wake (itype_cyl): inflow / convective outflow planes, Case-Cylinder-wake.f90:84-203
  double u1 = 1.0, u2 = 1.0, inflow_noise = 0.0;   // module param (inflow / outflow of the cylinder case)
  // channel forcing (Case-Channel.f90:396-420, parameters.f90:303-311): constant pressure gradient instead of the
  // constant flow rate, and the spin-up rotation
  bool cpg = false;
  double wrotation = 0.0;
  int spinup_time = 0, iin = 0;
};
struct Solver {
  SolverParams p;
  AxisScheme X, Y, Z;
  Stretch st;
  Poisson po;
  int nxm, nym, nzm;
  double xnu;
  double adt[5]{}, bdt[5]{}, cdt[5]{}, gdt[5]{};
  int iadvance_time = 1, ntime = 1;
  int itime = 0, itr = 1;
  std::vector<double> ux, uy, uz, px, py, pz, pp3;
  std::vector<double> dux[3], duy[3], duz[3];
  std::vector<std::vector<double>> work;
  // wall pressure-gradient terms captured by gradp and used by pre_correc (navier.f90:439-496, 560-746)
  std::vector<double> dpdyx1, dpdzx1, dpdyxn, dpdzxn, dpdxy1, dpdzy1, dpdxyn, dpdzyn, dpdxz1, dpdyz1, dpdxzn, dpdyzn;
  // wall velocities (src/module_param.f90:242-244), zero unless a case sets them, in the order
  // bxx1 bxy1 bxz1 bxxn bxyn bxzn | byx1 byy1 byz1 byxn byyn byzn | bzx1 bzy1 bzz1 bzxn bzyn bzzn
  std::vector<double> bw[18];
  std::vector<double> bxo, byo, bzo;   // inflow noise planes (random_number in the reference; inputs here)
  // immersed boundary (iibm = 1: body() around the projection; 2 / 3: reconstruction inside the derivative operators,
  // src/derive.f90:23-24, and the masked velocity in divergence, src/navier.f90:285-293); geometry is an input
  int iibm = 0;
  std::vector<double> ep1;
  double ubc[3] = {0.0, 0.0, 0.0};
  struct IbmStore {
    IbmGeom g;
    std::vector<int> nobj, nipif, nfpif;
    std::vector<double> xi, xf;
    bool set = false;
  } ibm[3];
  void ibm_prepass(int axis, double *arr, double lind);
  void init_cyl();                     // Case-Cylinder-wake.f90:205-279 with iin = 0
  void inflow();                       // Case-Cylinder-wake.f90:100-133
  void outflow();                      // Case-Cylinder-wake.f90:135-203
  void init();
  void init_tgv();
  void init_channel();        // Case-Channel.f90:25-107, iin = 0 (laminar profile + deterministic perturbation)
  double fcpg = 0.0;          // parameters.f90:310
  void momentum_forcing(double *dux1, double *duy1, double *duz1);   // case.f90:538-567 -> momentum_forcing_channel
  void boundary_conditions(); // case.f90 boundary_conditions -> boundary_conditions_channel
  void channel_cfr(std::vector<double> &u, double constant);
  void capture_wall_gradients(const double *px1, const double *py1, const double *pz1);
  void step();  // one full time step (iadvance_time sub-steps)
  void momentum_rhs_eq(double *dux1, double *duy1, double *duz1);
  void intt(std::vector<double> &var, std::vector<double> *dvar);
  void pre_correc();
  void divergence(double *pp3out, int nlock, double *tmax, double *tmoy);
  void gradp(double *px1, double *py1, double *pz1, const double *pp3in);
  void cor_vel();
  void postprocess_tgv(double out[4]);  // eek, eps, eps2, enst
  double *W(int i, size_t n);
};

void channel_cfr_apply(double *u, int nx, int ny, int nz, const double *ppy, double dy, double yly, double constant);
// src/ibm.f90:14-80: the "old school" solid body (iibm = 1): velocity zeroed inside the body, bracketed by corgp_IBM
void ibm_body(double *ux, double *uy, double *uz, const double *ep, size_t n);
void ibm_corgp(double *ux, double *uy, double *uz, const double *px, const double *py, const double *pz, size_t n, int nlock);

}  // namespace x3do
