// x3d_schemes.cuh -- host-side compact-scheme coefficients for the device-resident solver
// (product code; mirrors what the Fortran host computes in schemes(), src/schemes.f90).
#pragma once
#include <vector>
#include "x3d_common.cuh"

namespace x3d {

struct LU3 {  // one prepared tridiagonal: upper f, multipliers s, inverse pivots w (schemes.f90:413-439)
  std::vector<double> f, s, w;
};

struct AxisCoeffs {
  int n = 0, nm = 0, ncl1 = 0, ncln = 0;
  bool periodic = false;
  double d = 0, len = 0;
  x3d_deriv_coeffs c{};
  LU3 d1, d1p;        // ffx.. / ffxp..
  LU3 d2, d2p;        // sfx.. / sfxp..
  LU3 vp, vpp;        // cfx6.. / cfxp6..   (size nm)
  LU3 ivp, ivpp;      // cifx6.. / cifxp6.. (size nm)
  LU3 pv, pvp;        // cfi6.. / cfip6..   (size n)
  LU3 ipv, ipvp;      // cifi6.. / cifip6.. (size n)
};

struct SchemeOpts {
  int ifirstder = 4, isecondder = 4, ipinter = 3;
  double nu0nu = 4.0, cnu = 0.44;
};

AxisCoeffs make_axis_coeffs(int n, int ncl1, int ncln, double len, const SchemeOpts &o);

// stretched y mesh of stretching() (src/stretching.f90:96-318) for hosts that do not bring their own
struct StretchY {
  int istret = 0;
  double beta = 0.0, alpha = 0.0;
  std::vector<double> yp, ypi, ppy, pp2y, pp4y, ppyi, pp2yi, pp4yi;
};
StretchY make_stretching(int istret, double beta, double yly, int ny, int nym);

// set_filter_coefficients (src/filters.f90:62-219)
void make_filter_axis(int n, int ncl1, int ncln, double af, x3d_filter_coeffs &c, LU3 &plain, LU3 &p);

}  // namespace x3d
