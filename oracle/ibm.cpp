// ibm.cpp -- oracle restatement (TEST INFRASTRUCTURE) of the immersed-boundary pre-pass of the operators:
// lagpolx / lagpoly / lagpolz and polint (src/ibm.f90:83-343, 345-389).  When iibm = 2 every derx/dery/derz and
// derxx/deryy/derzz first rebuilds its INPUT inside the solid bodies by Lagrange interpolation through the body
// boundary (value 0) and up to npif fluid points on each side (src/derive.f90:23).  Pinned by tests/golden/ibm.npz.
#include <cmath>
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {

// Neville's algorithm, src/ibm.f90:345-389 (xa, ya: n points; returns the value at x)
static double polint(const double *xa, const double *ya, int n, double x) {
  double c[30], d[30];
  int ns = 1;
  double dif = std::fabs(x - xa[0]);
  for (int i = 1; i <= n; ++i) {
    const double dift = std::fabs(x - xa[i - 1]);
    if (dift < dif) { ns = i; dif = dift; }
    c[i - 1] = ya[i - 1];
    d[i - 1] = ya[i - 1];
  }
  double y = ya[ns - 1];
  ns = ns - 1;
  for (int m = 1; m <= n - 1; ++m) {
    for (int i = 1; i <= n - m; ++i) {
      const double ho = xa[i - 1] - x, hp = xa[i + m - 1] - x;
      const double w = c[i] - d[i - 1];
      double den = ho - hp;
      den = w / den;
      d[i - 1] = hp * den;
      c[i - 1] = ho * den;
    }
    double dy;
    if (2 * ns < n - m) dy = c[ns];
    else { dy = d[ns - 1]; ns = ns - 1; }
    y = y + dy;
  }
  return y;
}

// one direction; the line runs along `axis` of u(nx,ny,nz); geometry arrays are indexed by the two other indices in
// the order of the reference (x: (j,k), y: (i,k), z: (i,j)).  coords = node coordinates along the line; uniform
// directions locate the body faces by division as the reference does (lagpolx/z), y searches yp (lagpoly).
void lagpol(double *u, int nx, int ny, int nz, int axis, const IbmGeom &g, const double *coords, double d, double len) {
  const int n[3] = {nx, ny, nz};
  const int nl = n[axis];
  const int a_ax = axis == 0 ? 1 : 0, b_ax = axis == 2 ? 1 : 2;
  const int na = n[a_ax], nb = n[b_ax];
  const std::ptrdiff_t st[3] = {1, nx, static_cast<std::ptrdiff_t>(nx) * ny};
  for (int b = 0; b < nb; ++b)
    for (int a = 0; a < na; ++a) {
      const int nobj = g.nobj[a + static_cast<size_t>(na) * b];
      if (nobj == 0) continue;
      double *line = u + a * st[a_ax] + b * st[b_ax];
      auto U = [&](int q) -> double & { return line[(q - 1) * st[axis]]; };  // 1-based along the line
      auto X = [&](int q) { return axis == 1 ? coords[q - 1] : static_cast<double>(q - 1) * d; };
      for (int i = 1; i <= nobj; ++i) {
        double xa[10], ya[10];
        int ia = 0;
        const size_t gi = (i - 1) + static_cast<size_t>(g.nobjmax) * (a + static_cast<size_t>(na) * b);
        const size_t gp = i + static_cast<size_t>(g.nobjmax + 1) * (a + static_cast<size_t>(na) * b);
        const double xi = g.xi[gi], xf = g.xf[gi];
        int ipoli, ipolf;
        // first face
        int npf = g.npif;
        xa[ia] = xi; ya[ia] = 0.0; ++ia;
        if (xi > 0.0) {
          int ix;
          if (axis == 1) { ix = 1; while (coords[ix - 1] < xi) ix = ix + 1; ix = ix - 1; }
          else ix = static_cast<int>(xi / d + 1.0);
          ipoli = ix + 1;
          if (g.nipif[gp] < g.npif) npf = g.nipif[gp];
          for (int ip = 1; ip <= npf; ++ip) {
            if (g.izap == 1) { xa[ia] = axis == 1 ? coords[ix - ip - 1] : static_cast<double>(ix - 1) * d - ip * d; ya[ia] = U(ix - ip); }
            else { xa[ia] = axis == 1 ? coords[ix - ip] : static_cast<double>(ix - 1) * d - (ip - 1) * d; ya[ia] = U(ix - ip + 1); }
            ++ia;
          }
        } else {
          ipoli = 1;
        }
        // second face
        npf = g.npif;
        xa[ia] = xf; ya[ia] = 0.0; ++ia;
        if (xf < len) {
          int ix;
          if (axis == 1) { ix = 1; while (coords[ix - 1] < xf) ix = ix + 1; }
          else ix = static_cast<int>((xf + d) / d + 1.0);
          ipolf = ix - 1;
          if (g.nfpif[gp] < g.npif) npf = g.nfpif[gp];
          for (int ip = 1; ip <= npf; ++ip) {
            if (g.izap == 1) { xa[ia] = axis == 1 ? coords[ix + ip - 1] : static_cast<double>(ix - 1) * d + ip * d; ya[ia] = U(ix + ip); }
            else { xa[ia] = axis == 1 ? coords[ix + ip - 2] : static_cast<double>(ix - 1) * d + (ip - 1) * d; ya[ia] = U(ix + ip - 1); }
            ++ia;
          }
        } else {
          ipolf = nl;
        }
        const int na_pts = ia;
        for (int ipol = ipoli; ipol <= ipolf; ++ipol) U(ipol) = polint(xa, ya, na_pts, X(ipol));
      }
    }
}

}  // namespace x3do
