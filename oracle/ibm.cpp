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


// ---------------------------------------------------------------------------------------------------------------
// iibm = 3: cubic-spline reconstruction, cubsplx / cubsply / cubsplz and cubic_spline (src/ibm.f90:399-968).
// Differences from lagpol that are kept as they are in the reference: the imposed wall value `lind` (bcimp), ghost
// points with that value when a body touches a domain boundary, the analytic wall positions (ianal /= 0: passed in as
// ana_i / ana_f, the results of analitic_x / analitic_y), the equality test of a node with the wall position (x and y
// only), the "body spans the whole line" case (x and z only), and the value `y` that cubic_spline leaves untouched
// when x lies in none of its intervals (the caller's previous ypol is then stored again).

// natural ordering + clamped cubic spline, src/ibm.f90:880-968
static void cubic_spline(const double *xa, const double *ya, int n, double x, double &y) {
  double xaa[10] = {0}, yaa[10] = {0};
  int j = n / 2;
  for (int i = 1; i <= n; ++i) {
    if (i <= n / 2) { xaa[i - 1] = xa[j - 1]; yaa[i - 1] = ya[j - 1]; j = j - 1; }
    else { xaa[i - 1] = xa[i - 1]; yaa[i - 1] = ya[i - 1]; }
  }
  const double ypri = (yaa[2] - yaa[0]) / (xaa[2] - xaa[0]);
  const double yprf = (yaa[n - 1] - yaa[n - 3]) / (xaa[n - 1] - xaa[n - 3]);
  const int nk = n - 1, nc = nk - 1;
  double xx[10], aa[10], hh[10], alpha[10], ll[10], mm[10], zz[10], cc[10], bb[10], dd[10];   // 1-based below
  for (int i = 2; i <= nk; ++i) { aa[i - 1] = yaa[i - 1]; xx[i - 1] = xaa[i - 1]; }
  for (int i = 1; i <= nc - 1; ++i) hh[i] = xx[i + 1] - xx[i];
  alpha[1] = (3.0 * (aa[2] - aa[1])) / hh[1] - 3.0 * ypri;
  alpha[nc] = 3.0 * yprf - 3.0 * (aa[nc] - aa[nc - 1]) / hh[nc - 1];
  for (int i = 2; i <= nc - 1; ++i) alpha[i] = (3.0 / hh[i]) * (aa[i + 1] - aa[i]) - (3.0 / hh[i - 1]) * (aa[i] - aa[i - 1]);
  ll[1] = 2.0 * hh[1];
  mm[1] = 0.5;
  zz[1] = alpha[1] / ll[1];
  for (int i = 2; i <= nc - 1; ++i) {
    ll[i] = 2.0 * (xx[i + 1] - xx[i - 1]) - hh[i - 1] * mm[i - 1];
    mm[i] = hh[i] / ll[i];
    zz[i] = (alpha[i] - hh[i - 1] * zz[i - 1]) / ll[i];
  }
  ll[nc] = hh[nc - 1] * (2.0 - mm[nc - 1]);
  zz[nc] = (alpha[nc] - hh[nc - 1] * zz[nc - 1]) / ll[nc];
  cc[nc] = zz[nc];
  for (int q = nc - 1; q >= 1; --q) {
    cc[q] = zz[q] - mm[q] * cc[q + 1];
    bb[q] = (aa[q + 1] - aa[q]) / hh[q] - (hh[q] / 3.0) * (cc[q + 1] + 2.0 * cc[q]);
    dd[q] = (cc[q + 1] - cc[q]) / (3.0 * hh[q]);
  }
  for (int q = 2; q <= nc; ++q) {
    if (x <= xx[q] && x >= xx[q - 1]) {
      const double t = x - xx[q - 1];
      y = aa[q - 1] + bb[q - 1] * t + cc[q - 1] * (t * t) + dd[q - 1] * (t * t * t);
    }
  }
}

void cubspl(double *u, int nx, int ny, int nz, int axis, const IbmGeom &g, const double *coords, double d, double len, double lind,
            const double *ana_i, const double *ana_f) {
  const int n[3] = {nx, ny, nz};
  const int nl = n[axis];
  const int a_ax = axis == 0 ? 1 : 0, b_ax = axis == 2 ? 1 : 2;
  const int na = n[a_ax], nb = n[b_ax];
  const std::ptrdiff_t st[3] = {1, nx, static_cast<std::ptrdiff_t>(nx) * ny};
  const double bcimp = lind;
  double ypol = 0.0;   // a local of the reference routine that lives across lines
  for (int b = 0; b < nb; ++b)
    for (int a = 0; a < na; ++a) {
      const int nobj = g.nobj[a + static_cast<size_t>(na) * b];
      if (nobj == 0) continue;
      double *line = u + a * st[a_ax] + b * st[b_ax];
      auto U = [&](int q) -> double & { return line[(q - 1) * st[axis]]; };
      for (int i = 1; i <= nobj; ++i) {
        double xa[10] = {0}, ya[10] = {0};
        int ia = 0;
        const size_t gi = (i - 1) + static_cast<size_t>(g.nobjmax) * (a + static_cast<size_t>(na) * b);
        const size_t gp = i + static_cast<size_t>(g.nobjmax + 1) * (a + static_cast<size_t>(na) * b);
        const double xi = g.xi[gi], xf = g.xf[gi];
        const double ana_resi = ana_i ? ana_i[gi] : xi, ana_resf = ana_f ? ana_f[gi] : xf;
        int ipoli, ipolf, inxi = 0, inxf = 0;
        // ---- first wall
        int npf = g.npif;
        xa[ia] = ana_resi; ya[ia] = bcimp; ++ia;
        if (g.nipif[gp] < g.npif) npf = g.nipif[gp];
        if (xi > 0.0) {
          int ix;
          if (axis == 1) { ix = 1; while (coords[ix - 1] < xi) ix = ix + 1; ix = ix - 1; }
          else ix = static_cast<int>(xi / d + 1.0);
          ipoli = ix + 1;
          for (int ip = 1; ip <= npf; ++ip) {
            if (g.izap == 1) { xa[ia] = axis == 1 ? coords[ix - ip - 1] : static_cast<double>(ix - 1) * d - ip * d; ya[ia] = U(ix - ip); }
            else { xa[ia] = axis == 1 ? coords[ix - ip] : static_cast<double>(ix - 1) * d - (ip - 1) * d; ya[ia] = U(ix - ip + 1); }
            ++ia;
          }
        } else {  // the body starts on the domain boundary: ghost points carrying the wall value
          inxi = 1;
          int ix = 0;
          if (axis == 1) { ix = 1; while (coords[ix - 1] < xi) ix = ix + 1; ix = ix - 1; ipoli = ix + 1; }
          else { ix = static_cast<int>(xi / d); ipoli = axis == 0 ? ix + 1 : 1; }
          for (int ip = 1; ip <= npf; ++ip) {
            if (axis == 1) xa[ia] = g.izap == 1 ? coords[0] - (ip + 1) * d : coords[0] - (ip * d);
            else xa[ia] = g.izap == 1 ? static_cast<double>(ix - 1) * d - ip * d : static_cast<double>(ix - 1) * d - (ip - 1) * d;
            ya[ia] = bcimp;
            ++ia;
          }
        }
        // ---- second wall
        npf = g.npif;
        xa[ia] = ana_resf; ya[ia] = bcimp; ++ia;
        if (g.nfpif[gp] < g.npif) npf = g.nfpif[gp];
        if (xf < len) {
          int ix;
          if (axis == 1) { ix = 1; while (coords[ix - 1] < xf) ix = ix + 1; }
          else ix = static_cast<int>((xf + d) / d + 1.0);
          ipolf = ix - 1;
          for (int ip = 1; ip <= npf; ++ip) {
            if (g.izap == 1) { xa[ia] = axis == 1 ? coords[ix + ip - 1] : static_cast<double>(ix - 1) * d + ip * d; ya[ia] = U(ix + ip); }
            else { xa[ia] = axis == 1 ? coords[ix + ip - 2] : static_cast<double>(ix - 1) * d + (ip - 1) * d; ya[ia] = U(ix + ip - 1); }
            ++ia;
          }
        } else {
          inxf = 1;
          int ix;
          if (axis == 1) { ix = 1; while (ix <= nl && coords[ix - 1] < xf) ix = ix + 1; ipolf = ix - 1; }
          else { ix = static_cast<int>((xf + d) / d + 1.0); ipolf = axis == 0 ? ix - 1 : nl; }
          for (int ip = 1; ip <= npf; ++ip) {
            if (axis == 1) xa[ia] = g.izap == 1 ? coords[nl - 1] + (ip + 1) * d : coords[nl - 1] + ip * d;
            else xa[ia] = g.izap == 1 ? static_cast<double>(ix - 1) * d + ip * d : static_cast<double>(ix - 1) * d + (ip - 1) * d;
            ya[ia] = bcimp;
            ++ia;
          }
        }
        if (xi == xf) throw std::runtime_error("!! situation not supported by the IBM !!");
        const int na_pts = ia;
        for (int ipol = ipoli; ipol <= ipolf; ++ipol) {
          if (axis != 1 && inxf == 1 && inxi == 1) { U(ipol) = bcimp; continue; }
          const double xpol = axis == 1 ? coords[ipol - 1] : d * static_cast<double>(ipol - 1);
          if (axis != 2 && (xpol == ana_resi || xpol == ana_resf)) { U(ipol) = bcimp; continue; }
          cubic_spline(xa, ya, na_pts, xpol, ypol);
          U(ipol) = ypol;
        }
      }
    }
}

}  // namespace x3do
