// x3d_schemes.cu -- coefficients of the sixth-order compact schemes and the LU of their
// tridiagonal left-hand sides, as schemes() builds them (src/schemes.f90:413-1066).  Host code,
// O(n); used by the device-resident solver (a Fortran host passes its own arrays instead).
#include <cmath>
#include "x3d_schemes.cuh"

namespace x3d {

namespace {

struct Tri {  // lower b, diagonal c, upper f
  std::vector<double> b, c, f;
  explicit Tri(int n) : b(n, 0.0), c(n, 0.0), f(n, 0.0) {}
};

LU3 factor(const Tri &t) {  // prepare(), schemes.f90:413-439
  const int n = static_cast<int>(t.c.size());
  LU3 r;
  r.f = t.f;
  r.s.assign(n, 0.0);
  r.w = t.c;
  for (int i = 1; i < n; ++i) {
    r.s[i] = t.b[i - 1] / r.w[i - 1];
    r.w[i] = r.w[i] - t.f[i - 1] * r.s[i];
  }
  for (int i = 0; i < n; ++i) r.w[i] = 1.0 / r.w[i];
  return r;
}

// uniform interior band alpha,1,alpha
Tri band(int n, double al) {
  Tri t(n);
  for (int i = 0; i < n; ++i) { t.b[i] = al; t.c[i] = 1.0; t.f[i] = al; }
  t.f[n - 1] = 0.0;
  t.b[n - 1] = 0.0;
  return t;
}

}  // namespace

AxisCoeffs make_axis_coeffs(int n, int ncl1, int ncln, double len, const SchemeOpts &o) {
  AxisCoeffs A;
  A.n = n; A.ncl1 = ncl1; A.ncln = ncln; A.len = len;
  A.periodic = (ncl1 == 0 && ncln == 0);
  A.nm = A.periodic ? n : n - 1;               // parameters.f90:273-294
  const double d = len / static_cast<double>(A.nm);
  A.d = d;
  const double d2 = d * d;
  auto &c = A.c;
  // ---- stencil scalars ---------------------------------------------------------------
  if (o.ifirstder == 1) { c.alfai = 0.0; c.afi = 1.0 / (2.0 * d); c.bfi = 0.0; }             // schemes.f90:467-470
  else if (o.ifirstder == 4) { c.alfai = 1.0 / 3.0; c.afi = (7.0 / 9.0) / d; c.bfi = (1.0 / 36.0) / d; }  // :477-480
  else throw Error("ifirstder must be 1 or 4 (schemes.f90:471-486)");
  if (o.ifirstder != 1) {                                                                     // :505-519
    c.alfa1 = 2.0; c.af1 = -(5.0 / 2.0) / d; c.bf1 = 2.0 / d; c.cf1 = 0.5 / d; c.df1 = 0.0;
    c.alfa2 = 0.25; c.af2 = 0.75 / d;
    c.alfan = 2.0; c.afn = -(5.0 / 2.0) / d; c.bfn = 2.0 / d; c.cfn = 0.5 / d; c.dfn = 0.0;
    c.alfam = 0.25; c.afm = 0.75 / d;
  }
  if (o.isecondder == 1) { c.alsai = 0.0; c.asi = 1.0 / d2; }                                 // :636-641
  else if (o.isecondder == 4) { c.alsai = 2.0 / 11.0; c.asi = (12.0 / 11.0) / d2; c.bsi = (3.0 / 44.0) / d2; }  // :658-663
  else if (o.isecondder == 5) {                                                               // :674-688
    const double pi = std::acos(-1.0);
    const double dpis3 = 2.0 * pi / 3.0;
    const double xnpi2 = pi * pi * (1.0 + o.nu0nu);
    const double xmpi2 = dpis3 * dpis3 * (1.0 + o.cnu * o.nu0nu);
    const double den = 405.0 * xnpi2 - 640.0 * xmpi2 + 144.0;
    c.alsai = 0.5 - (320.0 * xmpi2 - 1296.0) / den;
    c.asi = -(4329.0 * xnpi2 / 8.0 - 32.0 * xmpi2 - 140.0 * xnpi2 * xmpi2 + 286.0) / den / d2;
    c.bsi = (2115.0 * xnpi2 - 1792.0 * xmpi2 - 280.0 * xnpi2 * xmpi2 + 1328.0) / den / (4.0 * d2);
    c.csi = -(7695.0 * xnpi2 / 8.0 + 288.0 * xmpi2 - 180.0 * xnpi2 * xmpi2 - 2574.0) / den / (9.0 * d2);
    c.dsi = (198.0 * xnpi2 + 128.0 * xmpi2 - 40.0 * xnpi2 * xmpi2 - 736.0) / den / (16.0 * d2);
  } else throw Error("isecondder must be 1, 4 or 5");
  c.alsa1 = 11.0; c.as1 = 13.0 / d2; c.bs1 = -27.0 / d2; c.cs1 = 15.0 / d2; c.ds1 = -1.0 / d2;  // :696-700
  c.alsa2 = (o.isecondder == 1) ? 0.0 : 0.1; c.as2 = (o.isecondder == 1) ? 1.0 / d2 : (6.0 / 5.0) / d2;
  c.alsa3 = 2.0 / 11.0; c.as3 = (12.0 / 11.0) / d2; c.bs3 = (3.0 / 44.0) / d2;
  c.alsa4 = 2.0 / 11.0; c.as4 = (12.0 / 11.0) / d2; c.bs4 = (3.0 / 44.0) / d2; c.cs4 = 0.0;
  c.alsan = c.alsa1; c.asn = c.as1; c.bsn = c.bs1; c.csn = c.cs1; c.dsn = c.ds1;                  // :719-723
  c.alsam = c.alsa2; c.asm_ = c.as2;
  c.alsat = c.alsa3; c.ast = c.as3; c.bst = c.bs3;
  c.alsatt = c.alsa4; c.astt = c.as4; c.bstt = c.bs4; c.cstt = 0.0;
  if (o.ifirstder == 1) { c.alcai6 = 0.0; c.aci6 = 1.0 / d; c.bci6 = 0.0; }                      // :891-899
  else { c.alcai6 = 9.0 / 62.0; c.aci6 = (63.0 / 62.0) / d; c.bci6 = (17.0 / 62.0) / 3.0 / d; }
  if (o.ifirstder == 1) { c.ailcai6 = 0.0; c.aici6 = 0.5; }                                      // :901-931
  else if (o.ipinter == 1) { c.ailcai6 = 0.3; c.aici6 = 0.75; c.bici6 = 1.0 / 20.0; }
  else if (o.ipinter == 2) {
    c.ailcai6 = 0.461658; c.dici6 = 0.00293016;
    c.aici6 = 1.0 / 64.0 * (75.0 + 70.0 * c.ailcai6 - 320.0 * c.dici6);
    c.bici6 = 1.0 / 128.0 * (126.0 * c.ailcai6 - 25.0 + 1152.0 * c.dici6);
    c.cici6 = 1.0 / 128.0 * (-10.0 * c.ailcai6 + 3.0 - 640.0 * c.dici6);
    c.aici6 /= 2.0; c.bici6 /= 2.0; c.cici6 /= 2.0; c.dici6 /= 2.0;
  } else if (o.ipinter == 3) {
    c.ailcai6 = 0.49;
    c.aici6 = 1.0 / 128.0 * (75.0 + 70.0 * c.ailcai6);
    c.bici6 = 1.0 / 256.0 * (126.0 * c.ailcai6 - 25.0);
    c.cici6 = 1.0 / 256.0 * (-10.0 * c.ailcai6 + 3.0);
  } else throw Error("ipinter must be 1, 2 or 3");
  if (n == 1) return A;
  if (n < 10) throw Error("compact schemes need at least 10 points per direction");
  // ---- first derivative LHS, :524-596 ------------------------------------------------------
  {
    const double al = c.alfai;
    Tri t = band(n, al);
    if (ncl1 == 0) t.c[0] = 2.0;
    else if (ncl1 == 1) t.f[0] = al + al;
    else { t.f[0] = c.alfa1; t.f[1] = c.alfa2; t.b[0] = c.alfa2; }
    if (ncln == 0) t.c[n - 1] = 1.0 + al * al;
    else if (ncln == 1) t.b[n - 2] = al + al;
    else { t.f[n - 2] = c.alfam; t.b[n - 3] = c.alfam; t.b[n - 2] = c.alfan; }
    A.d1 = factor(t);
    if (ncl1 == 1) t.f[0] = 0.0;
    if (ncln == 1) t.b[n - 2] = 0.0;
    A.d1p = factor(t);
  }
  // ---- second derivative LHS, :744-853 -----------------------------------------------------
  {
    const double al = c.alsai;
    Tri t = band(n, al);
    if (ncl1 == 0) t.c[0] = 2.0;
    else if (ncl1 == 1) t.f[0] = al + al;
    else { t.f[0] = c.alsa1; t.f[1] = c.alsa2; t.f[2] = c.alsa3; t.f[3] = c.alsa4; t.b[0] = c.alsa2; t.b[1] = c.alsa3; t.b[2] = c.alsa4; }
    if (ncln == 0) t.c[n - 1] = 1.0 + al * al;
    else if (ncln == 1) t.b[n - 2] = al + al;
    else {
      t.f[n - 4] = c.alsatt; t.f[n - 3] = c.alsat; t.f[n - 2] = c.alsam;
      t.b[n - 5] = c.alsatt; t.b[n - 4] = c.alsat; t.b[n - 3] = c.alsam; t.b[n - 2] = c.alsan;
    }
    A.d2p = factor(t);                    // sfxp: keeps sf(1)=2 alpha, sb(n-1)=2 alpha  (:848)
    if (ncl1 == 1) t.f[0] = 0.0;          // :843-845
    if (ncln == 1) t.b[n - 2] = 0.0;      // :850-853
    A.d2 = factor(t);
    if (ncl1 != 1 && ncln != 1) A.d2 = A.d2p;
    else if (ncln != 1) { /* only f(0) zeroed: already done above */ }
    // NOTE (:847-853): ss/sw are first prepared with sb(n-1) intact; they are re-prepared with
    // sb(n-1)=0 only when ncln==1, which is what the two lines above reproduce.
  }
  // ---- staggered LHS, :935-1063 --------------------------------------------------------------
  auto stag = [&](double al, LU3 &vp, LU3 &vpp, LU3 &pv, LU3 &pvp, bool deriv) {
    const int nm = A.nm;
    Tri tm = band(nm, al);  // pressure-mesh sized
    tm.c[0] = (ncl1 == 0) ? 2.0 : 1.0 + al;
    tm.c[nm - 1] = (ncln == 0) ? 1.0 + al * al : 1.0 + al;
    vp = factor(tm);
    Tri tmp = tm;
    if (deriv) tmp.f[0] = 0.0;            // cfxp6(1)=0 (:1034); cifxp6 is an unmodified copy
    vpp = factor(tmp);
    Tri tn = band(n, al);   // velocity-mesh sized
    tn.f[0] = al + al;
    tn.b[n - 2] = al + al;
    pv = factor(tn);
    Tri tnp = tn;
    if (deriv) tnp.f[0] = 0.0;            // cfip6(1)=0 (:1035)
    pvp = factor(tnp);
    if (ncln == 1 || ncln == 2) {         // :1044-1063
      if (deriv) { tmp.b[nm - 2] = 0.0; vpp = factor(tmp); }   // cbx6(nxm-1)=0
      // cibx6(nxm)=0 is already zero: cisxp6/ciwxp6 unchanged
      if (deriv) { tnp.b[n - 2] = 0.0; pvp = factor(tnp); }    // cbi6(nx-1)=0
      // cibi6(nx)=0 already zero
    }
  };
  stag(c.alcai6, A.vp, A.vpp, A.pv, A.pvp, true);
  stag(c.ailcai6, A.ivp, A.ivpp, A.ipv, A.ipvp, false);
  return A;
}

// src/stretching.f90:96-318: node (yp) and mid-point (ypi) coordinates of the mapped mesh and the metric
// factors that multiply the y derivatives
StretchY make_stretching(int istret, double beta, double yly, int ny, int nym) {
  if (istret < 1 || istret > 3 || !(beta > 0.0)) throw Error("stretching: istret must be 1..3 and beta > 0");
  StretchY S;
  S.istret = istret; S.beta = beta;
  const double pi = std::acos(-1.0);
  for (auto *v : {&S.yp, &S.ypi, &S.ppy, &S.pp2y, &S.pp4y, &S.ppyi, &S.pp2yi, &S.pp4yi}) v->assign(ny, 0.0);
  const double yinf = -yly / 2.0;
  const double alpha = std::fabs((-yinf - std::sqrt(pi * pi * beta * beta + yinf * yinf)) / (2.0 * beta * yinf));
  S.alpha = alpha;
  if (alpha == 0.0) throw Error("stretching: alpha = 0 is not supported");
  const double scale = (istret == 3) ? 0.5 / nym : 1.0 / nym;
  const double shift = (istret == 1) ? 0.0 : -0.5;
  // mapped coordinate of eta, :126-152 / :165-190
  auto coord = [&](double eta) {
    const double den1 = std::sqrt(alpha * beta + 1.0);
    const double xnum = den1 / std::sqrt(alpha / pi) / std::sqrt(beta) / std::sqrt(pi);
    const double den = 2.0 * std::sqrt(alpha / pi) * std::sqrt(beta) * pi * std::sqrt(pi);
    const double den3 = ((std::sin(pi * eta)) * (std::sin(pi * eta)) / beta / pi) + alpha / pi;
    const double den4 = 2.0 * alpha * beta - std::cos(2.0 * pi * eta) + 1.0;
    const double xnum1 = (std::atan(xnum * std::tan(pi * eta))) * den4 / den1 / den3 / den;
    const double cst = std::sqrt(beta) * pi / (2.0 * std::sqrt(alpha) * std::sqrt(alpha * beta + 1.0));
    const double off = (istret == 1) ? -yinf : yly;
    double y = 0.0;
    if (eta < 0.5) y = xnum1 - cst + off;
    if (eta == 0.5) y = 0.0 + off;
    if (eta > 0.5) y = xnum1 + cst + off;
    return (istret == 3) ? y * 2.0 : y;
  };
  std::vector<double> eta(ny), etai(ny);
  eta[0] = (istret == 1) ? 0.0 : -0.5;
  S.yp[0] = 0.0;
  for (int j = 1; j < ny; ++j) {
    eta[j] = static_cast<double>(j) * scale + shift;
    S.yp[j] = coord(eta[j]);
  }
  for (int j = 0; j < ny; ++j) {
    etai[j] = (static_cast<double>(j + 1) - 0.5) * scale + shift;
    S.ypi[j] = coord(etai[j]);
  }
  for (int j = 0; j < ny; ++j) {  // :262-286
    S.ppy[j] = yly * (alpha / pi + (1.0 / pi / beta) * std::sin(pi * eta[j]) * std::sin(pi * eta[j]));
    S.pp2y[j] = S.ppy[j] * S.ppy[j];
    S.pp4y[j] = (-2.0 / beta * std::cos(pi * eta[j]) * std::sin(pi * eta[j]));
    S.ppyi[j] = yly * (alpha / pi + (1.0 / pi / beta) * std::sin(pi * etai[j]) * std::sin(pi * etai[j]));
    S.pp2yi[j] = S.ppyi[j] * S.ppyi[j];
    S.pp4yi[j] = (-2.0 / beta * std::cos(pi * etai[j]) * std::sin(pi * etai[j]));
    if (istret == 3) { S.pp4y[j] /= 2.0; S.pp4yi[j] /= 2.0; }
  }
  return S;
}

// set_filter_coefficients (src/filters.f90:62-219): the parfiX/Y/Z scalars for the filter parameter af and the two
// prepared left-hand sides (plain: fiffx,fifsx,fifwx; p: fiffxp,fifsxp,fifwxp)
void make_filter_axis(int n, int ncl1, int ncln, double af, x3d_filter_coeffs &c, LU3 &plain, LU3 &p) {
  if (n < 6) throw Error("x3d_filter_axis: n must be at least 6");
  if (ncl1 < 0 || ncl1 > 2 || ncln < 0 || ncln > 2) throw Error("x3d_filter_axis: bad boundary condition");
  c = x3d_filter_coeffs{};
  // interior, Gaitonde & Visbal 1998 (:87-93)
  c.fiali = af;
  c.fiai = (11.0 + 10.0 * af) / 16.0;
  c.fibi = 0.5 * (15.0 + 34.0 * af) / 32.0;
  c.fici = 0.5 * (-3.0 + 6.0 * af) / 16.0;
  c.fidi = 0.5 * (1.0 - 2.0 * af) / 32.0;
  // boundary points 1 / n: not filtered; 2 / n-1: third order; 3 / n-2: fifth order (:94-138)
  c.fial1 = 0.0; c.fia1 = 1.0; c.fib1 = 0.0; c.fic1 = 0.0; c.fid1 = 0.0;
  c.fial2 = af; c.fia2 = 1.0 / 8.0 + 3.0 / 4.0 * af; c.fib2 = 5.0 / 8.0 + 3.0 / 4.0 * af; c.fic2 = 3.0 / 8.0 + af / 4.0; c.fid2 = -1.0 / 8.0 + af / 4.0;
  c.fial3 = af; c.fia3 = -1.0 / 32.0 + af / 16.0; c.fib3 = 5.0 / 32.0 + 11.0 / 16.0 * af; c.fic3 = 11.0 / 16.0 + 5.0 * af / 8.0;
  c.fid3 = 5.0 / 16.0 + 3.0 * af / 8.0; c.fie3 = -5.0 / 32.0 + 5.0 * af / 16.0; c.fif3 = 1.0 / 32.0 - af / 16.0;
  c.fialn = 0.0; c.fian = 1.0; c.fibn = 0.0; c.ficn = 0.0; c.fidn = 0.0;
  c.fialm = c.fial2; c.fiam = c.fia2; c.fibm = c.fib2; c.ficm = c.fic2; c.fidm = c.fid2;
  c.fialp = c.fial3; c.fiap = c.fia3; c.fibp = c.fib3; c.ficp = c.fic3; c.fidp = c.fid3; c.fiep = c.fie3; c.fifp = c.fif3;
  // tridiagonal left-hand side, :140-196 (0-based rows)
  Tri t(n);
  const double al = c.fiali;
  if (ncl1 == 0) { t.f[0] = al; t.f[1] = al; t.c[0] = 2.0; t.c[1] = 1.0; t.b[0] = al; t.b[1] = al; }
  else if (ncl1 == 1) { t.f[0] = al + al; t.f[1] = al; t.c[0] = 1.0; t.c[1] = 1.0; t.b[0] = al; t.b[1] = al; }
  else { t.f[0] = c.fial1; t.f[1] = c.fial2; t.c[0] = 1.0; t.c[1] = 1.0; t.b[0] = c.fial2; t.b[1] = al; }
  for (int i = 2; i <= n - 4; ++i) { t.f[i] = al; t.c[i] = 1.0; t.b[i] = al; }   // do i = 3, n-3 (after the ends in the reference; disjoint rows)
  if (ncln == 0) { t.f[n - 3] = al; t.f[n - 2] = al; t.f[n - 1] = 0.0; t.c[n - 3] = 1.0; t.c[n - 2] = 1.0; t.c[n - 1] = 1.0 + al * al;
                   t.b[n - 3] = al; t.b[n - 2] = al; t.b[n - 1] = 0.0; }
  else if (ncln == 1) { t.f[n - 3] = al; t.f[n - 2] = al; t.f[n - 1] = 0.0; t.c[n - 3] = 1.0; t.c[n - 2] = 1.0; t.c[n - 1] = 1.0;
                        t.b[n - 3] = al; t.b[n - 2] = al + al; t.b[n - 1] = 0.0; }
  else { t.f[n - 3] = al; t.f[n - 2] = c.fialm; t.f[n - 1] = 0.0; t.c[n - 3] = 1.0; t.c[n - 2] = 1.0; t.c[n - 1] = 1.0;
         t.b[n - 3] = c.fialm; t.b[n - 2] = c.fialn; t.b[n - 1] = 0.0; }
  p = factor(t);                       // prepare(fb, fc, ffp, fsp, fwp, n), :203
  if (ncl1 == 1) t.f[0] = 0.0;         // :205-210
  if (ncln == 1) t.b[n - 2] = 0.0;
  plain = factor(t);                   // prepare(fb, fc, ff, fs, fw, n)
}

}  // namespace x3d
