// stretch.cpp -- stretched-mesh pieces of the oracle (TEST INFRASTRUCTURE): stretching_full
// (src/stretching.f90:96-318), matrice_refinement (src/poisson.f90:1814-2249) and inversion5_v1/v2
// (src/tools.f90:1225-1498).  Pinned by tests/golden/poisson.npz (outputs of the reference's own statements).
#include <cmath>
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {

// stretching.f90:96-318; s.istret, s.beta, s.yly are inputs
void stretching(Stretch &s, int ny, int nym, int ncly1, int nclyn, bool ncly) {
  (void)ncly1; (void)nclyn; (void)ncly;
  const int istret = s.istret;
  const double beta = s.beta, yly = s.yly;
  const double pi = std::acos(-1.0);
  vec yeta(ny, 0.0), yetai(ny, 0.0);
  for (vec *v : {&s.yp, &s.ypi, &s.ppy, &s.pp2y, &s.pp4y, &s.ppyi, &s.pp2yi, &s.pp4yi}) v->assign(ny, 0.0);
  const double yinf = -yly / 2.0;
  double den = 2.0 * beta * yinf;
  double xnum = -yinf - std::sqrt(pi * pi * beta * beta + yinf * yinf);
  const double alpha = std::fabs(xnum / den);
  s.alpha = alpha;
  if (alpha != 0.0) {
    if (istret == 1) { s.yp[0] = 0.0; yeta[0] = 0.0; }
    if (istret == 2) { s.yp[0] = 0.0; yeta[0] = -0.5; }
    if (istret == 3) { s.yp[0] = 0.0; yeta[0] = -0.5; }
    for (int j = 2; j <= ny; ++j) {
      double &ye = yeta[j - 1];
      if (istret == 1) ye = static_cast<double>(j - 1) * (1.0 / nym);
      if (istret == 2) ye = static_cast<double>(j - 1) * (1.0 / nym) - 0.5;
      if (istret == 3) ye = static_cast<double>(j - 1) * (0.5 / nym) - 0.5;
      const double den1 = std::sqrt(alpha * beta + 1.0);
      xnum = den1 / std::sqrt(alpha / pi) / std::sqrt(beta) / std::sqrt(pi);
      den = 2.0 * std::sqrt(alpha / pi) * std::sqrt(beta) * pi * std::sqrt(pi);
      const double den3 = ((std::sin(pi * ye)) * (std::sin(pi * ye)) / beta / pi) + alpha / pi;
      const double den4 = 2.0 * alpha * beta - std::cos(2.0 * pi * ye) + 1.0;
      const double xnum1 = (std::atan(xnum * std::tan(pi * ye))) * den4 / den1 / den3 / den;
      const double cst = std::sqrt(beta) * pi / (2.0 * std::sqrt(alpha) * std::sqrt(alpha * beta + 1.0));
      double &yp = s.yp[j - 1];
      if (istret == 1) {
        if (ye < 0.5) yp = xnum1 - cst - yinf;
        if (ye == 0.5) yp = 0.0 - yinf;
        if (ye > 0.5) yp = xnum1 + cst - yinf;
      }
      if (istret == 2) {
        if (ye < 0.5) yp = xnum1 - cst + yly;
        if (ye == 0.5) yp = 0.0 + yly;
        if (ye > 0.5) yp = xnum1 + cst + yly;
      }
      if (istret == 3) {
        if (ye < 0.5) yp = (xnum1 - cst + yly) * 2.0;
        if (ye == 0.5) yp = (0.0 + yly) * 2.0;
        if (ye > 0.5) yp = (xnum1 + cst + yly) * 2.0;
      }
    }
    for (int j = 1; j <= ny; ++j) {
      double &ye = yetai[j - 1];
      if (istret == 1) ye = (static_cast<double>(j) - 0.5) * (1.0 / nym);
      if (istret == 2) ye = (static_cast<double>(j) - 0.5) * (1.0 / nym) - 0.5;
      if (istret == 3) ye = (static_cast<double>(j) - 0.5) * (0.5 / nym) - 0.5;
      const double den1 = std::sqrt(alpha * beta + 1.0);
      xnum = den1 / std::sqrt(alpha / pi) / std::sqrt(beta) / std::sqrt(pi);
      den = 2.0 * std::sqrt(alpha / pi) * std::sqrt(beta) * pi * std::sqrt(pi);
      const double den3 = ((std::sin(pi * ye)) * (std::sin(pi * ye)) / beta / pi) + alpha / pi;
      const double den4 = 2.0 * alpha * beta - std::cos(2.0 * pi * ye) + 1.0;
      const double xnum1 = (std::atan(xnum * std::tan(pi * ye))) * den4 / den1 / den3 / den;
      const double cst = std::sqrt(beta) * pi / (2.0 * std::sqrt(alpha) * std::sqrt(alpha * beta + 1.0));
      double &yp = s.ypi[j - 1];
      if (istret == 1) {
        if (ye < 0.5) yp = xnum1 - cst - yinf;
        if (ye == 0.5) yp = 0.0 - yinf;
        if (ye > 0.5) yp = xnum1 + cst - yinf;
      }
      if (istret == 2) {
        if (ye < 0.5) yp = xnum1 - cst + yly;
        if (ye == 0.5) yp = 0.0 + yly;
        if (ye > 0.5) yp = xnum1 + cst + yly;
      }
      if (istret == 3) {
        if (ye < 0.5) yp = (xnum1 - cst + yly) * 2.0;
        if (ye == 0.5) yp = (0.0 + yly) * 2.0;
        if (ye > 0.5) yp = (xnum1 + cst + yly) * 2.0;
      }
    }
  } else {
    s.yp[0] = -1.e10;
    s.ypi[0] = -1.e10;
    for (int j = 2; j <= ny; ++j) {
      yeta[j - 1] = static_cast<double>(j - 1) * (1.0 / ny);
      s.yp[j - 1] = -beta * std::cos(pi * yeta[j - 1]) / std::sin(yeta[j - 1] * pi);
      yetai[j - 1] = static_cast<double>(j - 1) * (1.0 / ny);
      s.ypi[j - 1] = -beta * std::cos(pi * yetai[j - 1]) / std::sin(yetai[j - 1] * pi);
    }
  }
  // metric terms, :262-286
  const double half4 = (istret == 3) ? 2.0 : 1.0;
  for (int j = 0; j < ny; ++j) {
    s.ppy[j] = yly * (alpha / pi + (1.0 / pi / beta) * std::sin(pi * yeta[j]) * std::sin(pi * yeta[j]));
    s.pp2y[j] = s.ppy[j] * s.ppy[j];
    s.pp4y[j] = (-2.0 / beta * std::cos(pi * yeta[j]) * std::sin(pi * yeta[j]));
    if (istret == 3) s.pp4y[j] = s.pp4y[j] / half4;
    s.ppyi[j] = yly * (alpha / pi + (1.0 / pi / beta) * std::sin(pi * yetai[j]) * std::sin(pi * yetai[j]));
    s.pp2yi[j] = s.ppyi[j] * s.ppyi[j];
    s.pp4yi[j] = (-2.0 / beta * std::cos(pi * yetai[j]) * std::sin(pi * yetai[j]));
    if (istret == 3) s.pp4yi[j] = s.pp4yi[j] / half4;
  }
}

namespace {
inline cplx cx(double a, double b) { return cplx(a, b); }
inline double rl(cplx c) { return c.real(); }
inline double iy(cplx c) { return c.imag(); }
}  // namespace

// poisson.f90:1814-2249.  a, a2: (nx, ny/2, nzh, 5); a3: (nx, nym, nzh, 5), i fastest, band index slowest.
void Poisson::matrice_refinement() {
  const int NY = sy->n;                 // module variable ny (velocity nodes); ny/2 below is NY/2
  const int nym = sy->nm;
  const int nyh = NY / 2;
  const double dx = sx->d, dy = sy->d, dz = sz->d;
  const auto &cxx = sx->c; const auto &cy = sy->c; const auto &cz = sz->c;
  const double pi = std::acos(-1.0);
  const cplx one_one(1.0, 1.0);
  std::vector<double> trx(nx), trx2(nx), trY(ny), trY2(ny), tzr(nzh), tzi(nzh), tzr2(nzh), tzi2(nzh);
  std::vector<cplx> transz(nzh);
  for (int i = 0; i < nx; ++i) {  // :1860-1871
    const double e = rl(exs[i]) * dx;
    const double tt = 2.0 * (cxx.bici6 * std::cos(e * 1.5) + cxx.cici6 * std::cos(e * 2.5) + cxx.dici6 * std::cos(e * 3.5));
    const double tt1 = 2.0 * cxx.aici6 * std::cos(e * 0.5);
    const double t1 = 1.0 + 2.0 * cxx.ailcai6 * std::cos(e);
    trx[i] = (tt1 + tt) / t1;
    trx2[i] = trx[i] * trx[i];
  }
  for (int j = 0; j < ny; ++j) {  // :1873-1885
    const double e = rl(eys[j]) * dy;
    const double tt = 2.0 * (cy.bici6 * std::cos(e * 1.5) + cy.cici6 * std::cos(e * 2.5) + cy.dici6 * std::cos(e * 3.5));
    const double tt1 = 2.0 * cy.aici6 * std::cos(e * 0.5);
    const double t1 = 1.0 + 2.0 * cy.ailcai6 * std::cos(e);
    trY[j] = (tt1 + tt) / t1;
    trY2[j] = trY[j] * trY[j];
  }
  for (int k = 0; k < nzh; ++k) {
    if (bcz == 0) {  // :1888-1904
      const double e = rl(ezs[k]) * dz;
      const double tt = 2.0 * (cz.bici6 * std::cos(e * 1.5) + cz.cici6 * std::cos(e * 2.5) + cz.dici6 * std::cos(e * 3.5));
      const double tt1 = 2.0 * cz.aici6 * std::cos(e * 0.5);
      const double t1 = 1.0 + 2.0 * cz.ailcai6 * std::cos(e);
      tzr[k] = (tt1 + tt) / t1;
      tzr2[k] = tzr[k] * tzr[k];
      tzi[k] = tzr[k];
      tzi2[k] = tzr2[k];
      transz[k] = one_one * tzr[k];
    } else {  // :1906-1926 (no dici6 term here)
      const double er = rl(ezs[k]) * dz, ei = iy(ezs[k]) * dz;
      const cplx ztt = 2.0 * cx(cz.bici6 * std::cos(er * 1.5) + cz.cici6 * std::cos(er * 2.5),
                                cz.bici6 * std::cos(ei * 1.5) + cz.cici6 * std::cos(ei * 2.5));
      const cplx ztt1 = 2.0 * cx(cz.aici6 * std::cos(er * 0.5), cz.aici6 * std::cos(ei * 0.5));
      const cplx zt1 = cx(1.0 + 2.0 * cz.ailcai6 * std::cos(er), 1.0 + 2.0 * cz.ailcai6 * std::cos(ei));
      tzr[k] = rl(ztt1 + ztt) / rl(zt1);
      tzr2[k] = tzr[k] * tzr[k];
      tzi[k] = iy(ztt1 + ztt) / iy(zt1);
      tzi2[k] = tzi[k] * tzi[k];
      transz[k] = cx(tzr[k], tzi[k]);
    }
  }
  auto prod = [&](int i, int jy, int k) {  // transx_rl(i) * cx(rl(yky(jy)) rl(transz(k)), iy(yky(jy)) iy(transz(k))), jy 1-based
    return trx[i] * cx(rl(yky[jy - 1]) * rl(transz[k]), iy(yky[jy - 1]) * iy(transz[k]));
  };
  if (istret == 1 || istret == 2) {
    const double xa0 = alpha / pi + 0.5 / beta / pi;
    const double xa1 = (istret == 1) ? +1.0 / 4.0 / beta / pi : -1.0 / 4.0 / beta / pi;
    const double xa0_2 = xa0 * xa0, xa1_2 = xa1 * xa1, xa01 = xa0 * xa1, xa0p1_2 = (xa0 + xa1) * (xa0 + xa1);
    const size_t nb = static_cast<size_t>(nx) * nyh * nzh;
    a.assign(nb * 5, cplx(0.0, 0.0));
    a2.assign(nb * 5, cplx(0.0, 0.0));
    std::vector<cplx> c22(static_cast<size_t>(nx) * nyh * nzh), c2(static_cast<size_t>(nx) * nyh * nzh);
    auto I3 = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * ((j - 1) + static_cast<size_t>(nyh) * k); };  // j 1-based
    auto A = [&](std::vector<cplx> &m, int i, int j, int k, int b) -> cplx & { return m[I3(i, j, k) + nb * (b - 1)]; };
    for (int k = 0; k < nzh; ++k)
      for (int j = 1; j <= nyh; ++j)
        for (int i = 0; i < nx; ++i) {
          c22[I3(i, j, k)] = prod(i, 2 * j - 1, k);
          c2[I3(i, j, k)] = prod(i, 2 * j, k);
        }
    auto diag = [&](const std::vector<cplx> &cw, int i, int j, int k, int jy, double c0, bool lo, bool hi) {
      // -( xk2 ty2 tz2 + zk2 ty2 tx2 + c0 cw(j)^2 + xa1_2 cw(j) (cw(j-1) [lo] + cw(j+1) [hi]) ), component-wise
      const cplx w = cw[I3(i, j, k)];
      const cplx wm = lo ? cw[I3(i, j - 1, k)] : cplx(0.0, 0.0), wp = hi ? cw[I3(i, j + 1, k)] : cplx(0.0, 0.0);
      const double ty2 = trY2[jy - 1];
      double nr, ni;
      if (lo && hi) {
        nr = rl(xk2[i]) * ty2 * tzr2[k] + rl(zk2[k]) * ty2 * trx2[i] + c0 * rl(w) * rl(w) + xa1_2 * rl(w) * (rl(wm) + rl(wp));
        ni = iy(xk2[i]) * ty2 * tzi2[k] + iy(zk2[k]) * ty2 * trx2[i] + c0 * iy(w) * iy(w) + xa1_2 * iy(w) * (iy(wm) + iy(wp));
      } else {
        const cplx wn = lo ? wm : wp;
        nr = rl(xk2[i]) * ty2 * tzr2[k] + rl(zk2[k]) * ty2 * trx2[i] + c0 * rl(w) * rl(w) + xa1_2 * rl(w) * rl(wn);
        ni = iy(xk2[i]) * ty2 * tzi2[k] + iy(zk2[k]) * ty2 * trx2[i] + c0 * iy(w) * iy(w) + xa1_2 * iy(w) * iy(wn);
      }
      return -cx(nr, ni);
    };
    for (int k = 0; k < nzh; ++k) {
      for (int j = 2; j <= nyh - 1; ++j)  // main diagonal, :1953-1976
        for (int i = 0; i < nx; ++i) {
          A(a, i, j, k, 3) = diag(c22, i, j, k, 2 * j - 1, xa0_2, true, true);
          A(a2, i, j, k, 3) = diag(c2, i, j, k, 2 * j, xa0_2, true, true);
        }
      for (int i = 0; i < nx; ++i) {  // :1978-2020
        A(a, i, 1, k, 3) = diag(c22, i, 1, k, 1, xa0_2, false, true);
        A(a, i, nyh, k, 3) = diag(c22, i, nyh, k, NY - 2, xa0_2, true, false);
        A(a2, i, 1, k, 3) = diag(c2, i, 1, k, 2, xa0_2 - xa1_2, false, true);
        A(a2, i, nyh, k, 3) = diag(c2, i, nyh, k, NY - 1, xa0p1_2, true, false);
      }
    }
    auto cmul2 = [&](cplx p, cplx q) { return cx(rl(p) * rl(q), iy(p) * iy(q)); };   // component-wise product
    auto cadd = [&](cplx p, cplx q) { return cx(rl(p) + rl(q), iy(p) + iy(q)); };
    for (int k = 0; k < nzh; ++k) {  // sup diag +1, :2024-2051
      for (int j = 2; j <= nyh - 1; ++j)
        for (int i = 0; i < nx; ++i) {
          A(a, i, j, k, 4) = xa01 * cmul2(c22[I3(i, j + 1, k)], cadd(c22[I3(i, j, k)], c22[I3(i, j + 1, k)]));
          A(a2, i, j, k, 4) = xa01 * cmul2(c2[I3(i, j + 1, k)], cadd(c2[I3(i, j, k)], c2[I3(i, j + 1, k)]));
        }
      for (int i = 0; i < nx; ++i) {
        const cplx w1 = c22[I3(i, 1, k)], w2 = c22[I3(i, 2, k)];
        A(a, i, 1, k, 4) = 2.0 * xa01 * cx(rl(w1) * rl(w2) + rl(w2) * rl(w2), iy(w1) * iy(w2) + iy(w2) * iy(w2));
        const cplx v1 = c2[I3(i, 1, k)], v2 = c2[I3(i, 2, k)];
        A(a2, i, 1, k, 4) = cx((xa0 - xa1) * xa1 * (rl(v1) * rl(v2)) + xa0 * xa1 * (rl(v2) * rl(v2)),
                               (xa0 - xa1) * xa1 * (iy(v1) * iy(v2)) + xa0 * xa1 * (iy(v2) * iy(v2)));
        const cplx vm = c2[I3(i, nyh - 1, k)], vn = c2[I3(i, nyh, k)];
        A(a2, i, nyh - 1, k, 4) = cx(xa0 * xa1 * rl(vm) * rl(vn) + (xa0 + xa1) * xa1 * (rl(vn) * rl(vn)),
                                     xa0 * xa1 * iy(vm) * iy(vn) + (xa0 + xa1) * xa1 * (iy(vn) * iy(vn)));
        A(a2, i, nyh, k, 4) = 0.0;
      }
    }
    for (int k = 0; k < nzh; ++k)  // sup diag +2, :2054-2072
      for (int i = 0; i < nx; ++i) {
        for (int j = 1; j <= nyh - 2; ++j) {
          const cplx p = c22[I3(i, j + 1, k)], q = c22[I3(i, j + 2, k)];
          A(a, i, j, k, 5) = xa1_2 * cx(-rl(p) * rl(q), -iy(p) * iy(q));
          const cplx p2 = c2[I3(i, j + 1, k)], q2 = c2[I3(i, j + 2, k)];
          A(a2, i, j, k, 5) = xa1_2 * cx(-rl(p2) * rl(q2), -iy(p2) * iy(q2));
        }
        A(a, i, 1, k, 5) = 2.0 * cx(rl(A(a, i, 1, k, 5)), iy(A(a, i, 1, k, 5)));
        A(a, i, nyh - 1, k, 5) = 0.0; A(a, i, nyh, k, 5) = 0.0;
        A(a2, i, nyh - 1, k, 5) = 0.0; A(a2, i, nyh, k, 5) = 0.0;
      }
    for (int k = 0; k < nzh; ++k)  // inf diag -1, :2075-2103
      for (int i = 0; i < nx; ++i) {
        for (int j = 2; j <= nyh; ++j) {
          A(a, i, j, k, 2) = xa01 * cmul2(c22[I3(i, j - 1, k)], cadd(c22[I3(i, j, k)], c22[I3(i, j - 1, k)]));
          A(a2, i, j, k, 2) = xa01 * cmul2(c2[I3(i, j - 1, k)], cadd(c2[I3(i, j, k)], c2[I3(i, j - 1, k)]));
        }
        A(a, i, 1, k, 2) = 0.0; A(a2, i, 1, k, 2) = 0.0;
        const cplx v1 = c2[I3(i, 1, k)], v2 = c2[I3(i, 2, k)];
        A(a2, i, 2, k, 2) = cx(xa0 * xa1 * (rl(v2) * rl(v1)) + (xa0 + xa1) * xa1 * (rl(v1) * rl(v1)),
                               xa0 * xa1 * (iy(v2) * iy(v1)) + (xa0 + xa1) * xa1 * (iy(v1) * iy(v1)));
        const cplx vm = c2[I3(i, nyh - 1, k)], vn = c2[I3(i, nyh, k)];
        A(a2, i, nyh, k, 2) = cx((xa0 + xa1) * xa1 * (rl(vn) * rl(vm)) + xa0 * xa1 * (rl(vm) * rl(vm)),
                                 (xa0 + xa1) * xa1 * (iy(vn) * iy(vm)) + xa0 * xa1 * (iy(vm) * iy(vm)));
      }
    for (int k = 0; k < nzh; ++k)  // inf diag -2, :2105-2118
      for (int i = 0; i < nx; ++i) {
        for (int j = 3; j <= nyh; ++j) {
          const cplx p = c22[I3(i, j - 1, k)], q = c22[I3(i, j - 2, k)];
          A(a, i, j, k, 1) = xa1_2 * cx(-rl(p) * rl(q), -iy(p) * iy(q));
          const cplx p2 = c2[I3(i, j - 1, k)], q2 = c2[I3(i, j - 2, k)];
          A(a2, i, j, k, 1) = xa1_2 * cx(-rl(p2) * rl(q2), -iy(p2) * iy(q2));
        }
        A(a, i, 1, k, 1) = 0.0; A(a, i, 2, k, 1) = 0.0; A(a2, i, 1, k, 1) = 0.0; A(a2, i, 2, k, 1) = 0.0;
      }
    for (int k = 0; k < nzh; ++k)  // not to have a singular matrix, :2120-2129
      for (int i = 0; i < nx; ++i)
        if (rl(xk2[i]) == 0.0 && rl(zk2[k]) == 0.0) {
          A(a, i, 1, k, 3) = one_one;
          A(a, i, 1, k, 4) = 0.0;
          A(a, i, 1, k, 5) = 0.0;
        }
  } else {  // istret = 3, :2131-2246
    const double xa0 = alpha / pi + 0.5 / beta / pi;
    const double xa1 = -1.0 / 4.0 / beta / pi;
    const double xa0_2 = xa0 * xa0, xa1_2 = xa1 * xa1, xa01 = xa0 * xa1;
    const size_t nb = static_cast<size_t>(nx) * nym * nzh;
    a3.assign(nb * 5, cplx(0.0, 0.0));
    std::vector<cplx> c22(nb);
    auto I3 = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * ((j - 1) + static_cast<size_t>(nym) * k); };
    auto A = [&](int i, int j, int k, int b) -> cplx & { return a3[I3(i, j, k) + nb * (b - 1)]; };
    for (int k = 0; k < nzh; ++k)
      for (int j = 1; j <= nym; ++j)
        for (int i = 0; i < nx; ++i) c22[I3(i, j, k)] = prod(i, j, k);
    for (int k = 0; k < nzh; ++k)
      for (int i = 0; i < nx; ++i) {
        for (int j = 1; j <= nym; ++j) {  // main diagonal, :2156-2198
          const cplx w = c22[I3(i, j, k)];
          const double ty2 = trY2[j - 1];
          double nr = rl(xk2[i]) * ty2 * tzr2[k] + rl(zk2[k]) * ty2 * trx2[i] + xa0_2 * rl(w) * rl(w);
          double ni = iy(xk2[i]) * ty2 * tzi2[k] + iy(zk2[k]) * ty2 * trx2[i] + xa0_2 * iy(w) * iy(w);
          if (j == 1) {
            nr += xa1_2 * rl(w) * rl(c22[I3(i, 2, k)]); ni += xa1_2 * iy(w) * iy(c22[I3(i, 2, k)]);
          } else if (j == nym) {
            nr += xa1_2 * rl(w) * rl(c22[I3(i, nym - 1, k)]); ni += xa1_2 * iy(w) * iy(c22[I3(i, nym - 1, k)]);
          } else {
            nr += xa1_2 * rl(w) * (rl(c22[I3(i, j - 1, k)]) + rl(c22[I3(i, j + 1, k)]));
            ni += xa1_2 * iy(w) * (iy(c22[I3(i, j - 1, k)]) + iy(c22[I3(i, j + 1, k)]));
          }
          A(i, j, k, 3) = -cx(nr, ni);
        }
        for (int j = 1; j <= nym - 1; ++j) {  // sup diag +1, :2201-2211 (row nym stays 0)
          const cplx p = c22[I3(i, j + 1, k)], w = c22[I3(i, j, k)];
          A(i, j, k, 4) = xa01 * cx(rl(p) * (rl(w) + rl(p)), iy(p) * (iy(w) + iy(p)));
        }
        for (int j = 1; j <= nym - 2; ++j) {  // sup diag +2, :2214-2223
          const cplx p = c22[I3(i, j + 1, k)], q = c22[I3(i, j + 2, k)];
          A(i, j, k, 5) = -xa1_2 * cx(rl(p) * rl(q), iy(p) * iy(q));
        }
        A(i, nym - 1, k, 5) = 0.0; A(i, nym, k, 5) = 0.0;
        for (int j = 2; j <= nym; ++j) {  // inf diag -1, :2226-2234
          const cplx p = c22[I3(i, j - 1, k)], w = c22[I3(i, j, k)];
          A(i, j, k, 2) = xa01 * cx(rl(p) * (rl(w) + rl(p)), iy(p) * (iy(w) + iy(p)));
        }
        A(i, 1, k, 2) = 0.0;
        for (int j = 3; j <= nym; ++j) {  // inf diag -2, :2237-2247
          const cplx p = c22[I3(i, j - 1, k)], q = c22[I3(i, j - 2, k)];
          A(i, j, k, 1) = -xa1_2 * cx(rl(p) * rl(q), iy(p) * iy(q));
        }
        A(i, 1, k, 1) = 0.0; A(i, 2, k, 1) = 0.0;
      }
    // :2240-2245 (rank 0 only, element (1,1,1))
    A(0, 1, 0, 3) = one_one;
    A(0, 1, 0, 4) = 0.0;
    A(0, 1, 0, 5) = 0.0;
  }
}

// tools.f90:1225-1360 (v1: works on a copy) and :1365-1498 (v2: eliminates in place); n = rows (ny/2 or nym).
// The elimination multipliers keep their previous value when a pivot is exactly zero (tmp1/tmp2 are only
// assigned under `if (pivot /= zero)`); loop order k outer, i inner as in the reference.
static void inversion5(cplx *aaa, cplx *eee, int nx, int n, int nz) {
  const double epsilon = 1.e-16;
  const size_t nb = static_cast<size_t>(nx) * n * nz;
  auto I3 = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * ((j - 1) + static_cast<size_t>(n) * k); };
  auto A = [&](int i, int j, int k, int b) -> cplx & { return aaa[I3(i, j, k) + nb * (b - 1)]; };
  auto E = [&](int i, int j, int k) -> cplx & { return eee[I3(i, j, k)]; };
  double tmp1 = 0.0, tmp2 = 0.0, tmp3 = 0.0, tmp4 = 0.0;
  std::vector<cplx> sr(static_cast<size_t>(nx) * nz);
  for (int m = 1; m <= n - 2; ++m)
    for (int ii = 1; ii <= 2; ++ii) {
      const int mi = m + ii;
      for (int k = 0; k < nz; ++k)
        for (int j = 0; j < nx; ++j) {
          if (A(j, m, k, 3).real() != 0.0) tmp1 = A(j, mi, k, 3 - ii).real() / A(j, m, k, 3).real();
          if (A(j, m, k, 3).imag() != 0.0) tmp2 = A(j, mi, k, 3 - ii).imag() / A(j, m, k, 3).imag();
          sr[j + static_cast<size_t>(nx) * k] = cplx(tmp1, tmp2);
          E(j, mi, k) = cplx(E(j, mi, k).real() - tmp1 * E(j, m, k).real(), E(j, mi, k).imag() - tmp2 * E(j, m, k).imag());
        }
      for (int jc = 4 - ii; jc <= 5 - ii; ++jc)
        for (int k = 0; k < nz; ++k)
          for (int j = 0; j < nx; ++j) {
            const cplx s = sr[j + static_cast<size_t>(nx) * k];
            A(j, mi, k, jc) = cplx(A(j, mi, k, jc).real() - s.real() * A(j, m, k, jc + ii).real(),
                                   A(j, mi, k, jc).imag() - s.imag() * A(j, m, k, jc + ii).imag());
          }
    }
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < nx; ++j) {
      const cplx p = A(j, n - 1, k, 3);
      tmp1 = std::fabs(p.real()) > epsilon ? A(j, n, k, 2).real() / p.real() : 0.0;
      tmp2 = std::fabs(p.imag()) > epsilon ? A(j, n, k, 2).imag() / p.imag() : 0.0;
      const cplx s(tmp1, tmp2);
      cplx b1(A(j, n, k, 3).real() - tmp1 * A(j, n - 1, k, 4).real(), A(j, n, k, 3).imag() - tmp2 * A(j, n - 1, k, 4).imag());
      if (std::fabs(b1.real()) > epsilon) {
        tmp1 = s.real() / b1.real();
        tmp3 = E(j, n, k).real() / b1.real() - tmp1 * E(j, n - 1, k).real();
      } else { tmp1 = 0.0; tmp3 = 0.0; }
      if (std::fabs(b1.imag()) > epsilon) {
        tmp2 = s.imag() / b1.imag();
        tmp4 = E(j, n, k).imag() / b1.imag() - tmp2 * E(j, n - 1, k).imag();
      } else { tmp2 = 0.0; tmp4 = 0.0; }
      E(j, n, k) = cplx(tmp3, tmp4);
      tmp1 = std::fabs(p.real()) > epsilon ? 1.0 / p.real() : 0.0;
      tmp2 = std::fabs(p.imag()) > epsilon ? 1.0 / p.imag() : 0.0;
      b1 = cplx(tmp1, tmp2);
      const cplx a1(A(j, n - 1, k, 4).real() * b1.real(), A(j, n - 1, k, 4).imag() * b1.imag());
      E(j, n - 1, k) = cplx(E(j, n - 1, k).real() * b1.real() - a1.real() * E(j, n, k).real(),
                            E(j, n - 1, k).imag() * b1.imag() - a1.imag() * E(j, n, k).imag());
    }
  for (int i = n - 2; i >= 1; --i)
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < nx; ++j) {
        const cplx p = A(j, i, k, 3);
        tmp1 = std::fabs(p.real()) > epsilon ? 1.0 / p.real() : 0.0;
        tmp2 = std::fabs(p.imag()) > epsilon ? 1.0 / p.imag() : 0.0;
        const cplx a1(A(j, i, k, 4).real() * tmp1, A(j, i, k, 4).imag() * tmp2);
        const cplx b1(A(j, i, k, 5).real() * tmp1, A(j, i, k, 5).imag() * tmp2);
        E(j, i, k) = cplx(E(j, i, k).real() * tmp1 - a1.real() * E(j, i + 1, k).real() - b1.real() * E(j, i + 2, k).real(),
                          E(j, i, k).imag() * tmp2 - a1.imag() * E(j, i + 1, k).imag() - b1.imag() * E(j, i + 2, k).imag());
      }
}

void inversion5_v1(const std::vector<cplx> &aaa_in, cplx *eee, int nx, int nyh, int nz) {
  std::vector<cplx> aaa = aaa_in;  // tools.f90:1258
  inversion5(aaa.data(), eee, nx, nyh, nz);
}
void inversion5_v2(std::vector<cplx> &aaa, cplx *eee, int nx, int nym, int nz) { inversion5(aaa.data(), eee, nx, nym, nz); }

}  // namespace x3do
