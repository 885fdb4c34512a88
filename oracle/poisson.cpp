// poisson.cpp -- oracle restatement of src/poisson.f90 on one rank (TEST INFRASTRUCTURE).
// All pencils coincide on one rank, so every transpose_* of the reference is the identity;
// arrays are (nx,ny,nz) with i fastest, spectral arrays (nx,ny,nz/2+1).
#include <cmath>
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {

namespace {
const double EPS = 1.e-16;  // poisson.f90:25

// decomp_2d_fft_3d, PHYSICAL_IN_Z: r2c along z, c2c along y, c2c along x (poisson.f90:330)
void fft3d_r2c(const double *in, cplx *out, int nx, int ny, int nz) {
  const int nzh = nz / 2 + 1;
  const std::ptrdiff_t nxy = static_cast<std::ptrdiff_t>(nx) * ny;
#pragma omp parallel
  {
    std::vector<cplx> line(nz);
#pragma omp for schedule(static)
    for (std::ptrdiff_t p = 0; p < nxy; ++p) {
      for (int k = 0; k < nz; ++k) line[k] = cplx(in[p + nxy * k], 0.0);
      fft_line(line.data(), nz, 1, -1);
      for (int k = 0; k < nzh; ++k) out[p + nxy * k] = line[k];
    }
#pragma omp for collapse(2) schedule(static)
    for (int k = 0; k < nzh; ++k)
      for (int i = 0; i < nx; ++i) fft_line(out + i + nxy * k, ny, nx, -1);
#pragma omp for collapse(2) schedule(static)
    for (int k = 0; k < nzh; ++k)
      for (int j = 0; j < ny; ++j) fft_line(out + nx * j + nxy * k, nx, 1, -1);
  }
}
// backward: c2c x, c2c y, c2r z (poisson.f90:405); unnormalised
void fft3d_c2r(cplx *in, double *out, int nx, int ny, int nz) {
  const int nzh = nz / 2 + 1;
  const std::ptrdiff_t nxy = static_cast<std::ptrdiff_t>(nx) * ny;
#pragma omp parallel
  {
    std::vector<cplx> line(nz);
#pragma omp for collapse(2) schedule(static)
    for (int k = 0; k < nzh; ++k)
      for (int j = 0; j < ny; ++j) fft_line(in + nx * j + nxy * k, nx, 1, +1);
#pragma omp for collapse(2) schedule(static)
    for (int k = 0; k < nzh; ++k)
      for (int i = 0; i < nx; ++i) fft_line(in + i + nxy * k, ny, nx, +1);
#pragma omp for schedule(static)
    for (std::ptrdiff_t p = 0; p < nxy; ++p) {
      for (int k = 0; k < nzh; ++k) line[k] = in[p + nxy * k];
      for (int k = nzh; k < nz; ++k) line[k] = std::conj(in[p + nxy * (nz - k)]);
      fft_line(line.data(), nz, 1, +1);
      for (int k = 0; k < nz; ++k) out[p + nxy * k] = line[k].real();
    }
  }
}

inline cplx cx(double a, double b) { return cplx(a, b); }

// even/odd reordering of a non-periodic axis before the FFT (poisson.f90:435-444)
void reorder_fwd(const double *in, double *out, int nx, int ny, int nz, int axis) {
  const int n[3] = {nx, ny, nz};
  const int na = n[axis];
  const std::ptrdiff_t st[3] = {1, nx, static_cast<std::ptrdiff_t>(nx) * ny};
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int idx[3] = {i, j, k};
        const int q = idx[axis];
        const int src = (q < na / 2) ? 2 * q : 2 * na - 2 * q - 1;
        const std::ptrdiff_t base = i * st[0] + j * st[1] + k * st[2];
        out[base] = in[base + (src - q) * st[axis]];
      }
}
// inverse reordering after the inverse FFT (poisson.f90:643-652)
void reorder_bwd(const double *in, double *out, int nx, int ny, int nz, int axis) {
  const int n[3] = {nx, ny, nz};
  const int na = n[axis];
  const std::ptrdiff_t st[3] = {1, nx, static_cast<std::ptrdiff_t>(nx) * ny};
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int idx[3] = {i, j, k};
        const int q = idx[axis];
        // out(2*i'-1) = in(i'), out(2*i') = in(n-i'+1)  (1-based)
        const int src = (q % 2 == 0) ? q / 2 : na - 1 - (q - 1) / 2;
        const std::ptrdiff_t base = i * st[0] + j * st[1] + k * st[2];
        out[base] = in[base + (src - q) * st[axis]];
      }
}
}  // namespace

// poisson.f90:1469-1526
void Poisson::abxyz() {
  const double PI = std::acos(-1.0);
  auto fill = [&](vec &a, vec &b, int n, int bc) {
    a.resize(n); b.resize(n);
    for (int i = 0; i < n; ++i) {
      const double arg = (bc == 0) ? static_cast<double>(i) * PI / static_cast<double>(n)
                                   : static_cast<double>(i) * PI * 0.5 / static_cast<double>(n);
      a[i] = std::sin(arg);
      b[i] = std::cos(arg);
    }
  };
  fill(ax, bx, nx, bcx);
  fill(ay, by, ny, bcy);
  fill(az, bz, nz, bcz);
}

// poisson.f90:1530-1810
void Poisson::waves() {
  const double pi = std::acos(-1.0), twopi = 2.0 * std::acos(-1.0);
  const int NX = sx->n, NY = sy->n, NZ = sz->n;        // module variables nx,ny,nz (velocity nodes)
  const int nxm = sx->nm, nym = sy->nm, nzm = sz->nm;
  const double dx = sx->d, dy = sy->d, dz = sz->d;
  const double xlx = sx->len, yly = sy->len, zlz = sz->len;
  const auto &cxx = sx->c; const auto &cy = sy->c; const auto &cz = sz->c;
  const cplx one_one(1.0, 1.0);
  xkx.assign(NX, 0.0); xk2.assign(NX, 0.0); exs.assign(NX, 0.0);
  yky.assign(NY, 0.0); yk2.assign(NY, 0.0); eys.assign(NY, 0.0);
  zkz.assign(NZ / 2 + 1, 0.0); zk2.assign(NZ / 2 + 1, 0.0); ezs.assign(NZ / 2 + 1, 0.0);
  // x, :1566-1596
  if (bcx == 0) {
    for (int i = 1; i <= NX / 2 + 1; ++i) {
      const double w = twopi * (i - 1) / NX;
      double wp = cxx.aci6 * 2.0 * dx * std::sin(w * 0.5) + cxx.bci6 * 2.0 * dx * std::sin(3.0 * 0.5 * w);
      wp = wp / (1.0 + 2.0 * cxx.alcai6 * std::cos(w));
      xkx[i - 1] = one_one * (NX * wp / xlx);
      exs[i - 1] = one_one * (NX * w / xlx);
      xk2[i - 1] = one_one * ((NX * wp / xlx) * (NX * wp / xlx));
    }
    for (int i = NX / 2 + 2; i <= NX; ++i) {
      xkx[i - 1] = xkx[NX - i + 1]; exs[i - 1] = exs[NX - i + 1]; xk2[i - 1] = xk2[NX - i + 1];
    }
  } else {
    for (int i = 1; i <= NX; ++i) {
      const double w = twopi * 0.5 * (i - 1) / nxm;
      double wp = cxx.aci6 * 2.0 * dx * std::sin(w * 0.5) + (cxx.bci6 * 2.0 * dx) * std::sin(3.0 * 0.5 * w);
      wp = wp / (1.0 + 2.0 * cxx.alcai6 * std::cos(w));
      xkx[i - 1] = one_one * static_cast<double>(nxm) * wp / xlx;
      exs[i - 1] = one_one * static_cast<double>(nxm) * w / xlx;
      xk2[i - 1] = one_one * ((nxm * wp / xlx) * (nxm * wp / xlx));
    }
    xkx[0] = 0.0; exs[0] = 0.0; xk2[0] = 0.0;
  }
  // y, :1599-1631
  if (bcy == 0) {
    for (int j = 1; j <= NY / 2 + 1; ++j) {
      const double w = twopi * (j - 1) / NY;
      double wp = cy.aci6 * 2.0 * dy * std::sin(w * 0.5) + cy.bci6 * 2.0 * dy * std::sin(3.0 * 0.5 * w);
      wp = wp / (1.0 + 2.0 * cy.alcai6 * std::cos(w));
      yky[j - 1] = (istret == 0) ? one_one * (NY * wp / yly) : one_one * (NY * wp);
      eys[j - 1] = one_one * (NY * w / yly);
      yk2[j - 1] = one_one * ((NY * wp / yly) * (NY * wp / yly));
    }
    for (int j = NY / 2 + 2; j <= NY; ++j) {
      yky[j - 1] = yky[NY - j + 1]; eys[j - 1] = eys[NY - j + 1]; yk2[j - 1] = yk2[NY - j + 1];
    }
  } else {
    for (int j = 1; j <= NY; ++j) {
      const double w = twopi * 0.5 * (j - 1) / nym;
      double wp = cy.aci6 * 2.0 * dy * std::sin(w * 0.5) + (cy.bci6 * 2.0 * dy) * std::sin(3.0 * 0.5 * w);
      wp = wp / (1.0 + 2.0 * cy.alcai6 * std::cos(w));
      yky[j - 1] = (istret == 0) ? one_one * (nym * wp / yly) : one_one * (nym * wp);
      eys[j - 1] = one_one * (nym * w / yly);
      yk2[j - 1] = one_one * ((nym * wp / yly) * (nym * wp / yly));
    }
    yky[0] = 0.0; eys[0] = 0.0; yk2[0] = 0.0;
  }
  // z, :1634-1659
  if (bcz == 0) {
    for (int k = 1; k <= NZ / 2 + 1; ++k) {
      const double w = twopi * (k - 1) / NZ;
      double wp = cz.aci6 * 2.0 * dz * std::sin(w * 0.5) + (cz.bci6 * 2.0 * dz) * std::sin(3.0 * 0.5 * w);
      wp = wp / (1.0 + 2.0 * cz.alcai6 * std::cos(w));
      zkz[k - 1] = one_one * (NZ * wp / zlz);
      ezs[k - 1] = one_one * (NZ * w / zlz);
      zk2[k - 1] = one_one * ((NZ * wp / zlz) * (NZ * wp / zlz));
    }
  } else {
    for (int k = 1; k <= NZ / 2 + 1; ++k) {
      const double w = pi * (k - 1) / nzm;
      const double w1 = pi * (nzm - k + 1) / nzm;
      double wp = cz.aci6 * 2.0 * dz * std::sin(w * 0.5) + (cz.bci6 * 2.0 * dz) * std::sin(3.0 * 0.5 * w);
      wp = wp / (1.0 + 2.0 * cz.alcai6 * std::cos(w));
      double w1p = cz.aci6 * 2.0 * dz * std::sin(w1 * 0.5) + (cz.bci6 * 2.0 * dz) * std::sin(3.0 * 0.5 * w1);
      w1p = w1p / (1.0 + 2.0 * cz.alcai6 * std::cos(w1));
      zkz[k - 1] = cx(nzm * wp / zlz, -nzm * w1p / zlz);
      ezs[k - 1] = cx(nzm * w / zlz, nzm * w1 / zlz);
      zk2[k - 1] = cx((nzm * wp / zlz) * (nzm * wp / zlz), (nzm * w1p / zlz) * (nzm * w1p / zlz));
    }
  }
  // kxyz, :1661-1807 (the y-pencil branch :1661-1700 holds the same formulas as :1703-1742)
  kxyz.assign(static_cast<size_t>(nx) * ny * nzh, 0.0);
  for (int k = 0; k < nzh; ++k) {
    const double rlezs = ezs[k].real() * dz, iyezs = ezs[k].imag() * dz;
    for (int j = 0; j < ny; ++j) {
      const double rleys = eys[j].real() * dy;
      for (int i = 0; i < nx; ++i) {
        const double rlexs = exs[i].real() * dx;
        const double xtt_rl = 2.0 * (cxx.bici6 * std::cos(rlexs * 1.5) + cxx.cici6 * std::cos(rlexs * 2.5) + cxx.dici6 * std::cos(rlexs * 3.5));
        const double ytt_rl = 2.0 * (cy.bici6 * std::cos(rleys * 1.5) + cy.cici6 * std::cos(rleys * 2.5) + cy.dici6 * std::cos(rleys * 3.5));
        const double xtt1_rl = 2.0 * cxx.aici6 * std::cos(rlexs * 0.5);
        const double ytt1_rl = 2.0 * cy.aici6 * std::cos(rleys * 0.5);
        const double xt1_rl = 1.0 + 2.0 * cxx.ailcai6 * std::cos(rlexs);
        const double yt1_rl = 1.0 + 2.0 * cy.ailcai6 * std::cos(rleys);
        cplx xyzk;
        if (bcz == 0) {
          const double ztt_rl = 2.0 * (cz.bici6 * std::cos(rlezs * 1.5) + cz.cici6 * std::cos(rlezs * 2.5) + cz.dici6 * std::cos(rlezs * 3.5));
          const double ztt1_rl = 2.0 * cz.aici6 * std::cos(rlezs * 0.5);
          const double zt1_rl = 1.0 + 2.0 * cz.ailcai6 * std::cos(rlezs);
          const double fy = (ytt1_rl + ytt_rl) / yt1_rl, fz = (ztt1_rl + ztt_rl) / zt1_rl, fx = (xtt1_rl + xtt_rl) / xt1_rl;
          const cplx xt2 = xk2[i] * ((fy * fz) * (fy * fz));
          const cplx yt2 = yk2[j] * ((fx * fz) * (fx * fz));
          const cplx zt2 = zk2[k] * ((fx * fy) * (fx * fy));
          xyzk = xt2 + yt2 + zt2;
        } else {
          const cplx ztt = 2.0 * cx(cz.bici6 * std::cos(rlezs * 1.5) + cz.cici6 * std::cos(rlezs * 2.5) + cz.dici6 * std::cos(rlezs * 3.5),
                                    cz.bici6 * std::cos(iyezs * 1.5) + cz.cici6 * std::cos(iyezs * 2.5) + cz.dici6 * std::cos(iyezs * 3.5));
          const cplx ztt1 = 2.0 * cx(cz.aici6 * std::cos(rlezs * 0.5), cz.aici6 * std::cos(iyezs * 0.5));
          const cplx zt1 = cx(1.0 + 2.0 * cz.ailcai6 * std::cos(rlezs), 1.0 + 2.0 * cz.ailcai6 * std::cos(iyezs));
          const cplx t1 = cx((ztt1 + ztt).real() / zt1.real(), (ztt1 + ztt).imag() / zt1.imag());
          const double t2 = (ytt1_rl + ytt_rl) / yt1_rl;
          const double t3 = (xtt1_rl + xtt_rl) / xt1_rl;
          const cplx t4 = (t2 * t2) * cx(t1.real() * t1.real(), t1.imag() * t1.imag());
          const cplx t5 = (t3 * t3) * cx(t1.real() * t1.real(), t1.imag() * t1.imag());
          const double t6 = (t3 * t2) * (t3 * t2);
          const cplx a1 = cx(t4.real() * xk2[i].real(), t4.imag() * xk2[i].imag());
          const cplx a2 = cx(t5.real() * yk2[j].real(), t5.imag() * yk2[j].imag());
          const cplx a3 = t6 * zk2[k];
          xyzk = a1 + a2 + a3;
        }
        kxyz[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k)] = xyzk;
      }
    }
  }
}

// decomp_2d_poisson_init, poisson.f90:73-255
void Poisson::init(const AxisScheme &x, const AxisScheme &y, const AxisScheme &z, const Stretch *stp) {
  sx = &x; sy = &y; sz = &z; st = stp;
  bcx = x.periodic ? 0 : 1; bcy = y.periodic ? 0 : 1; bcz = z.periodic ? 0 : 1;
  nx = x.nm; ny = y.nm; nz = z.nm;
  nzh = nz / 2 + 1;
  istret = stp ? stp->istret : 0;
  if (stp) { alpha = stp->alpha; beta = stp->beta; }
  if (!((bcx == 0 && bcy == 0 && bcz == 0) || (bcx == 1 && bcy == 0 && bcz == 0) || (bcx == 0 && bcy == 1 && bcz == 0) ||
        (bcx == 1 && bcy == 1)))
    throw std::runtime_error("boundary condition not supported (poisson.f90:107)");
  abxyz();
  waves();
  if (bcy == 1 && istret != 0) matrice_refinement();
}

void Poisson::solve(double *rhs) {
  if (bcx == 0 && bcy == 0 && bcz == 0) solve_000(rhs);
  else if (bcx == 1 && bcy == 0 && bcz == 0) solve_100(rhs);
  else if (bcx == 0 && bcy == 1 && bcz == 0) solve_010(rhs);
  else solve_11x(rhs);
}

#define CW(a, i, j, k) a[(i) + static_cast<size_t>(nx) * ((j) + static_cast<size_t>(ny) * (k))]

namespace {
// z rotation forward (poisson.f90:343-346) / backward of the non-periodic solvers (:625-628)
inline cplx rot_fwd(cplx c, double a, double b) { return cx(c.real() * b + c.imag() * a, c.imag() * b - c.real() * a); }
inline cplx rot_bwd(cplx c, double a, double b) { return cx(c.real() * b - c.imag() * a, c.imag() * b + c.real() * a); }
// four-way guarded division (poisson.f90:545-559)
inline cplx divide4(cplx c, cplx kk) {
  const double t1 = kk.real(), t2 = kk.imag();
  const bool z1 = std::fabs(t1) < EPS, z2 = std::fabs(t2) < EPS;
  if (z1 && z2) return cx(0.0, 0.0);
  if (z1) return cx(0.0, c.imag() / (-t2));
  if (z2) return cx(c.real() / (-t1), 0.0);
  return cx(c.real() / (-t1), c.imag() / (-t2));
}
}  // namespace

// poisson.f90:298-410
void Poisson::solve_000(double *rhs) {
  std::vector<cplx> cw1(static_cast<size_t>(nx) * ny * nzh);
  fft3d_r2c(rhs, cw1.data(), nx, ny, nz);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        cplx c = CW(cw1, i, j, k);
        c = c / static_cast<double>(nx) / static_cast<double>(ny) / static_cast<double>(nz);  // :333
        c = rot_fwd(c, az[k], bz[k]);
        c = rot_fwd(c, ay[j], by[j]);
        if (j + 1 > ny / 2 + 1) c = -c;
        c = rot_fwd(c, ax[i], bx[i]);
        if (i + 1 > nx / 2 + 1) c = -c;
        const double t1 = CW(kxyz, i, j, k).real(), t2 = CW(kxyz, i, j, k).imag();
        if (t1 < EPS || t2 < EPS) c = 0.0;  // :366
        else c = cx(c.real() / (-t1), c.imag() / (-t2));
        // backward, :381-398
        c = cx(c.real() * bz[k] - c.imag() * az[k], -c.imag() * bz[k] - c.real() * az[k]);
        c = cx(c.real() * by[j] + c.imag() * ay[j], c.imag() * by[j] - c.real() * ay[j]);
        if (j + 1 > ny / 2 + 1) c = -c;
        c = cx(c.real() * bx[i] + c.imag() * ax[i], -c.imag() * bx[i] + c.real() * ax[i]);
        if (i + 1 > nx / 2 + 1) c = -c;
        CW(cw1, i, j, k) = c;
      }
  fft3d_c2r(cw1.data(), rhs, nx, ny, nz);
}

namespace {
// half-sample DCT post-processing along one axis with the mirrored partner (poisson.f90:508-528)
// out(1)=in(1); out(i) = [half*] cx(xx1+xx4+xx5-xx8, -xx2+xx3+xx6+xx7)
template <bool HalfFirst>
inline cplx dct_post(cplx c, cplx p, double a, double b) {
  double xx1 = c.real() * b, xx2 = c.real() * a, xx3 = c.imag() * b, xx4 = c.imag() * a;
  double xx5 = p.real() * b, xx6 = p.real() * a, xx7 = p.imag() * b, xx8 = p.imag() * a;
  if (HalfFirst) {  // poisson_11x multiplies each product by half (:1146-1153)
    xx1 *= 0.5; xx2 *= 0.5; xx3 *= 0.5; xx4 *= 0.5; xx5 *= 0.5; xx6 *= 0.5; xx7 *= 0.5; xx8 *= 0.5;
    return cx(xx1 + xx4 + xx5 - xx8, -xx2 + xx3 + xx6 + xx7);
  }
  return 0.5 * cx(xx1 + xx4 + xx5 - xx8, -xx2 + xx3 + xx6 + xx7);  // :524-525
}
// backward (poisson.f90:575-588): cx(xx1-xx4+xx6+xx7, -(-xx2-xx3+xx5-xx8))
inline cplx dct_pre(cplx c, cplx p, double a, double b) {
  const double xx1 = c.real() * b, xx2 = c.real() * a, xx3 = c.imag() * b, xx4 = c.imag() * a;
  const double xx5 = p.real() * b, xx6 = p.real() * a, xx7 = p.imag() * b, xx8 = p.imag() * a;
  return cx(xx1 - xx4 + xx6 + xx7, -(-xx2 - xx3 + xx5 - xx8));
}
}  // namespace

// poisson.f90:413-659
void Poisson::solve_100(double *rhs) {
  const size_t nr = static_cast<size_t>(nx) * ny * nz, nsp = static_cast<size_t>(nx) * ny * nzh;
  std::vector<double> rw(nr);
  std::vector<cplx> cw1(nsp), cw1b(nsp);
  reorder_fwd(rhs, rw.data(), nx, ny, nz, 0);
  fft3d_r2c(rw.data(), cw1.data(), nx, ny, nz);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        cplx c = CW(cw1, i, j, k);
        c = c / static_cast<double>(nx) / static_cast<double>(ny) / static_cast<double>(nz);
        c = rot_fwd(c, az[k], bz[k]);
        c = rot_fwd(c, ay[j], by[j]);
        if (j + 1 > ny / 2 + 1) c = -c;
        CW(cw1, i, j, k) = c;
      }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j) {
      CW(cw1b, 0, j, k) = CW(cw1, 0, j, k);
      for (int i = 1; i < nx; ++i) CW(cw1b, i, j, k) = dct_post<false>(CW(cw1, i, j, k), CW(cw1, nx - i, j, k), ax[i], bx[i]);
      for (int i = 0; i < nx; ++i) CW(cw1b, i, j, k) = divide4(CW(cw1b, i, j, k), CW(kxyz, i, j, k));
      CW(cw1, 0, j, k) = CW(cw1b, 0, j, k);
      for (int i = 1; i < nx; ++i) CW(cw1, i, j, k) = dct_pre(CW(cw1b, i, j, k), CW(cw1b, nx - i, j, k), ax[i], bx[i]);
      for (int i = 0; i < nx; ++i) {
        cplx c = CW(cw1, i, j, k);
        c = rot_bwd(c, ay[j], by[j]);  // :610-612
        if (j + 1 > ny / 2 + 1) c = -c;
        c = rot_bwd(c, az[k], bz[k]);  // :627-628
        CW(cw1, i, j, k) = c;
      }
    }
  fft3d_c2r(cw1.data(), rw.data(), nx, ny, nz);
  reorder_bwd(rw.data(), rhs, nx, ny, nz, 0);
}

// poisson.f90:665-1013
void Poisson::solve_010(double *rhs) {
  const size_t nr = static_cast<size_t>(nx) * ny * nz, nsp = static_cast<size_t>(nx) * ny * nzh;
  std::vector<double> rw(nr);
  std::vector<cplx> cw1(nsp), cw2b(nsp);
  reorder_fwd(rhs, rw.data(), nx, ny, nz, 1);
  fft3d_r2c(rw.data(), cw1.data(), nx, ny, nz);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        cplx c = CW(cw1, i, j, k);
        c = c / static_cast<double>(nx) / static_cast<double>(ny) / static_cast<double>(nz);
        c = rot_fwd(c, az[k], bz[k]);
        c = rot_fwd(c, ax[i], bx[i]);
        if (i + 1 > nx / 2 + 1) c = -c;
        CW(cw1, i, j, k) = c;
      }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int i = 0; i < nx; ++i) {
      CW(cw2b, i, 0, k) = CW(cw1, i, 0, k);
      for (int j = 1; j < ny; ++j) CW(cw2b, i, j, k) = dct_post<false>(CW(cw1, i, j, k), CW(cw1, i, ny - j, k), ay[j], by[j]);
    }
  if (istret == 0) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nzh; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) CW(cw2b, i, j, k) = divide4(CW(cw2b, i, j, k), CW(kxyz, i, j, k));
  } else {
    matrice_refinement();  // :822 (rebuilt every solve)
    const int nyh = ny / 2;
    if (istret != 3) {
      std::vector<cplx> c2(static_cast<size_t>(nx) * nyh * nzh), c2c(static_cast<size_t>(nx) * nyh * nzh);
      for (int k = 0; k < nzh; ++k)
        for (int j = 0; j < nyh; ++j)
          for (int i = 0; i < nx; ++i) {
            c2[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)] = CW(cw2b, i, 2 * j, k);
            c2c[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)] = CW(cw2b, i, 2 * j + 1, k);
          }
      inversion5_v1(a, c2.data(), nx, nyh, nzh);
      inversion5_v1(a2, c2c.data(), nx, nyh, nzh);
      for (int k = 0; k < nzh; ++k)
        for (int j = 0; j < nyh; ++j)
          for (int i = 0; i < nx; ++i) {
            CW(cw2b, i, 2 * j, k) = c2[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)];
            CW(cw2b, i, 2 * j + 1, k) = c2c[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)];
          }
    } else {
      inversion5_v2(a3, cw2b.data(), nx, ny, nzh);
    }
  }
  // :902-908
  for (int k = 0; k < nzh; ++k)
    for (int i = 0; i < nx; ++i)
      if (i + 1 == nx / 2 + 1 && k + 1 == nz / 2 + 1)
        for (int j = 0; j < ny; ++j) CW(cw2b, i, j, k) = 0.0;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int i = 0; i < nx; ++i) {
      CW(cw1, i, 0, k) = CW(cw2b, i, 0, k);
      for (int j = 1; j < ny; ++j) CW(cw1, i, j, k) = dct_pre(CW(cw2b, i, j, k), CW(cw2b, i, ny - j, k), ay[j], by[j]);
    }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        cplx c = CW(cw1, i, j, k);
        c = rot_bwd(c, ax[i], bx[i]);  // :966-968
        if (i + 1 > nx / 2 + 1) c = -c;
        c = rot_bwd(c, az[k], bz[k]);
        CW(cw1, i, j, k) = c;
      }
  fft3d_c2r(cw1.data(), rw.data(), nx, ny, nz);
  reorder_bwd(rw.data(), rhs, nx, ny, nz, 1);
}

// poisson.f90:1019-1465
void Poisson::solve_11x(double *rhs) {
  const size_t nr = static_cast<size_t>(nx) * ny * nz, nsp = static_cast<size_t>(nx) * ny * nzh;
  std::vector<double> ra(nr), rb(nr);
  std::vector<cplx> cw1(nsp), cw1b(nsp);
  const double *cur = rhs;
  if (bcz == 1) { reorder_fwd(cur, ra.data(), nx, ny, nz, 2); cur = ra.data(); }  // :1047-1057
  reorder_fwd(cur, rb.data(), nx, ny, nz, 1);                                     // :1064-1073
  reorder_fwd(rb.data(), ra.data(), nx, ny, nz, 0);                               // :1083-1092
  fft3d_r2c(ra.data(), cw1.data(), nx, ny, nz);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        cplx c = CW(cw1, i, j, k);
        c = c / static_cast<double>(nx) / static_cast<double>(ny) / static_cast<double>(nz);
        CW(cw1, i, j, k) = rot_fwd(c, az[k], bz[k]);  // :1119-1128
      }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int i = 0; i < nx; ++i) {  // :1138-1158
      CW(cw1b, i, 0, k) = CW(cw1, i, 0, k);
      for (int j = 1; j < ny; ++j) CW(cw1b, i, j, k) = dct_post<true>(CW(cw1, i, j, k), CW(cw1, i, ny - j, k), ay[j], by[j]);
    }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j) {  // :1174-1194
      CW(cw1, 0, j, k) = CW(cw1b, 0, j, k);
      for (int i = 1; i < nx; ++i) CW(cw1, i, j, k) = dct_post<true>(CW(cw1b, i, j, k), CW(cw1b, nx - i, j, k), ax[i], bx[i]);
    }
  // cw1 now holds the reference's cw1b
  if (istret == 0) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nzh; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) CW(cw1, i, j, k) = divide4(CW(cw1, i, j, k), CW(kxyz, i, j, k));  // :1204-1224
  } else {
    matrice_refinement();  // :1232
    const int nyh = ny / 2;
    if (istret != 3) {
      std::vector<cplx> c2(static_cast<size_t>(nx) * nyh * nzh), c2c(static_cast<size_t>(nx) * nyh * nzh);
      for (int k = 0; k < nzh; ++k)
        for (int j = 0; j < nyh; ++j)
          for (int i = 0; i < nx; ++i) {
            c2[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)] = CW(cw1, i, 2 * j, k);
            c2c[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)] = CW(cw1, i, 2 * j + 1, k);
          }
      inversion5_v1(a, c2.data(), nx, nyh, nzh);
      inversion5_v1(a2, c2c.data(), nx, nyh, nzh);
      for (int k = 0; k < nzh; ++k)
        for (int j = 0; j < nyh; ++j)
          for (int i = 0; i < nx; ++i) {
            CW(cw1, i, 2 * j, k) = c2[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)];
            CW(cw1, i, 2 * j + 1, k) = c2c[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(nyh) * k)];
          }
    } else {
      inversion5_v2(a3, cw1.data(), nx, ny, nzh);
    }
  }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j) {  // :1338-1358
      CW(cw1b, 0, j, k) = CW(cw1, 0, j, k);
      for (int i = 1; i < nx; ++i) CW(cw1b, i, j, k) = dct_pre(CW(cw1, i, j, k), CW(cw1, nx - i, j, k), ax[i], bx[i]);
    }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int i = 0; i < nx; ++i) {  // :1368-1388
      CW(cw1, i, 0, k) = CW(cw1b, i, 0, k);
      for (int j = 1; j < ny; ++j) CW(cw1, i, j, k) = dct_pre(CW(cw1b, i, j, k), CW(cw1b, i, ny - j, k), ay[j], by[j]);
    }
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nzh; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) CW(cw1, i, j, k) = rot_bwd(CW(cw1, i, j, k), az[k], bz[k]);  // :1398-1407
  fft3d_c2r(cw1.data(), ra.data(), nx, ny, nz);
  double *cur2 = ra.data();
  if (bcz == 1) { reorder_bwd(cur2, rb.data(), nx, ny, nz, 2); cur2 = rb.data(); }  // :1422-1432
  double *other = (cur2 == ra.data()) ? rb.data() : ra.data();
  reorder_bwd(cur2, other, nx, ny, nz, 1);  // :1438-1447
  reorder_bwd(other, rhs, nx, ny, nz, 0);   // :1449-1458
}

}  // namespace x3do
