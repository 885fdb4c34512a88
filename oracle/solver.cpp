// solver.cpp -- oracle restatement of the time step on one rank (TEST INFRASTRUCTURE):
// src/xcompact3d.f90:29-102 (loop), src/transeq.f90:73-591 (momentum_rhs_eq),
// src/time_integrators.f90:18-187 (intt), src/navier.f90:206-499,502-789 (cor_vel, divergence,
// gradp, pre_correc), src/Case-TGV.f90:25-107,189-380 (init_tgv, postprocess_tgv),
// src/variables.f90:1388-1399 (RK3 coefficients), src/parameters.f90:273-304.
#include <cmath>
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {

double *Solver::W(int i, size_t n) {
  if (static_cast<int>(work.size()) <= i) work.resize(i + 1);
  if (work[i].size() < n) work[i].assign(n, 0.0);
  return work[i].data();
}

namespace {
OpDesc mk(OpKind kind, const AxisScheme &A, int npaire, const vec &f, const vec &s, const vec &w, const double *post = nullptr) {
  OpDesc op{};
  op.kind = kind; op.ncl1 = A.ncl1; op.ncln = A.ncln; op.periodic = A.periodic; op.npaire = npaire;
  op.n = A.n; op.nm = A.nm; op.f = f.data(); op.s = s.data(); op.w = w.data();
  op.c = &A.c; op.fc = &A.fc; op.post = post; op.rhs_only = false;
  return op;
}
// first / second derivative with the coefficient set the call sites pair with npaire (transeq.f90:125-130,442-444)
void der1(const AxisScheme &A, int axis, const int d[3], const double *u, double *t, int npaire, const double *post) {
  OpDesc op = npaire == 1 ? mk(D1, A, 1, A.ffp, A.fsp, A.fwp, post) : mk(D1, A, 0, A.ff, A.fs, A.fw, post);
  apply_op(op, axis, d, u, t);
}
void der2(const AxisScheme &A, int axis, const int d[3], const double *u, double *t, int npaire) {
  OpDesc op = npaire == 1 ? mk(D2, A, 1, A.sfp, A.ssp, A.swp) : mk(D2, A, 0, A.sf, A.ss, A.sw);
  apply_op(op, axis, d, u, t);
}
}  // namespace

void Solver::init() {
  X = make_axis(p.nx, p.ncl[0][0], p.ncl[0][1], p.xlx, p.opt);
  Y = make_axis(p.ny, p.ncl[1][0], p.ncl[1][1], p.yly, p.opt);
  Z = make_axis(p.nz, p.ncl[2][0], p.ncl[2][1], p.zlz, p.opt);
  nxm = X.nm; nym = Y.nm; nzm = Z.nm;
  xnu = 1.0 / p.re;  // parameters.f90:302
  if (p.cpg) {       // :305-311: re is Re_tau; viscosity from the estimated centre-line Reynolds number
    const double re_cent = std::pow(p.re / 0.116, 1.0 / 0.88);
    xnu = 1.0 / re_cent;
    fcpg = 2.0 / p.yly * ((p.re / re_cent) * (p.re / re_cent));
  }
  st.istret = p.istret;
  if (p.istret != 0) {
    st.beta = p.beta; st.yly = p.yly;
    stretching(st, p.ny, nym, p.ncl[1][0], p.ncl[1][1], Y.periodic);
  }
  const double dt = p.dt;
  if (p.itimescheme == 5) {  // variables.f90:1388-1399
    iadvance_time = 3;
    adt[0] = (8.0 / 15.0) * dt; bdt[0] = 0.0; gdt[0] = adt[0];
    adt[1] = (5.0 / 12.0) * dt; bdt[1] = (-17.0 / 60.0) * dt; gdt[1] = adt[1] + bdt[1];
    adt[2] = (3.0 / 4.0) * dt; bdt[2] = (-5.0 / 12.0) * dt; gdt[2] = adt[2] + bdt[2];
    ntime = 2;
  } else if (p.itimescheme == 1) {  // Euler, variables.f90:1343-1349
    iadvance_time = 1; adt[0] = dt; bdt[0] = 0.0; gdt[0] = adt[0] + bdt[0]; ntime = 1;
  } else if (p.itimescheme == 2) {  // AB2, :1355-1363
    iadvance_time = 1; adt[0] = 1.5 * dt; bdt[0] = -0.5 * dt; gdt[0] = adt[0] + bdt[0]; ntime = 2;
  } else if (p.itimescheme == 3) {  // AB3, :1364-1374
    iadvance_time = 1;
    adt[0] = (23.0 / 12.0) * dt; bdt[0] = -(16.0 / 12.0) * dt; cdt[0] = (5.0 / 12.0) * dt;
    gdt[0] = adt[0] + bdt[0] + cdt[0]; ntime = 3;
  } else {
    throw std::runtime_error("oracle solver: itimescheme 1 (Euler), 2 (AB2), 3 (AB3) and 5 (RK3) are restated");
  }
  po.init(X, Y, Z, p.istret ? &st : nullptr);
  const size_t n = static_cast<size_t>(p.nx) * p.ny * p.nz;
  ux.assign(n, 0.0); uy.assign(n, 0.0); uz.assign(n, 0.0);
  px.assign(n, 0.0); py.assign(n, 0.0); pz.assign(n, 0.0);
  pp3.assign(static_cast<size_t>(nxm) * nym * nzm, 0.0);
  for (int q = 0; q < ntime; ++q) { dux[q].assign(n, 0.0); duy[q].assign(n, 0.0); duz[q].assign(n, 0.0); }
  const size_t nyz = static_cast<size_t>(p.ny) * p.nz, nxz = static_cast<size_t>(p.nx) * p.nz, nxy = static_cast<size_t>(p.nx) * p.ny;
  for (auto *v : {&dpdyx1, &dpdzx1, &dpdyxn, &dpdzxn}) v->assign(nyz, 0.0);
  for (int q = 0; q < 6; ++q) bw[q].assign(nyz, 0.0);
  for (auto *v : {&bxo, &byo, &bzo}) v->assign(nyz, 0.0);
  for (auto *v : {&dpdxy1, &dpdzy1, &dpdxyn, &dpdzyn}) v->assign(nxz, 0.0);
  for (auto *v : {&dpdxz1, &dpdyz1, &dpdxzn, &dpdyzn}) v->assign(nxy, 0.0);
  for (int q = 6; q < 12; ++q) bw[q].assign(nxz, 0.0);
  for (int q = 12; q < 18; ++q) bw[q].assign(nxy, 0.0);
  itime = 0;
}

// Case-Channel.f90:71-94 (iin = 0): ux = 1 - y^2, uz = sin(x) + cos(z), walls at y = -+ yly/2
void Solver::init_channel() {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const double dx = X.d, dy = Y.d, dz = Z.d;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j) {
      const double y = (p.istret == 0) ? static_cast<double>(j) * dy - p.yly * 0.5 : st.yp[j] - p.yly * 0.5;
      for (int i = 0; i < nx; ++i) {
        const size_t id = i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k);
        ux[id] = 1.0 - y * y;
        uy[id] = 0.0;
        uz[id] = std::sin(static_cast<double>(i) * dx) + std::cos(static_cast<double>(k) * dz);
      }
    }
}

// Case-Channel.f90:220-261; ppy = nullptr on a uniform mesh (ppy = 1)
void channel_cfr_apply(double *u, int nx, int ny, int nz, const double *ppy, double dy, double yly, double constant) {
  double ub = 0.0;
  const double coeff = dy / (yly * static_cast<double>(nx * nz));
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j) {
      const double pj = ppy ? ppy[j] : 1.0;
      for (int i = 0; i < nx; ++i) ub = ub + u[i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k)] / pj;
    }
  ub = ub * coeff;
  const double can = -(constant - ub);
  const size_t n = static_cast<size_t>(nx) * ny * nz;
  for (size_t q = 0; q < n; ++q) u[q] = u[q] - can;
}
void Solver::channel_cfr(std::vector<double> &u, double constant) {
  channel_cfr_apply(u.data(), p.nx, p.ny, p.nz, p.istret ? st.ppy.data() : nullptr, Y.d, p.yly, constant);
}

// boundary_conditions_channel, Case-Channel.f90:150-170 (cpg = F, idir_stream = 1)
void Solver::boundary_conditions() {
  if (p.itype == 5) { inflow(); outflow(); }   // boundary_conditions_cyl, Case-Cylinder-wake.f90:84-98
  if (p.itype == 3 && !p.cpg) channel_cfr(ux, 2.0 / 3.0);
}

// Case-TGV.f90:58-100 (iin=1, no noise)
void Solver::init_tgv() {
  const double dx = X.d, dy = Y.d, dz = Z.d;
  const int nx = p.nx, ny = p.ny, nz = p.nz;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j) {
      const double z = static_cast<double>(k) * dz, y = static_cast<double>(j) * dy;
      for (int i = 0; i < nx; ++i) {
        const double x = static_cast<double>(i) * dx;
        const size_t id = i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k);
        ux[id] = +std::sin(x) * std::cos(y) * std::cos(z);
        uy[id] = -std::cos(x) * std::sin(y) * std::cos(z);
        uz[id] = 0.0;
      }
    }
}

// transeq.f90:73-591, incompressible, iimplicit=0, no LES / scalar / forcing
void Solver::momentum_rhs_eq(double *dux1, double *duy1, double *duz1) {
  const int d[3] = {p.nx, p.ny, p.nz};
  const size_t n = static_cast<size_t>(p.nx) * p.ny * p.nz;
  double *ta = W(0, n), *tb = W(1, n), *tc = W(2, n), *td = W(3, n), *te = W(4, n), *tf = W(5, n);
  double *tg1 = W(6, n), *th1 = W(7, n), *ti1 = W(8, n);
  double *tg2 = W(9, n), *th2 = W(10, n), *ti2 = W(11, n);
  double *tg3 = W(12, n), *th3 = W(13, n), *ti3 = W(14, n), *tj = W(15, n);
  double *u = ux.data(), *v = uy.data(), *w = uz.data();
  const double *ppy = p.istret ? st.ppy.data() : nullptr;
  const double half = 0.5;
  const double bx = ubc[0], by = ubc[1], bz = ubc[2];
  // with iibm = 2 / 3 every collocated derivative first rebuilds its input inside the bodies, in place -- the
  // velocity arrays included (src/derive.f90:23-24); lind = product of the wall velocities (src/transeq.f90:120-146)
  auto der1 = [&](const AxisScheme &A, int axis, const int dd[3], double *in, double *out, int npaire, const double *post, double lind = 0.0) {
    ibm_prepass(axis, in, lind);
    x3do::der1(A, axis, dd, in, out, npaire, post);
  };
  auto der2 = [&](const AxisScheme &A, int axis, const int dd[3], double *in, double *out, int npaire, double lind = 0.0) {
    ibm_prepass(axis, in, lind);
    x3do::der2(A, axis, dd, in, out, npaire);
  };
  // ---- x pencils, :114-146
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { ta[q] = u[q] * u[q]; tb[q] = u[q] * v[q]; tc[q] = u[q] * w[q]; }
  der1(X, 0, d, ta, td, 1, nullptr, bx * bx); der1(X, 0, d, tb, te, 0, nullptr, bx * by); der1(X, 0, d, tc, tf, 0, nullptr, bx * bz);
  der1(X, 0, d, u, ta, 0, nullptr, bx); der1(X, 0, d, v, tb, 1, nullptr, by); der1(X, 0, d, w, tc, 1, nullptr, bz);
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { tg1[q] = td[q] + u[q] * ta[q]; th1[q] = te[q] + u[q] * tb[q]; ti1[q] = tf[q] + u[q] * tc[q]; }
  // ---- y pencils, :188-219
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { td[q] = u[q] * v[q]; te[q] = v[q] * v[q]; tf[q] = w[q] * v[q]; }
  der1(Y, 1, d, td, tg2, 0, ppy, bx * by); der1(Y, 1, d, te, th2, 1, ppy, by * by); der1(Y, 1, d, tf, ti2, 0, ppy, bz * by);
  der1(Y, 1, d, u, td, 1, ppy, bx); der1(Y, 1, d, v, te, 0, ppy, by); der1(Y, 1, d, w, tf, 1, ppy, bz);
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { tg2[q] = tg2[q] + v[q] * td[q]; th2[q] = th2[q] + v[q] * te[q]; ti2[q] = ti2[q] + v[q] * tf[q]; }
  // ---- z pencils, :249-314
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { td[q] = u[q] * w[q]; te[q] = v[q] * w[q]; tf[q] = w[q] * w[q]; }
  der1(Z, 2, d, td, tg3, 0, nullptr, bx * bz); der1(Z, 2, d, te, th3, 0, nullptr, by * bz); der1(Z, 2, d, tf, ti3, 1, nullptr, bz * bz);
  der1(Z, 2, d, u, td, 1, nullptr, bx); der1(Z, 2, d, v, te, 1, nullptr, by); der1(Z, 2, d, w, tf, 0, nullptr, bz);
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { td[q] = tg3[q] + w[q] * td[q]; te[q] = th3[q] + w[q] * te[q]; tf[q] = ti3[q] + w[q] * tf[q]; }
  der2(Z, 2, d, u, ta, 1, bx); der2(Z, 2, d, v, tb, 1, by); der2(Z, 2, d, w, tc, 0, bz);
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) {
    td[q] = xnu * ta[q] - half * td[q]; te[q] = xnu * tb[q] - half * te[q]; tf[q] = xnu * tc[q] - half * tf[q];
    // back in y pencils, :323-325
    tg2[q] = td[q] - half * tg2[q]; th2[q] = te[q] - half * th2[q]; ti2[q] = tf[q] - half * ti2[q];
  }
  // ---- diffusive terms in y, :336-372
  der2(Y, 1, d, u, td, 1, bx); der2(Y, 1, d, v, te, 0, by); der2(Y, 1, d, w, tf, 1, bz);
  if (p.istret != 0) {
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    der1(Y, 1, d, u, tj, 1, ppy, bx);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k);
      td[q] = td[q] * st.pp2y[j] - st.pp4y[j] * tj[q];
    }
    der1(Y, 1, d, v, tj, 0, ppy, by);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k);
      te[q] = te[q] * st.pp2y[j] - st.pp4y[j] * tj[q];
    }
    der1(Y, 1, d, w, tj, 1, ppy, bz);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k);
      tf[q] = tf[q] * st.pp2y[j] - st.pp4y[j] * tj[q];
    }
  }
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { ta[q] = xnu * td[q] + tg2[q]; tb[q] = xnu * te[q] + th2[q]; tc[q] = xnu * tf[q] + ti2[q]; }  // :431-433
  // ---- diffusive terms in x and final sum, :442-470
  der2(X, 0, d, u, td, 0, bx); der2(X, 0, d, v, te, 1, by); der2(X, 0, d, w, tf, 1, bz);
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) {
    td[q] = xnu * td[q]; te[q] = xnu * te[q]; tf[q] = xnu * tf[q];
    dux1[q] = ta[q] - half * tg1[q] + td[q];
    duy1[q] = tb[q] - half * th1[q] + te[q];
    duz1[q] = tc[q] - half * ti1[q] + tf[q];
  }
}

// time_integrators.f90:71-74 (Euler), :151-157 (RK3)
void Solver::intt(std::vector<double> &var, std::vector<double> *dvar) {
  const size_t n = var.size();
  double *v = var.data();
  const double *d1 = dvar[0].data();
  if (p.itimescheme == 1) {
    const double g = gdt[itr - 1];
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < n; ++q) v[q] = g * d1[q] + v[q];
    return;
  }
  if (p.itimescheme == 2) {  // time_integrators.f90:75-84
    double *d2 = dvar[1].data();
    const double g = gdt[itr - 1], a = adt[itr - 1], b = bdt[itr - 1];
    if (itime == 1) {
#pragma omp parallel for schedule(static)
      for (size_t q = 0; q < n; ++q) v[q] = g * d1[q] + v[q];
    } else {
#pragma omp parallel for schedule(static)
      for (size_t q = 0; q < n; ++q) v[q] = a * d1[q] + b * d2[q] + v[q];
    }
    for (size_t q = 0; q < n; ++q) d2[q] = d1[q];
    return;
  }
  if (p.itimescheme == 3) {  // :85-100
    double *d2 = dvar[1].data(), *d3 = dvar[2].data();
    const double dt = p.dt, a = adt[itr - 1], b = bdt[itr - 1], c = cdt[itr - 1];
    if (itime == 1) {
#pragma omp parallel for schedule(static)
      for (size_t q = 0; q < n; ++q) v[q] = dt * d1[q] + v[q];
    } else if (itime == 2) {
#pragma omp parallel for schedule(static)
      for (size_t q = 0; q < n; ++q) { v[q] = 1.5 * dt * d1[q] - 0.5 * dt * d2[q] + v[q]; d3[q] = d2[q]; }
    } else {
#pragma omp parallel for schedule(static)
      for (size_t q = 0; q < n; ++q) { v[q] = a * d1[q] + b * d2[q] + c * d3[q] + v[q]; d3[q] = d2[q]; }
    }
    for (size_t q = 0; q < n; ++q) d2[q] = d1[q];
    return;
  }
  double *d2 = dvar[1].data();
  if (itr == 1) {
    const double g = gdt[0];
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < n; ++q) { v[q] = g * d1[q] + v[q]; d2[q] = d1[q]; }
  } else {
    const double a = adt[itr - 1], b = bdt[itr - 1];
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < n; ++q) { v[q] = a * d1[q] + b * d2[q] + v[q]; d2[q] = d1[q]; }
  }
}

// navier.f90:502-789.  Dirichlet faces (ncl = 2) are no-slip walls here: the case arrays b?? are zero
// (Case-Channel.f90:67); the tangential components get the wall pressure gradient stored by gradp.
void Solver::pre_correc() {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  auto id = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k); };
  const double g = gdt[itr - 1];
  std::vector<double> &bxx1 = bw[0], &bxy1 = bw[1], &bxz1 = bw[2], &bxxn = bw[3], &bxyn = bw[4], &bxzn = bw[5];
  std::vector<double> &byx1 = bw[6], &byy1 = bw[7], &byz1 = bw[8], &byxn = bw[9], &byyn = bw[10], &byzn = bw[11];
  std::vector<double> &bzx1 = bw[12], &bzy1 = bw[13], &bzz1 = bw[14], &bzxn = bw[15], &bzyn = bw[16], &bzzn = bw[17];
  // inflow / outflow flow-rate balance, :534-560 (itype channel, uniform, abl with nclx = 2 on both sides)
  if ((p.itype == 3 || p.itype == 11 || p.itype == 10) && p.ncl[0][0] == 2 && p.ncl[0][1] == 2) {
    double ut1 = 0.0, ut = 0.0;
    for (size_t q = 0; q < bxx1.size(); ++q) ut1 = ut1 + bxx1[q];
    ut1 = ut1 / static_cast<double>(ny * nz);
    for (size_t q = 0; q < bxxn.size(); ++q) ut = ut + bxxn[q];
    ut = ut / static_cast<double>(ny * nz);
    for (size_t q = 0; q < bxxn.size(); ++q) bxxn[q] = bxxn[q] - ut + ut1;
  }
  if (p.ncl[0][0] == 2)  // :564-579
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) {
      const size_t q = j + static_cast<size_t>(ny) * k;
      dpdyx1[q] *= g; dpdzx1[q] *= g;
      ux[id(0, j, k)] = bxx1[q]; uy[id(0, j, k)] = bxy1[q] + dpdyx1[q]; uz[id(0, j, k)] = bxz1[q] + dpdzx1[q];
    }
  if (p.ncl[0][1] == 2)  // :580-595
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) {
      const size_t q = j + static_cast<size_t>(ny) * k;
      dpdyxn[q] *= g; dpdzxn[q] *= g;
      ux[id(nx - 1, j, k)] = bxxn[q]; uy[id(nx - 1, j, k)] = bxyn[q] + dpdyxn[q]; uz[id(nx - 1, j, k)] = bxzn[q] + dpdzxn[q];
    }
  if (p.ncl[0][0] == 1) for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) ux[id(0, j, k)] = 0.0;        // :600-606
  if (p.ncl[0][1] == 1) for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) ux[id(nx - 1, j, k)] = 0.0;   // :607-613
  if (p.ncl[1][0] == 2)  // :616-642
    for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * k;
      dpdxy1[q] *= g; dpdzy1[q] *= g;
      ux[id(i, 0, k)] = byx1[q] + dpdxy1[q]; uy[id(i, 0, k)] = byy1[q]; uz[id(i, 0, k)] = byz1[q] + dpdzy1[q];
    }
  if (p.ncl[1][1] == 2)  // :644-689
    for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * k;
      dpdxyn[q] *= g; dpdzyn[q] *= g;
      ux[id(i, ny - 1, k)] = byxn[q] + dpdxyn[q]; uy[id(i, ny - 1, k)] = byyn[q]; uz[id(i, ny - 1, k)] = byzn[q] + dpdzyn[q];
    }
  if (p.ncl[1][0] == 1) for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) uy[id(i, 0, k)] = 0.0;        // :693-701
  if (p.ncl[1][1] == 1) for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) uy[id(i, ny - 1, k)] = 0.0;   // :703-711
  if (p.ncl[2][0] == 2)  // :712-728
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * j;
      dpdxz1[q] *= g; dpdyz1[q] *= g;
      ux[id(i, j, 0)] = bzx1[q] + dpdxz1[q]; uy[id(i, j, 0)] = bzy1[q] + dpdyz1[q]; uz[id(i, j, 0)] = bzz1[q];
    }
  if (p.ncl[2][1] == 2)  // :729-745
    for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
      const size_t q = i + static_cast<size_t>(nx) * j;
      dpdxzn[q] *= g; dpdyzn[q] *= g;
      ux[id(i, j, nz - 1)] = bzxn[q] + dpdxzn[q]; uy[id(i, j, nz - 1)] = bzyn[q] + dpdyzn[q]; uz[id(i, j, nz - 1)] = bzzn[q];
    }
  if (p.ncl[2][0] == 1) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) uz[id(i, j, 0)] = 0.0;        // :751-759
  if (p.ncl[2][1] == 1) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) uz[id(i, j, nz - 1)] = 0.0;   // :761-769
  if (iibm == 1) {  // :782-786, solid body old school
    const size_t n = ux.size();
    ibm_corgp(ux.data(), uy.data(), uz.data(), px.data(), py.data(), pz.data(), n, 1);
    ibm_body(ux.data(), uy.data(), uz.data(), ep1.data(), n);
    ibm_corgp(ux.data(), uy.data(), uz.data(), px.data(), py.data(), pz.data(), n, 2);
  }
}

void Solver::ibm_prepass(int axis, double *arr, double lind) {
  if (iibm != 2 && iibm != 3) return;
  if (!ibm[axis].set) throw std::runtime_error("oracle solver: iibm = 2 / 3 without geometry for this direction");
  const AxisScheme &A = axis == 0 ? X : (axis == 1 ? Y : Z);
  const double len = axis == 0 ? p.xlx : (axis == 1 ? p.yly : p.zlz);
  std::vector<double> yp_uniform;
  const double *coords = nullptr;
  if (axis == 1) {
    if (p.istret != 0) coords = st.yp.data();
    else { yp_uniform.resize(p.ny); for (int j = 0; j < p.ny; ++j) yp_uniform[j] = j * A.d; coords = yp_uniform.data(); }
  }
  if (iibm == 2) lagpol(arr, p.nx, p.ny, p.nz, axis, ibm[axis].g, coords, A.d, len);
  else cubspl(arr, p.nx, p.ny, p.nz, axis, ibm[axis].g, coords, A.d, len, lind, nullptr, nullptr);
}

// Case-Cylinder-wake.f90:205-279 with iin = 0 (no initial noise): uniform stream u1
void Solver::init_cyl() {
  for (size_t q = 0; q < ux.size(); ++q) { ux[q] = 0.0 + p.u1; uy[q] = 0.0; uz[q] = 0.0; }
  itime = 0;
}

// momentum_forcing_channel, Case-Channel.f90:396-420 (idir_stream = 1)
void Solver::momentum_forcing(double *dux1, double *duy1, double *) {
  if (p.itype != 3) return;
  const size_t n = ux.size();
  if (p.cpg)
    for (size_t q = 0; q < n; ++q) dux1[q] = dux1[q] + fcpg;
  if (itime < p.spinup_time && p.iin <= 2) {
    const double w = p.wrotation;
    for (size_t q = 0; q < n; ++q) { dux1[q] = dux1[q] - w * uy[q]; duy1[q] = duy1[q] + w * ux[q]; }
  }
}

// Case-Cylinder-wake.f90:100-133: inflow plane; bxo, byo, bzo are the reference's random_number planes
void Solver::inflow() {
  for (size_t q = 0; q < bw[0].size(); ++q) {
    bw[0][q] = p.u1 + bxo[q] * p.inflow_noise;
    bw[1][q] = 0.0 + byo[q] * p.inflow_noise;
    bw[2][q] = 0.0 + bzo[q] * p.inflow_noise;
  }
}

// Case-Cylinder-wake.f90:135-203: convective outflow, celerity from the plane nx - 1 or from u1 / u2
void Solver::outflow() {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  auto id = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k); };
  const double dx = X.d, udx = 1.0 / dx;
  double uxmax = -1609.0, uxmin = 1609.0;
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) {
    const double v = ux[id(nx - 2, j, k)];
    if (v > uxmax) uxmax = v;
    if (v < uxmin) uxmin = v;
  }
  const double g = gdt[itr - 1];
  double cx;
  if (p.u1 == 0.0) cx = (0.5 * (uxmax + uxmin)) * g * udx;
  else if (p.u1 == 1.0) cx = uxmax * g * udx;
  else if (p.u1 == 2.0) cx = p.u2 * g * udx;
  else cx = (0.5 * (p.u1 + p.u2)) * g * udx;
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) {
    const size_t q = j + static_cast<size_t>(ny) * k;
    bw[3][q] = ux[id(nx - 1, j, k)] - cx * (ux[id(nx - 1, j, k)] - ux[id(nx - 2, j, k)]);
    bw[4][q] = uy[id(nx - 1, j, k)] - cx * (uy[id(nx - 1, j, k)] - uy[id(nx - 2, j, k)]);
    bw[5][q] = uz[id(nx - 1, j, k)] - cx * (uz[id(nx - 1, j, k)] - uz[id(nx - 2, j, k)]);
  }
}

// src/ibm.f90:52-80 and :14-49
void ibm_body(double *ux, double *uy, double *uz, const double *ep, size_t n) {
  for (size_t q = 0; q < n; ++q) { ux[q] = (1.0 - ep[q]) * ux[q]; uy[q] = (1.0 - ep[q]) * uy[q]; uz[q] = (1.0 - ep[q]) * uz[q]; }
}
void ibm_corgp(double *ux, double *uy, double *uz, const double *px, const double *py, const double *pz, size_t n, int nlock) {
  if (nlock == 1) for (size_t q = 0; q < n; ++q) { ux[q] = -px[q] + ux[q]; uy[q] = -py[q] + uy[q]; uz[q] = -pz[q] + uz[q]; }
  if (nlock == 2) for (size_t q = 0; q < n; ++q) { ux[q] = px[q] + ux[q]; uy[q] = py[q] + uy[q]; uz[q] = pz[q] + uz[q]; }
}

// navier.f90:257-372
void Solver::divergence(double *pp3out, int nlock, double *tmax_out, double *tmoy_out) {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const size_t n1 = static_cast<size_t>(nxm) * ny * nz, n2 = static_cast<size_t>(nxm) * nym * nz;
  const size_t n3 = static_cast<size_t>(nxm) * nym * nzm;
  double *pp1 = W(20, n1), *pgy1 = W(21, n1), *pgz1 = W(22, n1);
  double *upi2 = W(23, n2), *duydypi2 = W(24, n2), *po3 = W(25, n3);
  const int dx1[3] = {nx, ny, nz};
  const double *ta1 = ux.data(), *tb1 = uy.data(), *tc1 = uz.data();
  if (iibm != 0) {  // :285-293
    const size_t n = static_cast<size_t>(nx) * ny * nz;
    double *a = W(26, n), *b = W(27, n), *c = W(28, n);
    for (size_t q = 0; q < n; ++q) {
      a[q] = (1.0 - ep1[q]) * ux[q] + ep1[q] * ubc[0];
      b[q] = (1.0 - ep1[q]) * uy[q] + ep1[q] * ubc[1];
      c[q] = (1.0 - ep1[q]) * uz[q] + ep1[q] * ubc[2];
    }
    ta1 = a; tb1 = b; tc1 = c;
  }
  apply_op(mk(DVP, X, 0, X.cfx6, X.csx6, X.cwx6), 0, dx1, ta1, pp1);        // :297
  apply_op(mk(IVP, X, 1, X.cifxp6, X.cisxp6, X.ciwxp6), 0, dx1, tb1, pgy1);  // :313
  apply_op(mk(IVP, X, 1, X.cifxp6, X.cisxp6, X.ciwxp6), 0, dx1, tc1, pgz1);  // :314
  const int dy2[3] = {nxm, ny, nz};
  const double *ppyi = p.istret ? st.ppyi.data() : nullptr;
  apply_op(mk(IVP, Y, 1, Y.cifxp6, Y.cisxp6, Y.ciwxp6), 1, dy2, pp1, upi2);           // :321
  apply_op(mk(DVP, Y, 0, Y.cfx6, Y.csx6, Y.cwx6, ppyi), 1, dy2, pgy1, duydypi2);       // :322
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n2; ++q) duydypi2[q] = duydypi2[q] + upi2[q];                 // :325
  apply_op(mk(IVP, Y, 1, Y.cifxp6, Y.cisxp6, Y.ciwxp6), 1, dy2, pgz1, upi2);           // :327
  const int dz3[3] = {nxm, nym, nz};
  apply_op(mk(IVP, Z, 1, Z.cifxp6, Z.cisxp6, Z.ciwxp6), 2, dz3, duydypi2, pp3out);     // :333
  apply_op(mk(DVP, Z, 0, Z.cfx6, Z.csx6, Z.cwx6), 2, dz3, upi2, po3);                  // :335
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n3; ++q) pp3out[q] = pp3out[q] + po3[q];                      // :339
  if (nlock == 2) {                                                                     // :341-347
    const double pres_ref = pp3out[static_cast<size_t>(nxm) * nym * (nzm - 1)];
    for (size_t q = 0; q < n3; ++q) pp3out[q] = pp3out[q] - pres_ref;
  }
  double tmax = -1609.0, tmoy = 0.0;                                                    // :349-359
  for (size_t q = 0; q < n3; ++q) { if (pp3out[q] > tmax) tmax = pp3out[q]; tmoy += std::fabs(pp3out[q]); }
  tmoy = tmoy / static_cast<double>(n3);
  if (tmax_out) *tmax_out = tmax;
  if (tmoy_out) *tmoy_out = tmoy;
}

// navier.f90:386-431
void Solver::gradp(double *px1, double *py1, double *pz1, const double *pp3in) {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const size_t n3 = static_cast<size_t>(nxm) * nym * nz, n2 = static_cast<size_t>(nxm) * ny * nz;
  double *ppi3 = W(30, n3), *pgz3 = W(31, n3), *ppi2 = W(32, n2), *pgy2 = W(33, n2), *pgzi2 = W(34, n2);
  auto pv = [&](OpKind k, const AxisScheme &A, const double *post) {
    const bool inter = (k == IPV);
    if (A.periodic) return inter ? mk(k, A, 1, A.cifx6, A.cisx6, A.ciwx6, post) : mk(k, A, 1, A.cfx6, A.csx6, A.cwx6, post);
    return inter ? mk(k, A, 1, A.cifip6, A.cisip6, A.ciwip6, post) : mk(k, A, 1, A.cfip6, A.csip6, A.cwip6, post);
  };
  const int dz[3] = {nxm, nym, nzm};
  apply_op(pv(IPV, Z, nullptr), 2, dz, pp3in, ppi3);  // :404
  apply_op(pv(DPV, Z, nullptr), 2, dz, pp3in, pgz3);  // :406
  const int dy[3] = {nxm, nym, nz};
  const double *ppy = p.istret ? st.ppy.data() : nullptr;
  apply_op(pv(IPV, Y, nullptr), 1, dy, ppi3, ppi2);   // :413
  apply_op(pv(DPV, Y, ppy), 1, dy, ppi3, pgy2);       // :415
  apply_op(pv(IPV, Y, nullptr), 1, dy, pgz3, pgzi2);  // :417
  const int dxx[3] = {nxm, ny, nz};
  apply_op(pv(DPV, X, nullptr), 0, dxx, ppi2, px1);   // :426
  apply_op(pv(IPV, X, nullptr), 0, dxx, pgy2, py1);   // :428
  apply_op(pv(IPV, X, nullptr), 0, dxx, pgzi2, pz1);  // :430
  capture_wall_gradients(px1, py1, pz1);
}

// wall pressure gradients for the next pre_correc, navier.f90:439-496 (the z faces store py1 / pz1 as the reference does)
void Solver::capture_wall_gradients(const double *px1, const double *py1, const double *pz1) {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  auto id = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k); };
  const double g = gdt[itr - 1];
  if (p.ncl[0][0] == 2) for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) { dpdyx1[j + static_cast<size_t>(ny) * k] = py1[id(0, j, k)] / g; dpdzx1[j + static_cast<size_t>(ny) * k] = pz1[id(0, j, k)] / g; }
  if (p.ncl[0][1] == 2) for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) { dpdyxn[j + static_cast<size_t>(ny) * k] = py1[id(nx - 1, j, k)] / g; dpdzxn[j + static_cast<size_t>(ny) * k] = pz1[id(nx - 1, j, k)] / g; }
  if (p.ncl[1][0] == 2) for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) { dpdxy1[i + static_cast<size_t>(nx) * k] = px1[id(i, 0, k)] / g; dpdzy1[i + static_cast<size_t>(nx) * k] = pz1[id(i, 0, k)] / g; }
  if (p.ncl[1][1] == 2) for (int k = 0; k < nz; ++k) for (int i = 0; i < nx; ++i) { dpdxyn[i + static_cast<size_t>(nx) * k] = px1[id(i, ny - 1, k)] / g; dpdzyn[i + static_cast<size_t>(nx) * k] = pz1[id(i, ny - 1, k)] / g; }
  if (p.ncl[2][0] == 2) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { dpdxz1[i + static_cast<size_t>(nx) * j] = py1[id(i, j, 0)] / g; dpdyz1[i + static_cast<size_t>(nx) * j] = pz1[id(i, j, 0)] / g; }
  if (p.ncl[2][1] == 2) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) { dpdxzn[i + static_cast<size_t>(nx) * j] = py1[id(i, j, nz - 1)] / g; dpdyzn[i + static_cast<size_t>(nx) * j] = pz1[id(i, j, nz - 1)] / g; }
}

// navier.f90:242-244
void Solver::cor_vel() {
  const size_t n = ux.size();
#pragma omp parallel for schedule(static)
  for (size_t q = 0; q < n; ++q) { ux[q] = ux[q] - px[q]; uy[q] = uy[q] - py[q]; uz[q] = uz[q] - pz[q]; }
}

// xcompact3d.f90:29-102
void Solver::step() {
  itime += 1;
  for (itr = 1; itr <= iadvance_time; ++itr) {
    boundary_conditions();                         // xcompact3d.f90:52
    momentum_rhs_eq(dux[0].data(), duy[0].data(), duz[0].data());
    momentum_forcing(dux[0].data(), duy[0].data(), duz[0].data());   // the end of momentum_rhs_eq, transeq.f90:539
    intt(ux, dux); intt(uy, duy); intt(uz, duz);
    pre_correc();
    divergence(pp3.data(), 1, nullptr, nullptr);  // solve_poisson, navier.f90:76
    po.solve(pp3.data());                          // :99
    gradp(px.data(), py.data(), pz.data(), pp3.data());  // :107
    cor_vel();
  }
}

// Case-TGV.f90:189-380
void Solver::postprocess_tgv(double out[4]) {
  const int nx = p.nx, ny = p.ny, nz = p.nz;
  const int d[3] = {nx, ny, nz};
  const size_t n = static_cast<size_t>(nx) * ny * nz;
  const int xs1 = (p.ncl[0][0] == 1) ? nx - 1 : nx, xs2 = (p.ncl[1][0] == 1) ? ny - 1 : ny, xs3 = (p.ncl[2][0] == 1) ? nz - 1 : nz;
  const int nxc = (p.ncl[0][0] == 1) ? nxm : nx, nyc = (p.ncl[1][0] == 1) ? nym : ny, nzc = (p.ncl[2][0] == 1) ? nzm : nz;
  double *ta1 = W(40, n), *tb1 = W(41, n), *tc1 = W(42, n), *td1 = W(43, n), *te1 = W(44, n), *tf1 = W(45, n);
  double *tg1 = W(46, n), *th1 = W(47, n), *ti1 = W(48, n);
  const double *u = ux.data(), *v = uy.data(), *w = uz.data();
  const double *ppy = p.istret ? st.ppy.data() : nullptr;
  der1(X, 0, d, u, ta1, 0, nullptr); der1(X, 0, d, v, tb1, 1, nullptr); der1(X, 0, d, w, tc1, 1, nullptr);  // :261-263
  der1(Y, 1, d, u, td1, 1, ppy); der1(Y, 1, d, v, te1, 0, ppy); der1(Y, 1, d, w, tf1, 1, ppy);              // :265-267
  der1(Z, 2, d, u, tg1, 1, nullptr); der1(Z, 2, d, v, th1, 1, nullptr); der1(Z, 2, d, w, ti1, 0, nullptr);  // :269-271
  double enst = 0.0, eps = 0.0, eek = 0.0, eps2 = 0.0;
  auto ID = [&](int i, int j, int k) { return i + static_cast<size_t>(nx) * (j + static_cast<size_t>(ny) * k); };
  // the reference accumulates serially in (k,j,i) order (:287-296); keep that order
  for (int k = 0; k < xs3; ++k) for (int j = 0; j < xs2; ++j) for (int i = 0; i < xs1; ++i) {
    const size_t q = ID(i, j, k);
    const double a = tf1[q] - th1[q], b = tg1[q] - tc1[q], c = tb1[q] - td1[q];
    enst = enst + 0.5 * (a * a + b * b + c * c);
  }
  enst = enst / (static_cast<double>(nxc) * nyc * nzc);
  for (int k = 0; k < xs3; ++k) for (int j = 0; j < xs2; ++j) for (int i = 0; i < xs1; ++i) {
    const size_t q = ID(i, j, k);
    auto sq = [](double x) { return x * x; };
    eps = eps + 0.5 * xnu * (sq(2.0 * ta1[q]) + sq(2.0 * te1[q]) + sq(2.0 * ti1[q]) + 2.0 * sq(td1[q] + tb1[q]) +
                             2.0 * sq(tg1[q] + tc1[q]) + 2.0 * sq(th1[q] + tf1[q]));
  }
  eps = eps / (static_cast<double>(nxc) * nyc * nzc);
  for (int k = 0; k < xs3; ++k) for (int j = 0; j < xs2; ++j) for (int i = 0; i < xs1; ++i) {
    const size_t q = ID(i, j, k);
    eek = eek + 0.5 * (u[q] * u[q] + v[q] * v[q] + w[q] * w[q]);
  }
  eek = eek / (static_cast<double>(nxc) * nyc * nzc);
  der2(X, 0, d, u, ta1, 0); der2(X, 0, d, v, tb1, 1); der2(X, 0, d, w, tc1, 1);  // :332-334
  der2(Y, 1, d, u, td1, 1); der2(Y, 1, d, v, te1, 0); der2(Y, 1, d, w, tf1, 1);  // :336-338
  der2(Z, 2, d, u, tg1, 1); der2(Z, 2, d, v, th1, 1); der2(Z, 2, d, w, ti1, 0);  // :340-342
  for (int k = 0; k < xs3; ++k) for (int j = 0; j < xs2; ++j) for (int i = 0; i < xs1; ++i) {
    const size_t q = ID(i, j, k);
    const double di = (-xnu) * (u[q] * (ta1[q] + td1[q] + tg1[q]) + v[q] * (tb1[q] + te1[q] + th1[q]) + w[q] * (tc1[q] + tf1[q] + ti1[q]));
    eps2 = eps2 + di;
  }
  eps2 = eps2 / (static_cast<double>(nxc) * nyc * nzc);
  out[0] = eek; out[1] = eps; out[2] = eps2; out[3] = enst;
}

}  // namespace x3do
