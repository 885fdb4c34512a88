// x3d_ibm.cu -- immersed-boundary pre-pass of the collocated operators and filters: reconstruction of the input inside
// the solid bodies, by Lagrange polynomials when iibm = 2 (lagpolx / lagpoly / lagpolz and polint, src/ibm.f90:83-389)
// and by clamped cubic splines when iibm = 3 (cubsplx / cubsply / cubsplz and cubic_spline, src/ibm.f90:399-968).  Per
// line the work is a handful of points (the body interior) and depends on the geometry the host's genepsi3d produced:
// one thread per line, lines without a body leave at once.
#include <string>
#include "x3d_ctx.cuh"

namespace x3d {

namespace {

// Neville's algorithm, src/ibm.f90:345-389
__device__ double polint(const double *xa, const double *ya, int n, double x) {
  double c[10], d[10];
  int ns = 1;
  double dif = fabs(x - xa[0]);
  for (int i = 1; i <= n; ++i) {
    const double dift = fabs(x - xa[i - 1]);
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

struct LagArgs {
  int nl, na, nb;              // line length, the two other extents (reference order)
  long long sl, sa, sb;        // strides of the line index and of (a, b) in u
  int nobjmax, npif, izap, searched;   // searched: locate the faces in coords (lagpoly), else by division (lagpolx/z)
  double d, len;
  const int *nobj, *nipif, *nfpif;
  const double *xi, *xf, *coords;
};

__global__ void k_lagpol(double *__restrict__ u, const LagArgs g) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(g.na) * g.nb) return;
  const int nobj = g.nobj[idx];
  if (nobj == 0) return;
  const int a = static_cast<int>(idx % g.na), b = static_cast<int>(idx / g.na);
  double *line = u + a * g.sa + b * g.sb;
  const double d = g.d;
  for (int i = 1; i <= nobj; ++i) {
    double xa[10], ya[10];
    int ia = 0;
    const long long gi = (i - 1) + static_cast<long long>(g.nobjmax) * idx;
    const long long gp = i + static_cast<long long>(g.nobjmax + 1) * idx;
    const double xi = g.xi[gi], xf = g.xf[gi];
    int ipoli, ipolf;
    int npf = g.npif;
    xa[ia] = xi; ya[ia] = 0.0; ++ia;
    if (xi > 0.0) {  // immersed: fluid points before the body, :117-131
      int ix;
      if (g.searched) { ix = 1; while (g.coords[ix - 1] < xi) ix = ix + 1; ix = ix - 1; }
      else ix = static_cast<int>(xi / d + 1.0);
      ipoli = ix + 1;
      if (g.nipif[gp] < g.npif) npf = g.nipif[gp];
      for (int ip = 1; ip <= npf; ++ip) {
        const int q = g.izap == 1 ? ix - ip : ix - ip + 1;   // 1-based point
        xa[ia] = g.searched ? g.coords[q - 1] : (g.izap == 1 ? static_cast<double>(ix - 1) * d - ip * d : static_cast<double>(ix - 1) * d - (ip - 1) * d);
        ya[ia] = line[(q - 1) * g.sl];
        ++ia;
      }
    } else {
      ipoli = 1;
    }
    npf = g.npif;
    xa[ia] = xf; ya[ia] = 0.0; ++ia;
    if (xf < g.len) {  // fluid points after the body, :137-151
      int ix;
      if (g.searched) { ix = 1; while (g.coords[ix - 1] < xf) ix = ix + 1; }
      else ix = static_cast<int>((xf + d) / d + 1.0);
      ipolf = ix - 1;
      if (g.nfpif[gp] < g.npif) npf = g.nfpif[gp];
      for (int ip = 1; ip <= npf; ++ip) {
        const int q = g.izap == 1 ? ix + ip : ix + ip - 1;
        xa[ia] = g.searched ? g.coords[q - 1] : (g.izap == 1 ? static_cast<double>(ix - 1) * d + ip * d : static_cast<double>(ix - 1) * d + (ip - 1) * d);
        ya[ia] = line[(q - 1) * g.sl];
        ++ia;
      }
    } else {
      ipolf = g.nl;
    }
    for (int ipol = ipoli; ipol <= ipolf; ++ipol) {
      const double xpol = g.searched ? g.coords[ipol - 1] : d * static_cast<double>(ipol - 1);
      line[(ipol - 1) * g.sl] = polint(xa, ya, ia, xpol);
    }
  }
}


// clamped cubic spline through the points of one body, src/ibm.f90:880-968 (y is left as it is when x lies in none of
// the intervals)
__device__ void cubic_spline(const double *xa, const double *ya, int n, double x, double &y) {
  double xaa[10], yaa[10];
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

struct SplArgs {
  LagArgs g;
  int axis;                    // 0, 1, 2: the three routines differ in their boundary cases
  double bcimp;                // lind: the value imposed on the walls
  const double *ana_i, *ana_f; // analytic wall positions (ianal /= 0), or null
};

// cubsplx / cubsply / cubsplz, src/ibm.f90:399-874
__global__ void k_cubspl(double *__restrict__ u, const SplArgs s) {
  const LagArgs &g = s.g;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(g.na) * g.nb) return;
  const int nobj = g.nobj[idx];
  if (nobj == 0) return;
  const int a = static_cast<int>(idx % g.na), b = static_cast<int>(idx / g.na);
  double *line = u + a * g.sa + b * g.sb;
  const double d = g.d, bcimp = s.bcimp;
  const int axis = s.axis, nl = g.nl;
  // the reference keeps ypol across lines (it is whatever the previous spline call produced when x matches no
  // interval, a case it leaves undefined); here it is per line, starting from the wall value
  double ypol = bcimp;
  for (int i = 1; i <= nobj; ++i) {
    double xa[10], ya[10];
    int ia = 0;
    const long long gi = (i - 1) + static_cast<long long>(g.nobjmax) * idx;
    const long long gp = i + static_cast<long long>(g.nobjmax + 1) * idx;
    const double xi = g.xi[gi], xf = g.xf[gi];
    const double ana_resi = s.ana_i ? s.ana_i[gi] : xi, ana_resf = s.ana_f ? s.ana_f[gi] : xf;
    int ipoli, ipolf, inxi = 0, inxf = 0;
    // ---- first wall
    int npf = g.npif;
    xa[ia] = ana_resi; ya[ia] = bcimp; ++ia;
    if (g.nipif[gp] < g.npif) npf = g.nipif[gp];
    if (xi > 0.0) {
      int ix;
      if (axis == 1) { ix = 1; while (g.coords[ix - 1] < xi) ix = ix + 1; ix = ix - 1; }
      else ix = static_cast<int>(xi / d + 1.0);
      ipoli = ix + 1;
      for (int ip = 1; ip <= npf; ++ip) {
        const int q = g.izap == 1 ? ix - ip : ix - ip + 1;
        xa[ia] = axis == 1 ? g.coords[q - 1] : (g.izap == 1 ? static_cast<double>(ix - 1) * d - ip * d : static_cast<double>(ix - 1) * d - (ip - 1) * d);
        ya[ia] = line[(q - 1) * g.sl];
        ++ia;
      }
    } else {  // the body starts on the domain boundary: ghost points that carry the wall value
      inxi = 1;
      int ix = 0;
      if (axis == 1) { ix = 1; while (g.coords[ix - 1] < xi) ix = ix + 1; ix = ix - 1; ipoli = ix + 1; }
      else { ix = static_cast<int>(xi / d); ipoli = axis == 0 ? ix + 1 : 1; }
      for (int ip = 1; ip <= npf; ++ip) {
        if (axis == 1) xa[ia] = g.izap == 1 ? g.coords[0] - (ip + 1) * d : g.coords[0] - (ip * d);
        else xa[ia] = g.izap == 1 ? static_cast<double>(ix - 1) * d - ip * d : static_cast<double>(ix - 1) * d - (ip - 1) * d;
        ya[ia] = bcimp;
        ++ia;
      }
    }
    // ---- second wall
    npf = g.npif;
    xa[ia] = ana_resf; ya[ia] = bcimp; ++ia;
    if (g.nfpif[gp] < g.npif) npf = g.nfpif[gp];
    if (xf < g.len) {
      int ix;
      if (axis == 1) { ix = 1; while (g.coords[ix - 1] < xf) ix = ix + 1; }
      else ix = static_cast<int>((xf + d) / d + 1.0);
      ipolf = ix - 1;
      for (int ip = 1; ip <= npf; ++ip) {
        const int q = g.izap == 1 ? ix + ip : ix + ip - 1;
        xa[ia] = axis == 1 ? g.coords[q - 1] : (g.izap == 1 ? static_cast<double>(ix - 1) * d + ip * d : static_cast<double>(ix - 1) * d + (ip - 1) * d);
        ya[ia] = line[(q - 1) * g.sl];
        ++ia;
      }
    } else {
      inxf = 1;
      int ix;
      if (axis == 1) { ix = 1; while (ix <= nl && g.coords[ix - 1] < xf) ix = ix + 1; ipolf = ix - 1; }
      else { ix = static_cast<int>((xf + d) / d + 1.0); ipolf = axis == 0 ? ix - 1 : nl; }
      for (int ip = 1; ip <= npf; ++ip) {
        if (axis == 1) xa[ia] = g.izap == 1 ? g.coords[nl - 1] + (ip + 1) * d : g.coords[nl - 1] + ip * d;
        else xa[ia] = g.izap == 1 ? static_cast<double>(ix - 1) * d + ip * d : static_cast<double>(ix - 1) * d + (ip - 1) * d;
        ya[ia] = bcimp;
        ++ia;
      }
    }
    if (xi == xf) continue;   // "situation not supported by the IBM" in the reference (it aborts); left untouched here
    if (ipolf > nl) ipolf = nl;
    for (int ipol = ipoli < 1 ? 1 : ipoli; ipol <= ipolf; ++ipol) {
      if (axis != 1 && inxf == 1 && inxi == 1) { line[(ipol - 1) * g.sl] = bcimp; continue; }
      const double xpol = axis == 1 ? g.coords[ipol - 1] : d * static_cast<double>(ipol - 1);
      if (axis != 2 && (xpol == ana_resi || xpol == ana_resf)) { line[(ipol - 1) * g.sl] = bcimp; continue; }
      cubic_spline(xa, ya, ia, xpol, ypol);
      line[(ipol - 1) * g.sl] = ypol;
    }
  }
}

}  // namespace

void set_ibm_geometry(Ctx &ctx, int axis, int nobjmax, int npif, int izap, int na, int nb, const int *nobj, const double *xi,
                      const double *xf, const int *nipif, const int *nfpif, const double *coords, int ncoords, double d, double len) {
  if (axis < 0 || axis > 2 || nobjmax < 1 || na < 1 || nb < 1 || !nobj || !xi || !xf || !nipif || !nfpif) throw Error("x3d_set_ibm_geometry: bad argument");
  if (npif < 1 || 2 * npif + 2 > 10) throw Error("x3d_set_ibm_geometry: npif must be 1..4 (xa, ya hold 10 points, src/ibm.f90:97)");
  if (axis == 1 && !coords && !ctx.st_yp.empty()) {   // the host's yp went in through x3d_set_stretching
    coords = ctx.st_yp.data();
    ncoords = static_cast<int>(ctx.st_yp.size());
  }
  if (axis == 1 && !coords) throw Error("x3d_set_ibm_geometry: lagpoly needs yp (pass it here or call x3d_set_stretching first)");
  X3D_CUDA(cudaSetDevice(ctx.device));
  Ctx::IbmAxis &G = ctx.ibm[axis];
  G.nobjmax = nobjmax; G.npif = npif; G.izap = izap; G.na = na; G.nb = nb; G.d = d; G.len = len; G.ncoords = coords ? ncoords : 0;
  const size_t nl = static_cast<size_t>(na) * nb;
  auto up = [&](DevBuf &b, const void *src, size_t bytes) {
    b.reserve(bytes);
    X3D_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx.stream));
  };
  up(G.nobj, nobj, nl * sizeof(int));
  up(G.xi, xi, nl * nobjmax * sizeof(double));
  up(G.xf, xf, nl * nobjmax * sizeof(double));
  up(G.nipif, nipif, nl * (nobjmax + 1) * sizeof(int));
  up(G.nfpif, nfpif, nl * (nobjmax + 1) * sizeof(int));
  if (coords) up(G.coords, coords, static_cast<size_t>(ncoords) * sizeof(double));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  G.set = true;
}

// geometry arguments of one direction for a pencil u(nx, ny, nz); `what` names the caller in error messages
static LagArgs ibm_args(Ctx &ctx, int axis, int nx, int ny, int nz, const char *what) {
  Ctx::IbmAxis &G = ctx.ibm[axis];
  if (!G.set) throw Error(std::string(what) + ": x3d_set_ibm_geometry has not been called for this direction");
  const int n[3] = {nx, ny, nz};
  const int a_ax = axis == 0 ? 1 : 0, b_ax = axis == 2 ? 1 : 2;
  if (G.na != n[a_ax] || G.nb != n[b_ax]) throw Error(std::string(what) + ": the geometry arrays do not match the pencil");
  if (axis == 1 && G.ncoords < ny) throw Error(std::string(what) + ": yp is shorter than the line");
  const long long st[3] = {1, nx, static_cast<long long>(nx) * ny};
  LagArgs g{};
  g.nl = n[axis]; g.na = G.na; g.nb = G.nb;
  g.sl = st[axis]; g.sa = st[a_ax]; g.sb = st[b_ax];
  g.nobjmax = G.nobjmax; g.npif = G.npif; g.izap = G.izap; g.searched = axis == 1 ? 1 : 0;
  g.d = G.d; g.len = G.len;
  g.nobj = static_cast<const int *>(G.nobj.p); g.nipif = static_cast<const int *>(G.nipif.p); g.nfpif = static_cast<const int *>(G.nfpif.p);
  g.xi = static_cast<const double *>(G.xi.p); g.xf = static_cast<const double *>(G.xf.p); g.coords = static_cast<const double *>(G.coords.p);
  return g;
}

void lagpol_device(Ctx &ctx, int axis, double *d_u, int nx, int ny, int nz) {
  const LagArgs g = ibm_args(ctx, axis, nx, ny, nz, "iibm = 2");
  const long long nl = static_cast<long long>(g.na) * g.nb;
  ProfScope ps(ctx, "ibm_lagpol(k_lagpol)");
  k_lagpol<<<static_cast<unsigned>((nl + 127) / 128), 128, 0, ctx.stream>>>(d_u, g);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

// analytic wall positions for the cubic-spline variant (ianal /= 0): what analitic_x / analitic_y return for xi / xf,
// same shape as xi / xf; null pointers go back to ianal = 0
void set_ibm_analytic(Ctx &ctx, int axis, const double *ana_i, const double *ana_f) {
  if (axis < 0 || axis > 2) throw Error("x3d_set_ibm_analytic: bad axis");
  Ctx::IbmAxis &G = ctx.ibm[axis];
  if (!ana_i || !ana_f) { G.analytic = false; return; }
  if (!G.set) throw Error("x3d_set_ibm_analytic: call x3d_set_ibm_geometry for this direction first");
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t bytes = static_cast<size_t>(G.na) * G.nb * G.nobjmax * sizeof(double);
  G.ana_i.reserve(bytes); G.ana_f.reserve(bytes);
  X3D_CUDA(cudaMemcpyAsync(G.ana_i.p, ana_i, bytes, cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(G.ana_f.p, ana_f, bytes, cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  G.analytic = true;
}

void cubspl_device(Ctx &ctx, int axis, double *d_u, int nx, int ny, int nz, double lind) {
  SplArgs s{};
  s.g = ibm_args(ctx, axis, nx, ny, nz, "iibm = 3");
  s.axis = axis; s.bcimp = lind;
  const Ctx::IbmAxis &G = ctx.ibm[axis];
  s.ana_i = G.analytic ? static_cast<const double *>(G.ana_i.p) : nullptr;
  s.ana_f = G.analytic ? static_cast<const double *>(G.ana_f.p) : nullptr;
  const long long nl = static_cast<long long>(s.g.na) * s.g.nb;
  ProfScope ps(ctx, "ibm_cubspl(k_cubspl)");
  k_cubspl<<<static_cast<unsigned>((nl + 127) / 128), 128, 0, ctx.stream>>>(d_u, s);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

// host-or-device entries (lagpolx(u), cubsplx(u, lind) etc.): spline = false / true
static void ibm_rebuild(Ctx &ctx, bool spline, int axis, double *u, int nx, int ny, int nz, double lind) {
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t bytes = static_cast<size_t>(nx) * ny * nz * sizeof(double);
  if (bytes == 0) return;
  const bool dev = is_device_ptr(u);
  double *d_u = u;
  if (!dev) {
    ctx.stage_in.reserve(bytes);
    X3D_CUDA(cudaMemcpyAsync(ctx.stage_in.p, u, bytes, cudaMemcpyHostToDevice, ctx.stream));
    d_u = static_cast<double *>(ctx.stage_in.p);
  }
  if (spline) cubspl_device(ctx, axis, d_u, nx, ny, nz, lind);
  else lagpol_device(ctx, axis, d_u, nx, ny, nz);
  if (!dev) {
    X3D_CUDA(cudaMemcpyAsync(u, d_u, bytes, cudaMemcpyDeviceToHost, ctx.stream));
    X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  }
}
void lagpol(Ctx &ctx, int axis, double *u, int nx, int ny, int nz) { ibm_rebuild(ctx, false, axis, u, nx, ny, nz, 0.0); }
void cubspl(Ctx &ctx, int axis, double *u, int nx, int ny, int nz, double lind) { ibm_rebuild(ctx, true, axis, u, nx, ny, nz, lind); }

}  // namespace x3d
