// x3d_ibm.cu -- immersed-boundary pre-pass of the collocated operators when iibm = 2: Lagrange reconstruction of the
// input inside the solid bodies (lagpolx / lagpoly / lagpolz and polint, src/ibm.f90:83-389).  Per line the work is a
// handful of points (the body interior) and depends on the geometry the host's genepsi3d produced: one thread per
// line, lines without a body leave at once.
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

}  // namespace

void set_ibm_geometry(Ctx &ctx, int axis, int nobjmax, int npif, int izap, int na, int nb, const int *nobj, const double *xi,
                      const double *xf, const int *nipif, const int *nfpif, const double *coords, int ncoords, double d, double len) {
  if (axis < 0 || axis > 2 || nobjmax < 1 || na < 1 || nb < 1 || !nobj || !xi || !xf || !nipif || !nfpif) throw Error("x3d_set_ibm_geometry: bad argument");
  if (npif < 1 || 2 * npif + 2 > 10) throw Error("x3d_set_ibm_geometry: npif must be 1..4 (xa, ya hold 10 points, src/ibm.f90:97)");
  if (axis == 1 && !coords) throw Error("x3d_set_ibm_geometry: lagpoly needs yp");
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

void lagpol_device(Ctx &ctx, int axis, double *d_u, int nx, int ny, int nz) {
  Ctx::IbmAxis &G = ctx.ibm[axis];
  if (!G.set) throw Error("iibm = 2: x3d_set_ibm_geometry has not been called for this direction");
  const int n[3] = {nx, ny, nz};
  const int a_ax = axis == 0 ? 1 : 0, b_ax = axis == 2 ? 1 : 2;
  if (G.na != n[a_ax] || G.nb != n[b_ax]) throw Error("iibm = 2: the geometry arrays do not match the pencil");
  if (axis == 1 && G.ncoords < ny) throw Error("iibm = 2: yp is shorter than the line");
  const long long st[3] = {1, nx, static_cast<long long>(nx) * ny};
  LagArgs g{};
  g.nl = n[axis]; g.na = G.na; g.nb = G.nb;
  g.sl = st[axis]; g.sa = st[a_ax]; g.sb = st[b_ax];
  g.nobjmax = G.nobjmax; g.npif = G.npif; g.izap = G.izap; g.searched = axis == 1 ? 1 : 0;
  g.d = G.d; g.len = G.len;
  g.nobj = static_cast<const int *>(G.nobj.p); g.nipif = static_cast<const int *>(G.nipif.p); g.nfpif = static_cast<const int *>(G.nfpif.p);
  g.xi = static_cast<const double *>(G.xi.p); g.xf = static_cast<const double *>(G.xf.p); g.coords = static_cast<const double *>(G.coords.p);
  const long long nl = static_cast<long long>(G.na) * G.nb;
  ProfScope ps(ctx, "ibm_lagpol(k_lagpol)");
  k_lagpol<<<static_cast<unsigned>((nl + 127) / 128), 128, 0, ctx.stream>>>(d_u, g);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

// host-or-device entry (lagpolx(u) etc.)
void lagpol(Ctx &ctx, int axis, double *u, int nx, int ny, int nz) {
  X3D_CUDA(cudaSetDevice(ctx.device));
  const size_t bytes = static_cast<size_t>(nx) * ny * nz * sizeof(double);
  if (bytes == 0) return;
  if (is_device_ptr(u)) { lagpol_device(ctx, axis, u, nx, ny, nz); return; }
  ctx.stage_in.reserve(bytes);
  X3D_CUDA(cudaMemcpyAsync(ctx.stage_in.p, u, bytes, cudaMemcpyHostToDevice, ctx.stream));
  lagpol_device(ctx, axis, static_cast<double *>(ctx.stage_in.p), nx, ny, nz);
  X3D_CUDA(cudaMemcpyAsync(u, ctx.stage_in.p, bytes, cudaMemcpyDeviceToHost, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}

}  // namespace x3d
