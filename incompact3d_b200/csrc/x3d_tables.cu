// x3d_tables.cu -- host-side construction of the device operator descriptors:
//   * interior stencil coefficients and explicit boundary-row tables for each of the
//     reference's operator variants (src/derive.f90, src/filters.f90), and
//   * the packed tridiagonal tables: LU arrays of prepare() (src/schemes.f90:413-439)
//     + the per-chunk "spike" products that let a line be swept by many threads, and
//     the precomputed Sherman-Morrison vector of the periodic variants
//     (src/derive.f90:30-59 recomputes it for every line; it is line independent).
#include <cmath>
#include <cstring>
#include "x3d_ctx.cuh"

namespace x3d {

int op_n_in(const OpCall &c) { return (c.kind == DPV || c.kind == IPV) ? c.nm : c.n; }
int op_n_out(const OpCall &c) { return (c.kind == DVP || c.kind == IVP) ? c.nm : c.n; }

namespace {

struct Tap { int off; double w; };

// interior stencil as (offset, weight) taps
int interior_taps(Kind k, double c0, const double c[4], Tap *t) {
  int n = 0;
  switch (k) {
    case D1: t[n++] = {1, c[0]}; t[n++] = {-1, -c[0]}; t[n++] = {2, c[1]}; t[n++] = {-2, -c[1]}; break;
    case D2: for (int q = 1; q <= 4; ++q) { t[n++] = {q, c[q - 1]}; t[n++] = {0, -2.0 * c[q - 1]}; t[n++] = {-q, c[q - 1]}; } break;
    case FIL: t[n++] = {0, c0}; for (int q = 1; q <= 3; ++q) { t[n++] = {q, c[q - 1]}; t[n++] = {-q, c[q - 1]}; } break;
    case DVP: t[n++] = {1, c[0]}; t[n++] = {0, -c[0]}; t[n++] = {2, c[1]}; t[n++] = {-1, -c[1]}; break;
    case IVP: for (int q = 1; q <= 4; ++q) { t[n++] = {q, c[q - 1]}; t[n++] = {-q + 1, c[q - 1]}; } break;
    case DPV: t[n++] = {0, c[0]}; t[n++] = {-1, -c[0]}; t[n++] = {1, c[1]}; t[n++] = {-2, -c[1]}; break;
    case IPV: for (int q = 1; q <= 4; ++q) { t[n++] = {q - 1, c[q - 1]}; t[n++] = {-q, c[q - 1]}; } break;
  }
  return n;
}

enum Ghost { G_NODE, G_HALF };

struct RowBuilder {
  int n_in;
  bool at_end;
  double *row;  // NBCOL weights
  void clear() { for (int q = 0; q < NBCOL; ++q) row[q] = 0.0; }
  void add(int idx, double w) {
    const int col = at_end ? idx - (n_in - NBCOL) : idx;
    if (col < 0 || col >= NBCOL) throw Error("boundary row reaches outside its table (line too short?)");
    row[col] += w;
  }
  // add through the mirror rule
  void add_ghost(int idx, double w, Ghost g, double sg) {
    if (idx < 0) {
      if (g == G_NODE) add(-idx, sg * w); else add(-idx - 1, w);
    } else if (idx > n_in - 1) {
      if (g == G_NODE) add(2 * (n_in - 1) - idx, sg * w); else add(2 * n_in - 1 - idx, w);
    } else {
      add(idx, w);
    }
  }
};

}  // namespace

void build_devop(const Ctx &ctx, const OpCall &call, DevOp &op) {
  std::memset(&op, 0, sizeof(op));
  const Kind k = call.kind;
  const int axis = call.axis;
  op.kind = k;
  op.n_in = op_n_in(call);
  op.n_out = op_n_out(call);
  op.rhs_only = call.rhs_only ? 1 : 0;
  op.has_post = call.post ? 1 : 0;
  const bool colloc = (k == D1 || k == D2 || k == FIL);
  if (colloc ? !(k == FIL ? ctx.have_fc[axis] : ctx.have_dc[axis]) : !ctx.have_dc[axis])
    throw Error("operator called before x3d_set_deriv_coeffs/x3d_set_filter_coeffs for this axis");
  const x3d_deriv_coeffs &c = ctx.dc[axis];
  const x3d_filter_coeffs &f = ctx.fc[axis];
  const bool per = colloc ? (call.ncl1 == 0 && call.ncln == 0) : call.periodic;
  op.periodic = per ? 1 : 0;
  switch (k) {
    case D1: op.c[0] = c.afi; op.c[1] = c.bfi; op.alpha = c.alfai; break;
    case D2: op.c[0] = c.asi; op.c[1] = c.bsi; op.c[2] = c.csi; op.c[3] = c.dsi; op.alpha = c.alsai; break;
    case FIL: op.c0 = f.fiai; op.c[0] = f.fibi; op.c[1] = f.fici; op.c[2] = f.fidi; op.alpha = f.fiali; break;
    case DVP: case DPV: op.c[0] = c.aci6; op.c[1] = c.bci6; op.alpha = c.alcai6; break;
    case IVP: case IPV: op.c[0] = c.aici6; op.c[1] = c.bici6; op.c[2] = c.cici6; op.c[3] = c.dici6; op.alpha = c.ailcai6; break;
  }
  // npaire values the reference actually implements (derive.f90)
  const int np = call.npaire;
  bool supported = true;
  if (colloc) {
    if ((call.ncl1 == 1 || call.ncln == 1) && np != 0 && np != 1) supported = false;
  } else if (!per) {
    if (k == DVP) supported = (axis == 1) ? (np == 0) : (np == 0 || np == 1);
    else supported = (np == 1);
  }
  op.untouched = supported ? 0 : 1;
  if (per || !supported) { op.nb = 0; return; }
  op.nb = NBROW;
  if (op.n_in < NBCOL || op.n_out < 2 * NBROW) throw Error("non-periodic line too short for the compact closures (need >= 9 points)");

  Tap taps[16];
  const int nt = interior_taps(k, op.c0, op.c, taps);
  const Ghost g = (k == DPV || k == IPV) ? G_HALF : G_NODE;
  const double sg = colloc ? (np == 1 ? 1.0 : -1.0) : 1.0;
  for (int end = 0; end < 2; ++end) {
    for (int r = 0; r < NBROW; ++r) {
      RowBuilder rb{op.n_in, end == 1, end ? op.wend[r] : op.wstart[r]};
      rb.clear();
      const int row = end ? op.n_out - NBROW + r : r;
      const int ncl = colloc ? (end ? call.ncln : call.ncl1) : 1;
      // distance from the boundary (0 = boundary row)
      const int dist = end ? (op.n_out - 1 - row) : row;
      const int N = op.n_in;
      bool done = false;
      if (colloc && ncl == 2) {
        // one-sided Dirichlet closures; e(j) = j-th point counted from the boundary
        auto e = [&](int j) { return end ? N - 1 - j : j; };
        const double sgn = 1.0;
        (void)sgn;
        if (k == D1) {
          if (dist == 0) {  // derive.f90:233 / :171 (mirrored with opposite sign)
            if (!end) { rb.add(e(0), c.af1); rb.add(e(1), c.bf1); rb.add(e(2), c.cf1); }
            else { rb.add(e(0), -c.afn); rb.add(e(1), -c.bfn); rb.add(e(2), -c.cfn); }
            done = true;
          } else if (dist == 1) {  // :234 / :170
            const double a2 = end ? c.afm : c.af2;
            rb.add(row + 1, a2); rb.add(row - 1, -a2);
            done = true;
          }
        } else if (k == D2) {  // derive.f90:1999-2036
          if (dist == 0) {
            if (!end) { rb.add(e(0), c.as1); rb.add(e(1), c.bs1); rb.add(e(2), c.cs1); rb.add(e(3), c.ds1); }
            else { rb.add(e(0), c.asn); rb.add(e(1), c.bsn); rb.add(e(2), c.csn); rb.add(e(3), c.dsn); }
          } else {
            double cc[3] = {0, 0, 0};
            if (dist == 1) cc[0] = end ? c.asm_ : c.as2;
            if (dist == 2) { cc[0] = end ? c.ast : c.as3; cc[1] = end ? c.bst : c.bs3; }
            if (dist == 3) { cc[0] = end ? c.astt : c.as4; cc[1] = end ? c.bstt : c.bs4; cc[2] = end ? c.cstt : c.cs4; }
            for (int q = 1; q <= 3; ++q)
              if (cc[q - 1] != 0.0) { rb.add(row + q, cc[q - 1]); rb.add(row, -2.0 * cc[q - 1]); rb.add(row - q, cc[q - 1]); }
          }
          done = true;
        } else {  // FIL, filters.f90:577-591
          if (dist == 0) { rb.add(e(0), 1.0); done = true; }
          else if (dist == 1) {
            if (!end) { rb.add(e(0), f.fia2); rb.add(e(1), f.fib2); rb.add(e(2), f.fic2); rb.add(e(3), f.fid2); }
            else { rb.add(e(0), f.fiam); rb.add(e(1), f.fibm); rb.add(e(2), f.ficm); rb.add(e(3), f.fidm); }
            done = true;
          } else if (dist == 2) {
            if (!end) { rb.add(e(0), f.fia3); rb.add(e(1), f.fib3); rb.add(e(2), f.fic3); rb.add(e(3), f.fid3); rb.add(e(4), f.fie3); rb.add(e(5), f.fif3); }
            else { rb.add(e(0), f.fiap); rb.add(e(1), f.fibp); rb.add(e(2), f.ficp); rb.add(e(3), f.fidp); rb.add(e(4), f.fiep); rb.add(e(5), f.fifp); }
            done = true;
          }
        }
      }
      if (done) continue;
      // forced zero rows
      if (colloc && ncl == 1 && np == 0 && dist == 0 && (k == D2 || k == FIL)) continue;  // derive.f90:1590,1649; filters.f90:349,361
      if (k == DPV && dist == 0) continue;                                               // derive.f90:4099,4108
      if (k == DVP && np == 0 && dist == 0) {  // derive.f90:3882-3883, 3892-3893
        const double a = op.c[0], b = op.c[1];
        if (!end) { rb.add(0, -a - 2.0 * b); rb.add(1, a + b); rb.add(2, b); }
        else { rb.add(N - 1, a + 2.0 * b); rb.add(N - 2, -a - b); rb.add(N - 3, -b); }
        continue;
      }
      // ghost-mirrored interior stencil
      for (int q = 0; q < nt; ++q) rb.add_ghost(row + taps[q].off, taps[q].w, g, sg);
      // derive.f90:1607-1614,1625-1632: in der??_11 (npaire=0) the c-term of the 4th row reads
      // -u(boundary) where the mirror rule gives +u(boundary)
      if (k == D2 && ncl == 1 && np == 0 && dist == 3) rb.add(end ? N - 1 : 0, -2.0 * op.c[2]);
    }
  }
}

// ---------------------------------------------------------------------------
TriTable::~TriTable() {
  if (d_rows) cudaFree(d_rows);
  if (d_scan) cudaFree(d_scan);
  if (d_chunk) cudaFree(d_chunk);
  if (d_rows_c) cudaFree(d_rows_c);
}

// Chunks far from both ends of a line have identical table rows (the LU recurrence of prepare() has converged) and a
// zero Sherman-Morrison entry: keep `h` head chunks, one generic chunk and h + 1 tail chunks (the last chunk may be
// partial).  h is the smallest count for which every column of every middle chunk equals the generic one (sixth-order
// derivatives: ~45 rows; the alpha = 0.49 interpolators: ~90 rows); the Sherman-Morrison column gets its own, larger
// count.  Returns h, or -1 when the table does not compress into the budget of the caller.
int compress_tri(const TriTable &T) {
  if (T.c_head != 0) return T.c_head;
  const int L = T.L, nc = T.nc;
  T.c_head = -1;
  if (T.h_rows.empty()) return -1;
  auto R = [&](int row, int col) { return T.h_rows[static_cast<size_t>(row) * TRI_W + col]; };
  const int cols[6] = {T_S, T_PF, T_W, T_FW, T_PB, T_POST};
  auto same_from = [&](int h) {
    for (int c = h + 1; c <= nc - h - 2; ++c)
      for (int m = 0; m < L; ++m)
        for (int q = 0; q < 6; ++q) {
          const double a = R(c * L + m, cols[q]), b = R(h * L + m, cols[q]);
          if (std::fabs(a - b) > 4e-16 * std::max(std::fabs(a), std::fabs(b))) return false;
        }
    return true;
  };
  double rsmax = 0.0;
  for (int i = 0; i < nc * L; ++i) rsmax = std::max(rsmax, std::fabs(R(i, T_RS)));
  auto rs_zero_from = [&](int h) {
    for (int c = h; c <= nc - h - 2; ++c)
      for (int m = 0; m < L; ++m)
        if (std::fabs(R(c * L + m, T_RS)) > 1e-18 * rsmax) return false;
    return true;
  };
  int h = -1, hr = -1;
  for (int t = std::max(1, (48 + L - 1) / L); t <= 4 && nc >= 2 * t + 3; ++t)
    if (same_from(t)) { h = t; break; }
  for (int t = 1; t <= 8 && nc >= 2 * t + 3; ++t)
    if (rs_zero_from(t)) { hr = t; break; }
  if (h < 0 || hr < 0) return -1;
  auto chunk_of = [&](int t, int hh) { return t < hh ? t : (t == hh ? hh : nc - hh - 1 + (t - hh - 1)); };
  const int np = (2 * h + 2) * L, npr = (2 * hr + 2) * L;
  std::vector<double> img(static_cast<size_t>(7) * np + npr, 0.0);
  for (int t = 0; t < 2 * h + 2; ++t)
    for (int m = 0; m < L; ++m)
      for (int col = 0; col < 7; ++col) img[static_cast<size_t>(col) * np + t * L + m] = R(chunk_of(t, h) * L + m, col);
  for (int t = 0; t < 2 * hr + 2; ++t)
    for (int m = 0; m < L; ++m) img[static_cast<size_t>(7) * np + t * L + m] = (t == hr) ? 0.0 : R(chunk_of(t, hr) * L + m, T_RS);
  X3D_CUDA(cudaMalloc(&T.d_rows_c, img.size() * sizeof(double)));
  X3D_CUDA(cudaMemcpy(T.d_rows_c, img.data(), img.size() * sizeof(double), cudaMemcpyHostToDevice));
  T.c_head = h;
  T.c_head_rs = hr;
  return h;
}

static uint64_t fnv(const void *p, size_t n, uint64_t h) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

const TriTable &get_tri(Ctx &ctx, const double *f, const double *s, const double *w, int n, int L, bool periodic,
                        double alpha, const double *post) {
  uint64_t h = 1469598103934665603ull;
  h = fnv(f, n * sizeof(double), h);
  h = fnv(s, n * sizeof(double), h);
  h = fnv(w, n * sizeof(double), h);
  if (post) h = fnv(post, n * sizeof(double), h);
  const int meta[4] = {n, L, periodic ? 1 : 0, post ? 1 : 0};
  h = fnv(meta, sizeof(meta), h);
  h = fnv(&alpha, sizeof(alpha), h);
  auto it = ctx.tri_cache.find(h);
  if (it != ctx.tri_cache.end()) return *it->second;

  auto T = std::make_unique<TriTable>();
  T->n = n; T->L = L; T->nc = (n + L - 1) / L;
  const int nc = T->nc, np = nc * L;
  std::vector<double> rows(static_cast<size_t>(np) * TRI_W, 0.0);
  auto R = [&](int i, int col) -> double & { return rows[static_cast<size_t>(i) * TRI_W + col]; };
  for (int i = 0; i < n; ++i) {
    R(i, T_S) = (i == 0) ? 0.0 : s[i];
    R(i, T_W) = w[i];
    R(i, T_FW) = (i == n - 1) ? 0.0 : f[i] * w[i];
    R(i, T_POST) = post ? post[i] : 1.0;
  }
  std::vector<double> chunk(2 * nc, 0.0);
  for (int c = 0; c < nc; ++c) {
    const int cs = c * L, ce = std::min(cs + L, n) - 1;
    double p = 1.0;
    for (int i = cs; i <= ce; ++i) { p *= -R(i, T_S); R(i, T_PF) = p; }
    chunk[c] = p;  // Af(c)
    p = 1.0;
    for (int i = ce; i >= cs; --i) { p *= -R(i, T_FW); R(i, T_PB) = p; }
    chunk[nc + c] = p;  // Ab(c)
  }
  if (periodic) {
    // r solves A' r = (-1,0,...,0,alpha)^T  (src/derive.f90:30-54)
    std::vector<double> r(n, 0.0);
    r[0] = -1.0; r[n - 1] = alpha;
    for (int i = 1; i < n; ++i) r[i] = r[i] - r[i - 1] * s[i];
    r[n - 1] = r[n - 1] * w[n - 1];
    for (int i = n - 2; i >= 0; --i) r[i] = (r[i] - f[i] * r[i + 1]) * w[i];
    const double den = 1.0 + r[0] - alpha * r[n - 1];
    for (int i = 0; i < n; ++i) R(i, T_RS) = r[i] / den;
  }
  // Kogge-Stone multipliers for warp-per-line kernels (lane = chunk)
  std::vector<double> scan(10 * 32, 0.0);
  if (nc <= 32) {
    for (int lev = 0; lev < 5; ++lev) {
      const int d = 1 << lev;
      for (int c = 0; c < nc; ++c) {
        if (c - d >= 0) { double p = 1.0; for (int m = 0; m < d; ++m) p *= chunk[c - m]; scan[lev * 32 + c] = p; }
        if (c + d <= nc - 1) { double p = 1.0; for (int m = 0; m < d; ++m) p *= chunk[nc + c + m]; scan[(5 + lev) * 32 + c] = p; }
      }
    }
  }
  X3D_CUDA(cudaMalloc(&T->d_rows, rows.size() * sizeof(double)));
  X3D_CUDA(cudaMalloc(&T->d_scan, scan.size() * sizeof(double)));
  X3D_CUDA(cudaMalloc(&T->d_chunk, chunk.size() * sizeof(double)));
  X3D_CUDA(cudaMemcpyAsync(T->d_rows, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(T->d_scan, scan.data(), scan.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(T->d_chunk, chunk.data(), chunk.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));  // host vectors go out of scope
  T->h_rows = std::move(rows);
  T->h_scan = std::move(scan);
  const TriTable &ref = *T;
  ctx.tri_cache[h] = std::move(T);
  return ref;
}

}  // namespace x3d
