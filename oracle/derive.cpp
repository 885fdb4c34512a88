// derive.cpp -- oracle restatement of src/derive.f90 (42 operators) and the 15
// filter kernels of src/filters.f90 (TEST INFRASTRUCTURE, see x3d_oracle.hpp).
//
// The reference writes every boundary row out by hand.  Here each family is one
// routine: the line is extended by ghost values that reproduce the hand-written
// rows EXACTLY (periodic wrap; +u mirror for npaire=1; -u mirror for npaire=0 --
// x-(-y) and x+(-y) round like x+y and x-y), and the rows that are not of the
// ghost form (one-sided Dirichlet closures, forced zeros, the "2*u_b - u" rows of
// der?vp, the -u(boundary) quirk of der??_11 npaire=0) are overwritten
// explicitly.  Lines are processed in panels of W so the compiler vectorises over
// independent lines, as the reference's inner i-loops do.
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include "x3d_oracle.hpp"

namespace x3do {

namespace {
constexpr int W = 8;  // lines per panel
constexpr int H = 4;  // ghost width

struct Panel {
  // E: ghost-extended input, rows -H .. n_in+H-1 ; T: output rows 0..n_out-1
  std::vector<double> Ebuf, Tbuf, Rbuf;
  double *E = nullptr, *T = nullptr;
  void size(int n_in, int n_out) {
    Ebuf.assign(static_cast<size_t>(n_in + 2 * H) * W, 0.0);
    Tbuf.assign(static_cast<size_t>(n_out) * W, 0.0);
    E = Ebuf.data() + H * W;
    T = Tbuf.data();
  }
};

#define LANES for (int l = 0; l < W; ++l)

inline void fill_ghost_periodic(double *E, int n) {
  for (int k = 1; k <= H; ++k) LANES {
    E[(-k) * W + l] = E[(n - k) * W + l];
    E[(n - 1 + k) * W + l] = E[(k - 1) * W + l];
  }
}
// mirror about the NODE 0 / NODE n-1 (velocity-type arrays): u(-k) = sg*u(k)
inline void fill_ghost_node_start(double *E, double sg) {
  for (int k = 1; k <= H; ++k) LANES E[(-k) * W + l] = sg * E[k * W + l];
}
inline void fill_ghost_node_end(double *E, int n, double sg) {
  for (int k = 1; k <= H; ++k) LANES E[(n - 1 + k) * W + l] = sg * E[(n - 1 - k) * W + l];
}
// mirror about the half node before point 0 / after point n-1 (pressure-type arrays)
inline void fill_ghost_half_start(double *E) {
  for (int k = 1; k <= H; ++k) LANES E[(-k) * W + l] = E[(k - 1) * W + l];
}
inline void fill_ghost_half_end(double *E, int n) {
  for (int k = 1; k <= H; ++k) LANES E[(n - 1 + k) * W + l] = E[(n - k) * W + l];
}

// Thomas sweeps, src/derive.f90:45-54 (and every other operator)
inline void thomas(double *T, int n, const double *f, const double *s, const double *w) {
  for (int i = 1; i < n; ++i) LANES T[i * W + l] = T[i * W + l] - T[(i - 1) * W + l] * s[i];
  LANES T[(n - 1) * W + l] = T[(n - 1) * W + l] * w[n - 1];
  for (int i = n - 2; i >= 0; --i) LANES T[i * W + l] = (T[i * W + l] - f[i] * T[(i + 1) * W + l]) * w[i];
}
// Sherman-Morrison correction of the periodic operators, src/derive.f90:30-59
inline void periodic_solve(double *T, int n, const double *f, const double *s, const double *w, double alfa) {
  std::vector<double> r(n, 0.0);
  r[0] = -1.0;
  r[n - 1] = alfa;
  for (int i = 1; i < n; ++i) r[i] = r[i] - r[i - 1] * s[i];
  r[n - 1] = r[n - 1] * w[n - 1];
  for (int i = n - 2; i >= 0; --i) r[i] = (r[i] - f[i] * r[i + 1]) * w[i];
  thomas(T, n, f, s, w);
  LANES {
    const double sl = (T[l] - alfa * T[(n - 1) * W + l]) / (1.0 + r[0] - alfa * r[n - 1]);
    for (int i = 0; i < n; ++i) T[i * W + l] = T[i * W + l] - sl * r[i];
  }
}

#define U(i) E[(i) * W + l]

// ---- first derivative, src/derive.f90:7-1350 ---------------------------------
bool rhs_d1(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n;
  const auto &c = *op.c;
  const bool has11 = (op.ncl1 == 1 || op.ncln == 1);
  if (has11 && op.npaire != 0 && op.npaire != 1) return false;  // both "if (npaire==..)" skipped
  const double sg = op.npaire == 1 ? 1.0 : -1.0;
  if (op.ncl1 == 0) fill_ghost_periodic(E, n);
  if (op.ncl1 == 1) fill_ghost_node_start(E, sg);
  if (op.ncln == 1) fill_ghost_node_end(E, n, sg);
  const double a = c.afi, b = c.bfi;
  for (int i = 0; i < n; ++i) LANES T[i * W + l] = a * (U(i + 1) - U(i - 1)) + b * (U(i + 2) - U(i - 2));
  if (op.ncl1 == 2) LANES {  // :233-234
    T[0 * W + l] = c.af1 * U(0) + c.bf1 * U(1) + c.cf1 * U(2);
    T[1 * W + l] = c.af2 * (U(2) - U(0));
  }
  if (op.ncln == 2) LANES {  // :170-171
    T[(n - 2) * W + l] = c.afm * (U(n - 1) - U(n - 3));
    T[(n - 1) * W + l] = (-c.afn * U(n - 1)) - c.bfn * U(n - 2) - c.cfn * U(n - 3);
  }
  return true;
}

// ---- second derivative, src/derive.f90:1354-3791 ------------------------------
inline double d2term(const double *E, int i, int k, int l) {
  return U(i + k) - U(i) - U(i) + U(i - k);
}
bool rhs_d2(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n;
  const auto &c = *op.c;
  const bool has11 = (op.ncl1 == 1 || op.ncln == 1);
  if (has11 && op.npaire != 0 && op.npaire != 1) return false;
  const double sg = op.npaire == 1 ? 1.0 : -1.0;
  if (op.ncl1 == 0) fill_ghost_periodic(E, n);
  if (op.ncl1 == 1) fill_ghost_node_start(E, sg);
  if (op.ncln == 1) fill_ghost_node_end(E, n, sg);
  const double a = c.asi, b = c.bsi, cc = c.csi, d = c.dsi;
  for (int i = 0; i < n; ++i) LANES
    T[i * W + l] = a * d2term(E, i, 1, l) + b * d2term(E, i, 2, l) + cc * d2term(E, i, 3, l) + d * d2term(E, i, 4, l);
  if (op.ncl1 == 1 && op.npaire == 0) LANES {
    T[0 * W + l] = 0.0;  // :1590
    // :1607-1614  row 4 (1-based): the c-term reads -ux(1), not the mirror value
    const int i = 3;
    T[i * W + l] = a * d2term(E, i, 1, l) + b * d2term(E, i, 2, l) + cc * (U(i + 3) - U(i) - U(i) - U(0)) +
                   d * d2term(E, i, 4, l);
  }
  if (op.ncln == 1 && op.npaire == 0) LANES {
    T[(n - 1) * W + l] = 0.0;  // :1649
    // :1625-1632  row nx-3: the c-term reads -ux(nx)
    const int i = n - 4;
    T[i * W + l] = a * d2term(E, i, 1, l) + b * d2term(E, i, 2, l) + cc * (-U(n - 1) - U(i) - U(i) + U(i - 3)) +
                   d * d2term(E, i, 4, l);
  }
  if (op.ncl1 == 2) LANES {  // :1999-2012
    T[0 * W + l] = c.as1 * U(0) + c.bs1 * U(1) + c.cs1 * U(2) + c.ds1 * U(3);
    T[1 * W + l] = c.as2 * d2term(E, 1, 1, l);
    T[2 * W + l] = c.as3 * d2term(E, 2, 1, l) + c.bs3 * d2term(E, 2, 2, l);
    T[3 * W + l] = c.as4 * d2term(E, 3, 1, l) + c.bs4 * d2term(E, 3, 2, l) + c.cs4 * d2term(E, 3, 3, l);
  }
  if (op.ncln == 2) LANES {  // :2023-2036
    T[(n - 4) * W + l] = c.astt * d2term(E, n - 4, 1, l) + c.bstt * d2term(E, n - 4, 2, l) + c.cstt * d2term(E, n - 4, 3, l);
    T[(n - 3) * W + l] = c.ast * d2term(E, n - 3, 1, l) + c.bst * d2term(E, n - 3, 2, l);
    T[(n - 2) * W + l] = c.asm_ * d2term(E, n - 2, 1, l);
    T[(n - 1) * W + l] = c.asn * U(n - 1) + c.bsn * U(n - 2) + c.csn * U(n - 3) + c.dsn * U(n - 4);
  }
  return true;
}

// ---- filters, src/filters.f90:221-1377 ------------------------------------------
bool rhs_fil(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n;
  const auto &c = *op.fc;
  const bool has11 = (op.ncl1 == 1 || op.ncln == 1);
  if (has11 && op.npaire != 0 && op.npaire != 1) return false;
  const double sg = op.npaire == 1 ? 1.0 : -1.0;
  if (op.ncl1 == 0) fill_ghost_periodic(E, n);
  if (op.ncl1 == 1) fill_ghost_node_start(E, sg);
  if (op.ncln == 1) fill_ghost_node_end(E, n, sg);
  for (int i = 0; i < n; ++i) LANES
    T[i * W + l] = c.fiai * U(i) + c.fibi * (U(i + 1) + U(i - 1)) + c.fici * (U(i + 2) + U(i - 2)) +
                   c.fidi * (U(i + 3) + U(i - 3));
  if (op.ncl1 == 1 && op.npaire == 0) LANES T[l] = 0.0;                      // :349
  if (op.ncln == 1 && op.npaire == 0) LANES T[(n - 1) * W + l] = 0.0;        // :361
  if (op.ncl1 == 2) LANES {  // :577-581
    T[0 * W + l] = U(0);
    T[1 * W + l] = c.fia2 * U(0) + c.fib2 * U(1) + c.fic2 * U(2) + c.fid2 * U(3);
    T[2 * W + l] = c.fia3 * U(0) + c.fib3 * U(1) + c.fic3 * U(2) + c.fid3 * U(3) + c.fie3 * U(4) + c.fif3 * U(5);
  }
  if (op.ncln == 2) LANES {  // :587-591
    T[(n - 1) * W + l] = U(n - 1);
    T[(n - 2) * W + l] = c.fiam * U(n - 1) + c.fibm * U(n - 2) + c.ficm * U(n - 3) + c.fidm * U(n - 4);
    T[(n - 3) * W + l] = c.fiap * U(n - 1) + c.fibp * U(n - 2) + c.ficp * U(n - 3) + c.fidp * U(n - 4) +
                         c.fiep * U(n - 5) + c.fifp * U(n - 6);
  }
  return true;
}

// ---- staggered: velocity -> pressure mesh, src/derive.f90:3796,3911,4265,4442,4920,5105 ----
bool rhs_dvp(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n, nm = op.nm;
  const auto &c = *op.c;
  const double a = c.aci6, b = c.bci6;
  if (op.periodic) {
    fill_ghost_periodic(E, n);
  } else {
    if (op.npaire != 0 && op.npaire != 1) return false;
    fill_ghost_node_start(E, 1.0);
    fill_ghost_node_end(E, n, 1.0);
  }
  for (int i = 0; i < nm; ++i) LANES T[i * W + l] = a * (U(i + 1) - U(i)) + b * (U(i + 2) - U(i - 1));
  if (!op.periodic && op.npaire == 0) LANES {  // :3882-3893
    T[l] = a * (U(1) - U(0)) + b * (U(2) - 2.0 * U(0) + U(1));
    T[(nm - 1) * W + l] = a * (U(n - 1) - U(nm - 1)) + b * (2.0 * U(n - 1) - U(nm - 1) - U(nm - 2));
  }
  return true;
}
bool rhs_ivp(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n, nm = op.nm;
  const auto &c = *op.c;
  if (op.periodic) {
    fill_ghost_periodic(E, n);
  } else {
    if (op.npaire != 1) return false;  // only npaire==1 exists, :3991
    fill_ghost_node_start(E, 1.0);
    fill_ghost_node_end(E, n, 1.0);
  }
  for (int i = 0; i < nm; ++i) LANES
    T[i * W + l] = c.aici6 * (U(i + 1) + U(i)) + c.bici6 * (U(i + 2) + U(i - 1)) + c.cici6 * (U(i + 3) + U(i - 2)) +
                   c.dici6 * (U(i + 4) + U(i - 3));
  return true;
}
// ---- staggered: pressure -> velocity mesh, src/derive.f90:4041,4126,4587,4775,5287,5426 ----
bool rhs_dpv(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n, nm = op.nm;
  const auto &c = *op.c;
  const double a = c.aci6, b = c.bci6;
  if (op.periodic) {
    fill_ghost_periodic(E, nm);
  } else {
    if (op.npaire != 1) return false;  // :4096
    fill_ghost_half_start(E);
    fill_ghost_half_end(E, nm);
  }
  for (int i = 0; i < n; ++i) LANES T[i * W + l] = a * (U(i) - U(i - 1)) + b * (U(i + 1) - U(i - 2));
  if (!op.periodic) LANES {  // :4099,:4108
    T[l] = 0.0;
    T[(n - 1) * W + l] = 0.0;
  }
  return true;
}
bool rhs_ipv(const OpDesc &op, const double *Ein, double *T) {
  double *E = const_cast<double *>(Ein);
  const int n = op.n, nm = op.nm;
  const auto &c = *op.c;
  if (op.periodic) {
    fill_ghost_periodic(E, nm);
  } else {
    if (op.npaire != 1) return false;  // :4207
    fill_ghost_half_start(E);
    fill_ghost_half_end(E, nm);
  }
  for (int i = 0; i < n; ++i) LANES
    T[i * W + l] = c.aici6 * (U(i) + U(i - 1)) + c.bici6 * (U(i + 1) + U(i - 2)) + c.cici6 * (U(i + 2) + U(i - 3)) +
                   c.dici6 * (U(i + 3) + U(i - 4));
  return true;
}
#undef U

double sm_alpha(const OpDesc &op) {
  switch (op.kind) {
    case D1: return op.c->alfai;
    case D2: return op.c->alsai;
    case FIL: return op.fc->fiali;
    case DVP: case DPV: return op.c->alcai6;
    default: return op.c->ailcai6;
  }
}
bool is_periodic(const OpDesc &op) {
  if (op.kind == D1 || op.kind == D2 || op.kind == FIL) return op.ncl1 == 0 && op.ncln == 0;
  return op.periodic;
}
}  // namespace

int op_n_in(const OpDesc &op) {
  if (op.kind == DPV || op.kind == IPV) return op.nm;
  return op.n;
}
int op_n_out(const OpDesc &op) {
  if (op.kind == DVP || op.kind == IVP) return op.nm;
  return op.n;
}

void apply_op(const OpDesc &op, int axis, const int dims_in[3], const double *u, double *t) {
  const int n_in = op_n_in(op), n_out = op_n_out(op);
  if (dims_in[axis] != n_in) throw std::runtime_error("apply_op: line extent mismatch");
  int dims_out[3] = {dims_in[0], dims_in[1], dims_in[2]};
  dims_out[axis] = n_out;
  const size_t tot_out = static_cast<size_t>(dims_out[0]) * dims_out[1] * dims_out[2];
  // n == 1 shortcuts of the z operators (derive.f90:874,2939,4937,5121,5305,5445)
  if (op.n == 1 && axis == 2) {
    if (op.kind == IVP || op.kind == IPV) {
      if (op.nm == 1) { std::memcpy(t, u, tot_out * sizeof(double)); return; }
    } else if (op.kind == FIL) {  // filters.f90:1008-1011 (filz_00 only)
      if (op.ncl1 == 0 && op.ncln == 0) { std::memcpy(t, u, tot_out * sizeof(double)); return; }
    } else {
      std::fill(t, t + tot_out, 0.0);
      return;
    }
  }
  // deryvp implements only npaire==0 in its non-periodic branch (derive.f90:4525)
  const bool force_skip = (op.kind == DVP && axis == 1 && !op.periodic && op.npaire != 0);
  // line addressing: element q of line (a,b) sits at base(a,b) + q*stride
  std::ptrdiff_t s_in, s_out;
  long nlines_fast, nlines_slow;        // fast index contiguous in memory (except axis 0)
  std::ptrdiff_t fast_in, fast_out, slow_in, slow_out;
  if (axis == 0) {
    s_in = s_out = 1;
    nlines_fast = static_cast<long>(dims_in[1]) * dims_in[2]; nlines_slow = 1;
    fast_in = n_in; fast_out = n_out; slow_in = slow_out = 0;
  } else if (axis == 1) {
    s_in = s_out = dims_in[0];
    nlines_fast = dims_in[0]; nlines_slow = dims_in[2];
    fast_in = fast_out = 1;
    slow_in = static_cast<std::ptrdiff_t>(dims_in[0]) * n_in;
    slow_out = static_cast<std::ptrdiff_t>(dims_in[0]) * n_out;
  } else {
    s_in = s_out = static_cast<std::ptrdiff_t>(dims_in[0]) * dims_in[1];
    nlines_fast = static_cast<long>(dims_in[0]) * dims_in[1]; nlines_slow = 1;
    fast_in = fast_out = 1; slow_in = slow_out = 0;
  }
  const bool per = is_periodic(op);
  const double alfa = sm_alpha(op);
  const long npanels = (nlines_fast + W - 1) / W;
#pragma omp parallel
  {
    Panel P;
    P.size(n_in, n_out);
#pragma omp for collapse(2) schedule(static)
    for (long sl = 0; sl < nlines_slow; ++sl) {
      for (long pn = 0; pn < npanels; ++pn) {
        const long f0 = pn * W;
        const int cnt = static_cast<int>(std::min<long>(W, nlines_fast - f0));
        double *E = P.E, *T = P.T;
        for (int l = 0; l < W; ++l) {
          const long fl = f0 + std::min(l, cnt - 1);
          const double *src = u + sl * slow_in + fl * fast_in;
          for (int i = 0; i < n_in; ++i) E[i * W + l] = src[i * s_in];
        }
        bool ok = false;
        switch (op.kind) {
          case D1: ok = rhs_d1(op, E, T); break;
          case D2: ok = rhs_d2(op, E, T); break;
          case FIL: ok = rhs_fil(op, E, T); break;
          case DVP: ok = rhs_dvp(op, E, T); break;
          case IVP: ok = rhs_ivp(op, E, T); break;
          case DPV: ok = rhs_dpv(op, E, T); break;
          case IPV: ok = rhs_ipv(op, E, T); break;
        }
        if (force_skip) ok = false;
        if (!ok) {
          // unsupported npaire: the reference skips RHS and solve, leaving t untouched,
          // but still runs the trailing "if (istret.ne.0) t = t*pp" loop (derive.f90:4572-4580)
          if (op.post && !op.rhs_only)
            for (int l = 0; l < cnt; ++l) {
              double *dst = t + sl * slow_out + (f0 + l) * fast_out;
              for (int i = 0; i < n_out; ++i) dst[i * s_out] = dst[i * s_out] * op.post[i];
            }
          continue;
        }
        if (!op.rhs_only) {
          if (per) periodic_solve(T, n_out, op.f, op.s, op.w, alfa);
          else thomas(T, n_out, op.f, op.s, op.w);
          if (op.post)  // derive.f90:409-417, 4572-4580, 4905-4913
            for (int i = 0; i < n_out; ++i) LANES T[i * W + l] = T[i * W + l] * op.post[i];
        }
        for (int l = 0; l < cnt; ++l) {
          double *dst = t + sl * slow_out + (f0 + l) * fast_out;
          for (int i = 0; i < n_out; ++i) dst[i * s_out] = T[i * W + l];
        }
      }
    }
  }
}

}  // namespace x3do
