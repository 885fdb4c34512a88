// x3d_mom_kernels.cuh -- fused momentum-RHS kernels for periodic directions (sm_100a).
//
// For one direction d with advecting velocity a = u_d, the reference computes for each velocity component c
//   conv_c = D1(c a) + a D1(c)          (skew-symmetric convection, src/transeq.f90:114-146,188-219,240-274)
//   diff_c = D2(c)                      (src/transeq.f90:301-303,336-338,442-444)
// with nine operator calls and three elementwise passes over nine full fields.  Here one kernel does all
// nine line solves of a direction per tile position and writes  r_c = xnu diff_c - 1/2 conv_c : every
// velocity component is read once and every result written once (6 array passes instead of ~45).
//
// k_mom_pair (y / z lines) has the structure of k_pair: 128B-swizzled TMA tiles of 16 lanes x whole line,
// warp = lane pair, thread = chunk of L rows, shuffle carries, a TMA producer warp, mbarriers only.  A tile
// position needs three tiles (a, c1, c2) that fill the three ring slots; they are consumed in the order
// c1, c2, a (each solved in place and stored), and every slot is refilled for the next position as soon as
// its store has left shared memory, so loads stay one to two component-times ahead of their use.
//
// Coefficient tables are compressed: for a periodic operator the LU rows of prepare() (src/schemes.f90:413-439)
// reach their floating-point fixed point a few rows away from row 1, so all chunks except the first three and
// the last three share one table (x3d_mom.cu checks this on the host and otherwise disables the fused path).
#pragma once
#include "x3d_ops_kernels.cuh"

namespace x3d {

struct MomMaps {
  CUtensorMap in[3], halo[3], out[3];
};
struct MomGeom {
  int nbx;
  long long npos;       // tile positions = nbx * nouter (y/z) or ceil(nlines / 16) (x)
  int slot_bytes;
  int nbox, br;         // data boxes per tile (y/z)
  // x lines: 16 contiguous lines per tile, moved by 1-D bulk copies
  const double *fin[3];
  double *fout[3];
  long long nlines;
  int pitch;            // doubles per line in shared memory (n + 8)
  int n;                // line length
  int nc;               // chunks per line
  int ia, ic1, ic2;     // field index of the advecting velocity and of the two other components
  double xnu;
  int add;              // 0: out = r, 1: out += r (TMA reduce-add stores)
  // ---- cyclic (table-free) line solves, see pair_solve_cyclic ----
  struct Cyc {
    double rho;         // root of alpha rho^2 + rho + alpha = 0 inside the unit circle
    double rhoL, rhoR;  // rho^L, rho^rem: weight of a whole / of the last (partial) chunk in the carry look-back
    double esc;         // rho^-(L - rem): rescales what enters / leaves the padded rows of the last chunk
    double gamma, delta;// what the padded rows of the last chunk (they hold rho^(j+1) y_last after the forward sweep) add to
                        // its backward start value and to the carry that reaches its last real row, per unit of y_last
    double scale;       // what the solution is multiplied by when it enters r: xnu / c (D2), -1/2 / c (D1), c = -alpha / rho
    int K;              // chunks of look-back: |rho|^(K L) < 1e-18
  } cy1, cy2;
  int rem;              // rows of the last chunk
  // ---- long lines by overlap-save segments ----
  // A line of ntot > 544 rows does not fit a tile.  The cyclic recurrences forget their past geometrically (|rho|^k), so a
  // tile of n = seg_S + 2 seg_H consecutive rows (wrapping at the ends of the line), solved as if IT were periodic, is
  // exact to double precision on its inner seg_S rows once |rho|^(seg_H - 4) < 1e-17: what the artificial wrap and the
  // stencil rows next to it get wrong has decayed before it reaches them.  nseg tiles per line; only the inner rows
  // are stored.  nseg = 1: seg_S = n = ntot, seg_H = 0 (the whole line, truly periodic).
  int ntot, seg_S, seg_H, nseg;
  double *iu_out[3];    // INTT: where the new velocity goes (= iu unless the line is segmented: neighbours still read the old one)
  // ---- time integration folded into the x kernel (INTT), src/time_integrators.f90:71-74,151-157 ----
  //   N = sum (+ extra) + r_x ;  u <- ca N + cb old + u ;  old <- N (when store_old)
  const double *isum[3], *iextra[3], *iold_in[3];
  double *iu[3], *iold_out[3];
  double ca, cb;
  int use_old, store_old, has_extra;
};
struct MomTabs {        // device: [3][mom_tabs(L) L] double2 each ((s,Pf) (w,fw) (Pb,rs)), and [10][32] scan multipliers
  const double2 *c1, *c2;
  const double *scan1, *scan2;
  double ff1, ff2;     // the operators' constant super-diagonal (ffx = alfaix ... of src/schemes.f90)
};
// chunk tables: H head chunks, 1 generic, H+1 tail chunks (the last chunk may hold a single row); H covers >= 45 rows (the Sherman-Morrison vector of the
// 6th-order schemes decays by 0.382 per row: 0.382^45 = 1.5e-19)
__host__ __device__ constexpr int mom_head(int L) { return L >= 15 ? 3 : (L >= 9 ? 5 : 9); }
__host__ __device__ constexpr int mom_tabs(int L) { return 2 * mom_head(L) + 2; }
// 8 consumer warps (two warpgroups) + one producer warpgroup (one active thread).  Registers are allocated
// per warpgroup: the producer group shrinks to 24 registers per thread and the consumers grow to 240
// (setmaxnreg), which is what lets a thread hold two 17-row chunk pairs without spilling.
constexpr int MOM_THREADS = 32 * (PAIR_WARPS + 4);

// periodic line solve of one chunk pair held in x[] (as in k_pair)
template <int L>
__device__ __forceinline__ void pair_solve_periodic(dd2 (&x)[L], const double2 *__restrict__ cSP, const double2 *__restrict__ cWF,
                                                    const double2 *__restrict__ cBR, const double *__restrict__ scan, int lane, int nc,
                                                    bool live, double alpha, double ff, int q0, int n) {
  X3D_UNROLL
  for (int m = 1; m < L; ++m) x[m] = fma2(-cSP[m].x, x[m - 1], x[m]);
  dd2 v = x[L - 1];
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    const dd2 o = shfl_up2(v, 1 << lev);
    v = fma2(scan[lev * 32 + lane], o, v);
  }
  dd2 cin = shfl_up2(v, 1);
  if (lane == 0) cin = {0.0, 0.0};
  {
    dd2 xn = {0.0, 0.0};
    X3D_UNROLL
    for (int m = L - 1; m >= 0; --m) {
      // (t(i) - ff t(i+1)) fw(i) as in src/derive.f90:47-50; ff is one number for a periodic operator
      const dd2 tt = fma2(cSP[m].y, cin, x[m]);
      xn = cWF[m].x * fma2(-ff, xn, tt);
      x[m] = xn;
    }
  }
  v = x[0];
  if (!live) v = {0.0, 0.0};
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    const dd2 o = shfl_down2(v, 1 << lev);
    v = fma2(scan[(5 + lev) * 32 + lane], o, v);
  }
  dd2 cb = shfl_down2(v, 1);
  if (lane >= nc - 1) cb = {0.0, 0.0};
  X3D_UNROLL
  for (int m = 0; m < L; ++m) x[m] = fma2(cBR[m].x, cb, x[m]);
  dd2 xl = {0.0, 0.0};
  X3D_UNROLL
  for (int m = 0; m < L; ++m)
    if (q0 + m == n - 1) xl = x[m];
  const dd2 x0 = shfl2(x[0], 0);
  const dd2 xe = shfl2(xl, nc - 1);
  const dd2 sf = {x0.x - alpha * xe.x, x0.y - alpha * xe.y};  // src/derive.f90:55-59
  X3D_UNROLL
  for (int m = 0; m < L; ++m) {
    const double rs = cBR[m].y;
    x[m].x = fma(-sf.x, rs, x[m].x);
    x[m].y = fma(-sf.y, rs, x[m].y);
  }
}

// Cyclic, table-free solve of the periodic compact system  alpha x(i-1) + x(i) + alpha x(i+1) = r(i)  for one chunk
// pair held in x[].  The circulant matrix factors exactly as  c (I - rho S-)(I - rho S+)  (S-, S+: cyclic shifts,
// alpha rho^2 + rho + alpha = 0, |rho| < 1, c = -alpha / rho), so the solve is two first-order recurrences with ONE
// constant multiplier: y(i) = r(i) + rho y(i-1), z(i) = y(i) + rho z(i+1), x = z / c.  Every chunk is the same: no
// boundary rows, no coefficient table, no Sherman-Morrison correction -- the reference's Thomas + Sherman-Morrison
// algorithm (src/derive.f90:45-59) solves the same well-conditioned system (condition (1+2 alpha)/(1-2 alpha) <= 5), so
// the two results agree to a few ulp.  Each sweep runs twice: a Horner pass gives the chunk's zero-carry end value,
// the K previous (next) chunks' end values are combined by shuffles into the exact carry (|rho|^(K L) < 1e-18 makes the
// look-back exact in double precision, and cyclic indexing makes it periodic), then the sweep is repeated with the
// carry.  4 FMA per row and system, no shared-memory traffic.  On return x holds z (the caller applies 1/c).
template <int L>
__device__ __forceinline__ void pair_solve_cyclic(dd2 (&x)[L], const MomGeom::Cyc &cy, int lane, int nc) {
  const double rho = cy.rho;
  // the last chunk holds rem < L real rows followed by zero right-hand sides.  Its padded rows are not masked: what they
  // do is known in closed form (pure decay of the last real value) and is taken out with three per-lane constants.
  const bool last = lane == nc - 1;
  const double escl = last ? cy.esc : 1.0, gl = last ? cy.gamma : 0.0, dl = last ? cy.delta : 0.0;
  // ---- forward: y(i) = r(i) + rho y(i-1)
  dd2 e = x[0];
  X3D_UNROLL
  for (int m = 1; m < L; ++m) e = fma2(rho, e, x[m]);
  e = escl * e;
  dd2 acc = {0.0, 0.0};
#pragma unroll 1
  for (int k = cy.K; k >= 1; --k) {
    int src = lane - k;
    src += src < 0 ? nc : 0;
    const dd2 ev = shfl2(e, src);
    acc = fma2(src == nc - 1 ? cy.rhoR : cy.rhoL, acc, ev);
  }
  dd2 t = acc;
  X3D_UNROLL
  for (int m = 0; m < L; ++m) { t = fma2(rho, t, x[m]); x[m] = t; }
  const dd2 yl = escl * t;   // last chunk: y at its last real row
  // ---- backward: z(i) = y(i) + rho z(i+1)
  dd2 b = x[L - 1];
  X3D_UNROLL
  for (int m = L - 2; m >= 0; --m) b = fma2(rho, b, x[m]);
  b = fma2(-gl, yl, b);
  acc = {0.0, 0.0};
#pragma unroll 1
  for (int k = cy.K; k >= 1; --k) {
    int src = lane + k;
    src -= src >= nc ? nc : 0;
    const dd2 bv = shfl2(b, src);
    acc = fma2(src == nc - 1 ? cy.rhoR : cy.rhoL, acc, bv);
  }
  t = escl * fma2(-dl, yl, acc);
  X3D_UNROLL
  for (int m = L - 1; m >= 0; --m) { t = fma2(rho, t, x[m]); x[m] = t; }
}

// Two independent cyclic solves (different operators, same lines) interleaved statement by statement: every sweep is a
// serial chain of dependent FMAs, and with 8 consumer warps per SM two chains per thread (the two lanes of the pair) do
// not cover the FP64 latency (ncu: fp64 pipe 54 %, stall reason "wait"); four do.  Same arithmetic as two calls of
// pair_solve_cyclic; the look-back runs max(K1, K2) rounds for both (further terms are part of the exact sum).
template <int L>
__device__ __forceinline__ void pair_solve_cyclic2(dd2 (&x1)[L], const MomGeom::Cyc &c1, dd2 (&x2)[L], const MomGeom::Cyc &c2, int lane, int nc) {
  const double r1 = c1.rho, r2 = c2.rho;
  const bool last = lane == nc - 1;
  const double e1s = last ? c1.esc : 1.0, g1 = last ? c1.gamma : 0.0, d1 = last ? c1.delta : 0.0;
  const double e2s = last ? c2.esc : 1.0, g2 = last ? c2.gamma : 0.0, d2 = last ? c2.delta : 0.0;
  const int K = c1.K > c2.K ? c1.K : c2.K;
  dd2 e1 = x1[0], e2 = x2[0];
  X3D_UNROLL
  for (int m = 1; m < L; ++m) { e1 = fma2(r1, e1, x1[m]); e2 = fma2(r2, e2, x2[m]); }
  e1 = e1s * e1; e2 = e2s * e2;
  dd2 a1 = {0.0, 0.0}, a2 = {0.0, 0.0};
#pragma unroll 1
  for (int k = K; k >= 1; --k) {
    int src = lane - k;
    src += src < 0 ? nc : 0;
    const dd2 v1 = shfl2(e1, src), v2 = shfl2(e2, src);
    const bool sl = src == nc - 1;
    a1 = fma2(sl ? c1.rhoR : c1.rhoL, a1, v1);
    a2 = fma2(sl ? c2.rhoR : c2.rhoL, a2, v2);
  }
  dd2 t1 = a1, t2 = a2;
  X3D_UNROLL
  for (int m = 0; m < L; ++m) { t1 = fma2(r1, t1, x1[m]); x1[m] = t1; t2 = fma2(r2, t2, x2[m]); x2[m] = t2; }
  const dd2 y1 = e1s * t1, y2 = e2s * t2;
  dd2 b1 = x1[L - 1], b2 = x2[L - 1];
  X3D_UNROLL
  for (int m = L - 2; m >= 0; --m) { b1 = fma2(r1, b1, x1[m]); b2 = fma2(r2, b2, x2[m]); }
  b1 = fma2(-g1, y1, b1); b2 = fma2(-g2, y2, b2);
  a1 = {0.0, 0.0}; a2 = {0.0, 0.0};
#pragma unroll 1
  for (int k = K; k >= 1; --k) {
    int src = lane + k;
    src -= src >= nc ? nc : 0;
    const dd2 v1 = shfl2(b1, src), v2 = shfl2(b2, src);
    const bool sl = src == nc - 1;
    a1 = fma2(sl ? c1.rhoR : c1.rhoL, a1, v1);
    a2 = fma2(sl ? c2.rhoR : c2.rhoL, a2, v2);
  }
  t1 = e1s * fma2(-d1, y1, a1); t2 = e2s * fma2(-d2, y2, a2);
  X3D_UNROLL
  for (int m = L - 1; m >= 0; --m) { t1 = fma2(r1, t1, x1[m]); x1[m] = t1; t2 = fma2(r2, t2, x2[m]); x2[m] = t2; }
}

// how a thread reaches element j of its window (two lines at once) inside a ring slot
template <bool XD>
struct TileAcc;
template <>
struct TileAcc<false> {  // y / z: [row][16 lanes], 128B swizzle, rows 8 .. 8+n-1 hold the data
  int off[8];
  __device__ __forceinline__ void init(int jw, int q0, int) {
    const int base = q0 + 8 - HALO;
    X3D_UNROLL
    for (int p = 0; p < 8; ++p) off[p] = base * 8 + (jw ^ ((base + p) & 7));
  }
  // one 16-byte access per lane pair (dd2 itself is only 8-byte aligned, which would split it in two)
  __device__ __forceinline__ dd2 ld(const unsigned char *slot, int j) const {
    const double2 t = reinterpret_cast<const double2 *>(slot)[off[j & 7] + 8 * j];
    return {t.x, t.y};
  }
  __device__ __forceinline__ void st(unsigned char *slot, int j, dd2 v) const {
    reinterpret_cast<double2 *>(slot)[off[j & 7] + 8 * j] = make_double2(v.x, v.y);
  }
};
template <>
struct TileAcc<true> {  // x: 16 contiguous lines of pitch n+8 doubles (4 ghosts each side); the pair is two lines
  int o0, o1;
  __device__ __forceinline__ void init(int jw, int q0, int pitch) { o0 = 2 * jw * pitch + q0; o1 = o0 + pitch; }
  __device__ __forceinline__ dd2 ld(const unsigned char *slot, int j) const {
    const double *d = reinterpret_cast<const double *>(slot);
    return {d[o0 + j], d[o1 + j]};
  }
  __device__ __forceinline__ void st(unsigned char *slot, int j, dd2 v) const {
    double *d = reinterpret_cast<double *>(slot);
    d[o0 + j] = v.x;
    d[o1 + j] = v.y;
  }
};

// CYC: table-free cyclic solves (pair_solve_cyclic) instead of the partitioned Thomas tables.
// INTT (x lines only): the time integration is folded into the kernel: instead of storing r_x the consumers stream
// sum (+ extra), the stored right-hand side and u of their two lines, and write u and the new stored right-hand side.
template <int L, int NT2, bool XD, bool CYC, bool INTT>
__global__ void __launch_bounds__(MOM_THREADS, 1)
    k_mom_pair(const __grid_constant__ DevOp op1, const __grid_constant__ DevOp op2, const __grid_constant__ MomMaps maps,
               const MomTabs tb, const __grid_constant__ MomGeom g) {
  static_assert(!INTT || XD, "the time integration is folded into the x kernel only");
  constexpr int NWIN = L + 2 * HALO;
  constexpr int NB = 3;
  constexpr int MOM_TABS = mom_tabs(L), MOM_H = mom_head(L);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int slot_bytes = g.slot_bytes;
  double2 *c1 = reinterpret_cast<double2 *>(smem_raw + NB * slot_bytes);  // [3][MOM_TABS L]
  double2 *c2 = c1 + 3 * MOM_TABS * L;
  double *scan1 = reinterpret_cast<double *>(c2 + 3 * MOM_TABS * L);     // [10][32]
  double *scan2 = scan1 + 320;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(scan2 + 320);
  unsigned long long *done = full + NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = g.n, nc = g.nc;
  if constexpr (CYC) {   // no tables; the area stays as padding behind the last slot (the last chunk's window overruns into it)
    for (int idx = threadIdx.x; idx < 3 * MOM_TABS * L; idx += blockDim.x) { c1[idx] = make_double2(0.0, 0.0); c2[idx] = make_double2(0.0, 0.0); }
  } else {
    for (int idx = threadIdx.x; idx < 3 * MOM_TABS * L; idx += blockDim.x) { c1[idx] = tb.c1[idx]; c2[idx] = tb.c2[idx]; }
    for (int idx = threadIdx.x; idx < 320; idx += blockDim.x) { scan1[idx] = tb.scan1[idx]; scan2[idx] = tb.scan2[idx]; }
  }
  if (threadIdx.x == 0) {
    X3D_UNROLL
    for (int b = 0; b < NB; ++b) { mbar_init(full + b, 1); mbar_init(done + b, PAIR_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < g.npos ? (g.npos - first + step - 1) / step : 0;  // positions of this CTA
  // tile q of position p (q = 0: c1, 1: c2, 2: a, the order of consumption) lives in slot (p + 1 + q) % 3
  const int fld[3] = {g.ic1, g.ic2, g.ia};

  if (warp >= PAIR_WARPS) {
    // ---------------- TMA producer ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp != PAIR_WARPS || lane != 0) return;
    auto load = [&](long long p, int q) {
      const long long pos = first + p * step;
      const int slot = static_cast<int>((p + 1 + q) % 3);
      unsigned char *dst = smem_raw + slot * slot_bytes;
      const int seg = static_cast<int>(pos % g.nseg);
      const long long blk = pos / g.nseg;
      int row0 = seg * g.seg_S - g.seg_H;           // first row of the tile in the line (wraps)
      row0 += row0 < 0 ? g.ntot : 0;
      if constexpr (XD) {
        const long long line0 = blk * 16;
        const int nl = static_cast<int>(g.nlines - line0 < 16 ? g.nlines - line0 : 16);
        const double *src = g.fin[fld[q]] + line0 * g.ntot;
        mbar_expect_tx(full + slot, static_cast<unsigned>(nl) * n * 8u);
        const int n1 = g.ntot - row0 < n ? g.ntot - row0 : n;   // rows up to the end of the line, then from its start
        for (int l = 0; l < nl; ++l) {
          const double *ls = src + static_cast<long long>(l) * g.ntot;
          bulk_g2s(dst + (l * g.pitch + HALO) * 8, ls + row0, static_cast<unsigned>(n1) * 8u, full + slot);
          if (n1 < n) bulk_g2s(dst + (l * g.pitch + HALO + n1) * 8, ls, static_cast<unsigned>(n - n1) * 8u, full + slot);
        }
      } else {
        const int bx = static_cast<int>(blk % g.nbx), by = static_cast<int>(blk / g.nbx);
        const CUtensorMap *tm = &maps.in[fld[q]], *th = &maps.halo[fld[q]];
        mbar_expect_tx(full + slot, (static_cast<unsigned>(g.nbox) * g.br + 16u) * 128u);
        for (int b = 0; b < g.nbox; ++b) {            // boxes never straddle the end of the line (ntot, row0 are multiples of br)
          int r = row0 + b * g.br;
          r -= r >= g.ntot ? g.ntot : 0;
          tma_load_3d(dst + (8 + b * g.br) * 128, tm, bx * 16, r, by, full + slot);
        }
        int rp = row0 - 8, rn = row0 + n;             // the 8 rows before and after the tile: the solver's own wrap ghosts
        rp += rp < 0 ? g.ntot : 0;
        rn -= rn >= g.ntot ? g.ntot : 0;
        tma_load_3d(dst, th, bx * 16, rp, by, full + slot);
        tma_load_3d(dst + (8 + n) * 128, th, bx * 16, rn, by, full + slot);
      }
    };
    if (mine > 0) { load(0, 2); load(0, 0); load(0, 1); }
    for (long long p = 0; p < mine; ++p) {
      const long long pos = first + p * step;
      for (int q = 0; q < 3; ++q) {
        const int slot = static_cast<int>((p + 1 + q) % 3);
        mbar_wait(done + slot, static_cast<unsigned>(p & 1));
        const unsigned char *src = smem_raw + slot * slot_bytes;
        if constexpr (INTT) {
          // the consumers have written their results to global memory themselves: the slot is free
          (void)src;
        } else if constexpr (XD) {
          const int seg = static_cast<int>(pos % g.nseg);
          const long long line0 = (pos / g.nseg) * 16;
          const int nl = static_cast<int>(g.nlines - line0 < 16 ? g.nlines - line0 : 16);
          double *dst = g.fout[fld[q]] + line0 * g.ntot + seg * g.seg_S;     // the inner seg_S rows of the tile
          const unsigned bytes = static_cast<unsigned>(g.seg_S) * 8u;
          if (g.add == 0) {
            for (int l = 0; l < nl; ++l)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + static_cast<long long>(l) * g.ntot),
                           "r"(smem_u32(src + (l * g.pitch + HALO + g.seg_H) * 8)), "r"(bytes)
                           : "memory");
          } else {
            for (int l = 0; l < nl; ++l)
              asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst + static_cast<long long>(l) * g.ntot),
                           "r"(smem_u32(src + (l * g.pitch + HALO + g.seg_H) * 8)), "r"(bytes)
                           : "memory");
          }
        } else {
          const int seg = static_cast<int>(pos % g.nseg);
          const long long blk = pos / g.nseg;
          const int bx = static_cast<int>(blk % g.nbx), by = static_cast<int>(blk / g.nbx);
          const CUtensorMap *tm = &maps.out[fld[q]];
          const int nbs = g.seg_S / g.br;              // the inner seg_S rows of the tile
          const unsigned char *s0 = src + (8 + g.seg_H) * 128;
          if (g.add == 0) for (int b = 0; b < nbs; ++b) tma_store_3d(tm, bx * 16, seg * g.seg_S + b * g.br, by, s0 + b * g.br * 128);
          else for (int b = 0; b < nbs; ++b) tma_red_add_3d(tm, bx * 16, seg * g.seg_S + b * g.br, by, s0 + b * g.br * 128);
        }
        bulk_commit();
        if (p + 1 < mine) {
          // the slot is free once the store has left shared memory; for the next position it takes
          // c1's slot -> a, c2's slot -> c1, a's slot -> c2 (one to two component times ahead of its use)
          bulk_wait_read<0>();
          load(p + 1, (q + 2) % 3);
        }
      }
    }
    bulk_wait_read<0>();
    return;
  }

  // ---------------- consumers ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
  const int jw = warp;
  const int cl = lane < nc ? lane : nc - 1;
  const bool live = lane < nc;
  const int q0 = cl * L;
  TileAcc<XD> acc;
  acc.init(jw, q0, g.pitch);
  const int tab = cl < MOM_H ? cl : (cl >= nc - MOM_H - 1 ? MOM_H + 1 + cl - (nc - MOM_H - 1) : MOM_H);
  const double2 *s1 = c1 + tab * L, *w1 = c1 + MOM_TABS * L + tab * L, *b1 = c1 + 2 * MOM_TABS * L + tab * L;
  const double2 *s2 = c2 + tab * L, *w2 = c2 + MOM_TABS * L + tab * L, *b2 = c2 + 2 * MOM_TABS * L + tab * L;
  const double xnu = g.xnu;
  // x lines: the wrap ghosts of a freshly loaded tile are copied inside shared memory, each warp for its two lines
  auto ghosts = [&](unsigned char *slot) {
    if constexpr (XD) {
      if (lane < 8) {
        double *d = reinterpret_cast<double *>(slot) + (2 * jw + (lane >> 2)) * g.pitch + HALO;
        const int gq = lane & 3;
        d[-1 - gq] = d[n - 1 - gq];
        d[n + gq] = d[gq];
      }
      __syncwarp();
    }
  };
  for (long long p = 0; p < mine; ++p) {
    const unsigned par = static_cast<unsigned>(p & 1);
    const int slot_a = static_cast<int>((p + 3) % 3);
    unsigned char *bufA = smem_raw + slot_a * slot_bytes;
    mbar_wait(full + slot_a, par);
    ghosts(bufA);
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
      const int slot = static_cast<int>((p + 1 + q) % 3);
      unsigned char *bufC = smem_raw + slot * slot_bytes;
      if (q < 2) { mbar_wait(full + slot, par); ghosts(bufC); }
      dd2 r[L];
      {  // xnu * D2(c) and - 1/2 a D1(c): both right-hand sides from one read of the window of c
        dd2 x[L];
        {
          dd2 win[NWIN];
          X3D_UNROLL
          for (int j = 0; j < NWIN; ++j) win[j] = acc.ld(bufC, j);
          X3D_UNROLL
          for (int m = 0; m < L; ++m) {
            const dd2 v2 = rhs_interior<D2, NT2, NWIN, dd2>(op2, win, m);
            const dd2 v1 = rhs_interior<D1, 2, NWIN, dd2>(op1, win, m);
            const bool ok = live && q0 + m < n;
            r[m].x = ok ? v2.x : 0.0;
            r[m].y = ok ? v2.y : 0.0;
            x[m].x = ok ? v1.x : 0.0;
            x[m].y = ok ? v1.y : 0.0;
          }
        }
        if constexpr (CYC) {
          pair_solve_cyclic2<L>(r, g.cy2, x, g.cy1, lane, nc);
          X3D_UNROLL
          for (int m = 0; m < L; ++m) r[m] = g.cy2.scale * r[m];
          const double k1 = g.cy1.scale;
          X3D_UNROLL
          for (int m = 0; m < L; ++m) {
            const dd2 a = acc.ld(bufA, m + HALO);
            r[m].x = fma(k1 * a.x, x[m].x, r[m].x);
            r[m].y = fma(k1 * a.y, x[m].y, r[m].y);
          }
        } else {
          pair_solve_periodic<L>(r, s2, w2, b2, scan2, lane, nc, live, op2.alpha, tb.ff2, q0, n);
          X3D_UNROLL
          for (int m = 0; m < L; ++m) r[m] = xnu * r[m];
          pair_solve_periodic<L>(x, s1, w1, b1, scan1, lane, nc, live, op1.alpha, tb.ff1, q0, n);
          X3D_UNROLL
          for (int m = 0; m < L; ++m) {
            const dd2 a = acc.ld(bufA, m + HALO);
            r[m].x = fma(-0.5 * a.x, x[m].x, r[m].x);
            r[m].y = fma(-0.5 * a.y, x[m].y, r[m].y);
          }
        }
      }
      __syncwarp();
      {  // - 1/2 D1(c a)
        dd2 x[L];
        {
          dd2 win[NWIN];
          X3D_UNROLL
          for (int j = 0; j < NWIN; ++j) {
            const dd2 cc = acc.ld(bufC, j);
            const dd2 aa = acc.ld(bufA, j);
            win[j] = {cc.x * aa.x, cc.y * aa.y};
          }
          X3D_UNROLL
          for (int m = 0; m < L; ++m) {
            const dd2 v = rhs_interior<D1, 2, NWIN, dd2>(op1, win, m);
            const bool ok = live && q0 + m < n;
            x[m].x = ok ? v.x : 0.0;
            x[m].y = ok ? v.y : 0.0;
          }
        }
        if constexpr (CYC) {
          pair_solve_cyclic<L>(x, g.cy1, lane, nc);
          const double k1 = g.cy1.scale;
          X3D_UNROLL
          for (int m = 0; m < L; ++m) {
            r[m].x = fma(k1, x[m].x, r[m].x);
            r[m].y = fma(k1, x[m].y, r[m].y);
          }
        } else {
          pair_solve_periodic<L>(x, s1, w1, b1, scan1, lane, nc, live, op1.alpha, tb.ff1, q0, n);
          X3D_UNROLL
          for (int m = 0; m < L; ++m) {
            r[m].x = fma(-0.5, x[m].x, r[m].x);
            r[m].y = fma(-0.5, x[m].y, r[m].y);
          }
        }
      }
      __syncwarp();  // every lane has read its windows of c (and of a when c == a)
      if (live) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m < n) acc.st(bufC, m + HALO, r[m]);
      }
      if constexpr (INTT) {
        // time integration of the warp's two lines, streamed with 16-byte coalesced accesses:
        //   N = sum (+ extra) + r ;  u <- ca N + cb old + u ;  old <- N
        __syncwarp();
        const int f = fld[q];
        const long long posq = first + p * step;
        const int seg = static_cast<int>(posq % g.nseg);
        const long long line0 = (posq / g.nseg) * 16 + 2 * jw;
        const double ca = g.ca, cb = g.cb;
        // all loads of a batch (8 x 32 lanes x 16 bytes per array) are issued before the first use: the warp has
        // nothing else to overlap the HBM latency with, so it needs the bytes in flight (the solve's registers are dead here)
        constexpr int NBT = 8;
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {
          if (line0 + l >= g.nlines) break;
          const double2 *rs = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(bufC) + (2 * jw + l) * g.pitch + HALO + g.seg_H);
          const long long gb = ((line0 + l) * static_cast<long long>(g.ntot) + seg * g.seg_S) / 2;   // the inner seg_S rows
          const double2 *gs = reinterpret_cast<const double2 *>(g.isum[f]) + gb;
          const double2 *ge = reinterpret_cast<const double2 *>(g.iextra[f]) + gb;
          const double2 *go = reinterpret_cast<const double2 *>(g.iold_in[f]) + gb;
          const double2 *gu = reinterpret_cast<const double2 *>(g.iu[f]) + gb;
          double2 *gw = reinterpret_cast<double2 *>(g.iu_out[f]) + gb;
          double2 *gn = reinterpret_cast<double2 *>(g.iold_out[f]) + gb;
          const int nh = g.seg_S / 2;
#pragma unroll 1
          for (int i0 = 0; i0 < nh; i0 += 32 * NBT) {
            double2 Sv[NBT], Uv[NBT], Ov[NBT];
            X3D_UNROLL
            for (int t = 0; t < NBT; ++t) {
              const int i = i0 + lane + 32 * t;
              Sv[t] = i < nh ? __ldcs(gs + i) : make_double2(0.0, 0.0);
            }
            X3D_UNROLL
            for (int t = 0; t < NBT; ++t) {
              const int i = i0 + lane + 32 * t;
              Uv[t] = i < nh ? gu[i] : make_double2(0.0, 0.0);
            }
            if (g.use_old) {
              X3D_UNROLL
              for (int t = 0; t < NBT; ++t) {
                const int i = i0 + lane + 32 * t;
                Ov[t] = i < nh ? __ldcs(go + i) : make_double2(0.0, 0.0);
              }
            } else {
              X3D_UNROLL
              for (int t = 0; t < NBT; ++t) Ov[t] = make_double2(0.0, 0.0);
            }
            if (g.has_extra) {
              X3D_UNROLL
              for (int t = 0; t < NBT; ++t) {
                const int i = i0 + lane + 32 * t;
                if (i < nh) { const double2 E = __ldcs(ge + i); Sv[t].x = E.x + Sv[t].x; Sv[t].y = E.y + Sv[t].y; }
              }
            }
            X3D_UNROLL
            for (int t = 0; t < NBT; ++t) {
              const int i = i0 + lane + 32 * t;
              if (i < nh) {
                const double2 rr = rs[i];
                double2 N = Sv[t];
                if (g.has_extra) { N.x = N.x + rr.x; N.y = N.y + rr.y; }   // extra + (sum + r) as intt3_fused
                else { N.x += rr.x; N.y += rr.y; }
                double2 U = Uv[t];
                U.x = ca * N.x + cb * Ov[t].x + U.x;
                U.y = ca * N.y + cb * Ov[t].y + U.y;
                gw[i] = U;
                if (g.store_old) __stcs(gn + i, N);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(done + slot);
      } else {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(done + slot);
      }
    }
  }
}

}  // namespace x3d
