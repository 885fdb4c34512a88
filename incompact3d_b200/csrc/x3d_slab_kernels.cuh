// x3d_slab_kernels.cuh -- the z part of the fused momentum right-hand side WITHOUT pencil transposes (sm_100a).
//
// With slabs (p_row = 1) every rank holds nzl = nz / P consecutive z planes of the velocity in its x / y pencils.  The
// reference transposes the three components to z pencils, differentiates along z and transposes the three results
// back (src/transeq.f90:236-320): 6 of the 12 field transposes of a sub-step.  Here each rank solves its OWN nzl rows
// of every z line and the ranks exchange a few planes instead of whole fields.
//
// The periodic compact system tri(alpha, 1, alpha) x = r factors as c (I - rho S-)(I - rho S+) (see pair_solve_cyclic),
// i.e. y(i) = r(i) + rho y(i-1), z(i) = y(i) + rho z(i+1), x = z / c, and both recurrences forget geometrically
// (|rho| = 0.38 for the first, 0.19 for the second derivative).  For the rows i = 0 .. n-1 of a slab
//     z(i) = z0(i) + A(i) Yin + B(i) Zin ,   A(i) = rho^(i+1) (1 - rho^(2(n-i))) / (1 - rho^2) ,  B(i) = rho^(n-i)
// where z0 is the solve with zero carries, Yin = y(-1) is the forward value at the last row of the previous slab and
// Zin = z(n) the backward value at the first row of the next slab.  With |rho|^n < 1e-17 (n >= 64 rows):
//     Yin(g) = Yout(g-1) ,  Zin(g) = Z0(g+1) + A(0) Yout(g)
// with Yout = the zero-carry forward value at a slab's last row and Z0 = its zero-carry backward value at row 0.  So:
//   1. halo exchange: 4 planes of each velocity component from both neighbours (the right-hand-side stencils);
//   2. k_mom_slab: the nine zero-carry solves of the slab, result stored; Yout and Z0 of every line and system written
//      to carry planes (18 planes);
//   3. carry exchange: Yout to the next rank, Z0 to the previous one;
//   4. k_zfix: the rank-one corrections A Yin + B Zin on the rows next to the slab faces (where they are above
//      double precision), added to the stored result.
// 42 planes cross NVLink per sub-step instead of 6 (P-1)/P fields, and the result is already in the x / y pencils.
//
// k_mom_slab is k_mom_pair<L = 9, XD = false, CYC = true> with: ghost rows from halo arrays instead of the line's own
// wrap, look-backs that stop at the slab faces, carries emitted, and -- because a slab line has few chunks (nc = 8 for 64
// rows) -- G = 32 / W lines per warp (W = 8, 16, 32 lanes per line), a ring slot holding G tiles of 16 lanes.
#pragma once
#include "x3d_mom_kernels.cuh"

namespace x3d {

struct SlabGeom {
  int nbx;               // 16-lane tiles across the plane
  long long npos;        // tile groups (G tiles each)
  long long ntiles;
  int gshift;            // G = 1 << gshift tiles per ring slot, W = 32 >> gshift lanes (chunks) per line
  int sub_bytes, slot_bytes;
  int nbox, br;
  int n, nc, rem;
  int ia, ic1, ic2;
  MomGeom::Cyc cy1, cy2;
  int add;
  double *carry;         // [18][nlanes]: Yout of the 9 systems (3 components in the order c1, c2, a; D2(c), D1(c), D1(c a)), then Z0
  long long nlanes;
};

// zero-carry solves of one chunk pair of a slab line (pair_solve_cyclic2 without the wrap); ln = chunk index, lbase = first lane of the
// line's lane group; yo = y at the last row of the slab (valid in the last chunk)
template <int L>
__device__ __forceinline__ void pair_solve_open2(dd2 (&x1)[L], const MomGeom::Cyc &c1, dd2 (&x2)[L], const MomGeom::Cyc &c2, int ln, int lbase,
                                                 int nc, dd2 &yo1, dd2 &yo2) {
  const double r1 = c1.rho, r2 = c2.rho;
  const bool last = ln == nc - 1;
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
    const int src = ln - k;
    const bool ok = src >= 0;
    dd2 v1 = shfl2(e1, lbase + (ok ? src : 0)), v2 = shfl2(e2, lbase + (ok ? src : 0));
    if (!ok) { v1 = {0.0, 0.0}; v2 = {0.0, 0.0}; }
    a1 = fma2(c1.rhoL, a1, v1);
    a2 = fma2(c2.rhoL, a2, v2);
  }
  dd2 t1 = a1, t2 = a2;
  X3D_UNROLL
  for (int m = 0; m < L; ++m) { t1 = fma2(r1, t1, x1[m]); x1[m] = t1; t2 = fma2(r2, t2, x2[m]); x2[m] = t2; }
  const dd2 y1 = e1s * t1, y2 = e2s * t2;
  yo1 = y1; yo2 = y2;
  dd2 b1 = x1[L - 1], b2 = x2[L - 1];
  X3D_UNROLL
  for (int m = L - 2; m >= 0; --m) { b1 = fma2(r1, b1, x1[m]); b2 = fma2(r2, b2, x2[m]); }
  b1 = fma2(-g1, y1, b1); b2 = fma2(-g2, y2, b2);
  a1 = {0.0, 0.0}; a2 = {0.0, 0.0};
#pragma unroll 1
  for (int k = K; k >= 1; --k) {
    const int src = ln + k;
    const bool ok = src <= nc - 1;
    dd2 v1 = shfl2(b1, lbase + (ok ? src : 0)), v2 = shfl2(b2, lbase + (ok ? src : 0));
    if (!ok) { v1 = {0.0, 0.0}; v2 = {0.0, 0.0}; }
    const bool sl = src == nc - 1;
    a1 = fma2(sl ? c1.rhoR : c1.rhoL, a1, v1);
    a2 = fma2(sl ? c2.rhoR : c2.rhoL, a2, v2);
  }
  t1 = e1s * fma2(-d1, y1, a1); t2 = e2s * fma2(-d2, y2, a2);
  X3D_UNROLL
  for (int m = L - 1; m >= 0; --m) { t1 = fma2(r1, t1, x1[m]); x1[m] = t1; t2 = fma2(r2, t2, x2[m]); x2[m] = t2; }
}

template <int L>
__device__ __forceinline__ void pair_solve_open(dd2 (&x)[L], const MomGeom::Cyc &cy, int ln, int lbase, int nc, dd2 &yo) {
  const double rho = cy.rho;
  const bool last = ln == nc - 1;
  const double escl = last ? cy.esc : 1.0, gl = last ? cy.gamma : 0.0, dl = last ? cy.delta : 0.0;
  dd2 e = x[0];
  X3D_UNROLL
  for (int m = 1; m < L; ++m) e = fma2(rho, e, x[m]);
  e = escl * e;
  dd2 acc = {0.0, 0.0};
#pragma unroll 1
  for (int k = cy.K; k >= 1; --k) {
    const int src = ln - k;
    const bool ok = src >= 0;
    dd2 ev = shfl2(e, lbase + (ok ? src : 0));
    if (!ok) ev = {0.0, 0.0};
    acc = fma2(cy.rhoL, acc, ev);
  }
  dd2 t = acc;
  X3D_UNROLL
  for (int m = 0; m < L; ++m) { t = fma2(rho, t, x[m]); x[m] = t; }
  const dd2 yl = escl * t;
  yo = yl;
  dd2 b = x[L - 1];
  X3D_UNROLL
  for (int m = L - 2; m >= 0; --m) b = fma2(rho, b, x[m]);
  b = fma2(-gl, yl, b);
  acc = {0.0, 0.0};
#pragma unroll 1
  for (int k = cy.K; k >= 1; --k) {
    const int src = ln + k;
    const bool ok = src <= nc - 1;
    dd2 bv = shfl2(b, lbase + (ok ? src : 0));
    if (!ok) bv = {0.0, 0.0};
    acc = fma2(src == nc - 1 ? cy.rhoR : cy.rhoL, acc, bv);
  }
  t = escl * fma2(-dl, yl, acc);
  X3D_UNROLL
  for (int m = L - 1; m >= 0; --m) { t = fma2(rho, t, x[m]); x[m] = t; }
}

// maps.in / maps.out: the slab's fields as (lanes, n rows, 1); maps.halo: the halo arrays (lanes, 16 rows, 1), rows 4..7 =
// the 4 planes below the slab, rows 8..11 = the 4 planes above it
template <int L, int NT2>
__global__ void __launch_bounds__(MOM_THREADS, 1)
    k_mom_slab(const __grid_constant__ DevOp op1, const __grid_constant__ DevOp op2, const __grid_constant__ MomMaps maps,
               const __grid_constant__ SlabGeom g) {
  constexpr int NWIN = L + 2 * HALO;
  constexpr int NB = 3;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int slot_bytes = g.slot_bytes, sub_bytes = g.sub_bytes;
  double *pad = reinterpret_cast<double *>(smem_raw + NB * slot_bytes);   // zeroed: the last chunk's window overruns the last tile
  unsigned long long *full = reinterpret_cast<unsigned long long *>(pad + 512);
  unsigned long long *done = full + NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = g.n, nc = g.nc;
  const int G = 1 << g.gshift;
  for (int idx = threadIdx.x; idx < 512; idx += blockDim.x) pad[idx] = 0.0;
  if (threadIdx.x == 0) {
    X3D_UNROLL
    for (int b = 0; b < NB; ++b) { mbar_init(full + b, 1); mbar_init(done + b, PAIR_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < g.npos ? (g.npos - first + step - 1) / step : 0;
  const int fld[3] = {g.ic1, g.ic2, g.ia};

  if (warp >= PAIR_WARPS) {
    // ---------------- TMA producer ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp != PAIR_WARPS || lane != 0) return;
    auto load = [&](long long p, int q) {
      const long long pos = first + p * step;
      const int slot = static_cast<int>((p + 1 + q) % 3);
      const CUtensorMap *tm = &maps.in[fld[q]], *th = &maps.halo[fld[q]];
      mbar_expect_tx(full + slot, static_cast<unsigned>(G) * (static_cast<unsigned>(g.nbox) * g.br + 16u) * 128u);
      for (int s = 0; s < G; ++s) {
        const long long tile = pos * G + s;
        const int bx = static_cast<int>(tile % g.nbx), by = static_cast<int>(tile / g.nbx);
        unsigned char *dst = smem_raw + slot * slot_bytes + s * sub_bytes;
        for (int b = 0; b < g.nbox; ++b) tma_load_3d(dst + (8 + b * g.br) * 128, tm, bx * 16, b * g.br, by, full + slot);
        tma_load_3d(dst, th, bx * 16, 0, by, full + slot);
        tma_load_3d(dst + (8 + n) * 128, th, bx * 16, 8, by, full + slot);
      }
    };
    if (mine > 0) { load(0, 2); load(0, 0); load(0, 1); }
    for (long long p = 0; p < mine; ++p) {
      const long long pos = first + p * step;
      for (int q = 0; q < 3; ++q) {
        const int slot = static_cast<int>((p + 1 + q) % 3);
        mbar_wait(done + slot, static_cast<unsigned>(p & 1));
        const CUtensorMap *tm = &maps.out[fld[q]];
        for (int s = 0; s < G; ++s) {
          const long long tile = pos * G + s;
          const int bx = static_cast<int>(tile % g.nbx), by = static_cast<int>(tile / g.nbx);
          const unsigned char *s0 = smem_raw + slot * slot_bytes + s * sub_bytes + 8 * 128;
          if (g.add == 0) for (int b = 0; b < g.nbox; ++b) tma_store_3d(tm, bx * 16, b * g.br, by, s0 + b * g.br * 128);
          else for (int b = 0; b < g.nbox; ++b) tma_red_add_3d(tm, bx * 16, b * g.br, by, s0 + b * g.br * 128);
        }
        bulk_commit();
        if (p + 1 < mine) {
          bulk_wait_read<0>();
          load(p + 1, (q + 2) % 3);
        }
      }
    }
    bulk_wait_read<0>();
    return;
  }

  // ---------------- consumers: warp = G lane pairs (one per tile of the slot), W lanes = chunks per line ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
  const int jw = warp;
  const int W = 32 >> g.gshift;
  const int ln = lane & (W - 1), lbase = lane - ln, grp = lane >> (5 - g.gshift);
  const int cl = ln < nc ? ln : nc - 1;
  const bool live = ln < nc;
  const int q0 = cl * L;
  TileAcc<false> acc;
  acc.init(jw, q0, 0);
  const int sub = grp * sub_bytes;
  for (long long p = 0; p < mine; ++p) {
    const unsigned par = static_cast<unsigned>(p & 1);
    const long long tile = (first + p * step) * G + grp;
    const long long gl0 = tile * 16 + 2 * jw;          // first of this thread's two lines in the plane
    const bool emit = gl0 + 1 < g.nlanes;
    const int slot_a = static_cast<int>((p + 3) % 3);
    unsigned char *bufA = smem_raw + slot_a * slot_bytes + sub;
    mbar_wait(full + slot_a, par);
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {
      const int slot = static_cast<int>((p + 1 + q) % 3);
      unsigned char *bufC = smem_raw + slot * slot_bytes + sub;
      if (q < 2) mbar_wait(full + slot, par);
      double *cq = g.carry + static_cast<long long>(3 * q) * g.nlanes + gl0;
      const long long zoff = 9 * g.nlanes;
      dd2 r[L];
      {
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
        dd2 yo2, yo1;
        pair_solve_open2<L>(r, g.cy2, x, g.cy1, ln, lbase, nc, yo2, yo1);
        if (emit) {
          if (ln == nc - 1) {
            *reinterpret_cast<double2 *>(cq) = make_double2(yo2.x, yo2.y);
            *reinterpret_cast<double2 *>(cq + g.nlanes) = make_double2(yo1.x, yo1.y);
          }
          if (ln == 0) {
            *reinterpret_cast<double2 *>(cq + zoff) = make_double2(r[0].x, r[0].y);
            *reinterpret_cast<double2 *>(cq + zoff + g.nlanes) = make_double2(x[0].x, x[0].y);
          }
        }
        X3D_UNROLL
        for (int m = 0; m < L; ++m) r[m] = g.cy2.scale * r[m];
        const double k1 = g.cy1.scale;
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          const dd2 a = acc.ld(bufA, m + HALO);
          r[m].x = fma(k1 * a.x, x[m].x, r[m].x);
          r[m].y = fma(k1 * a.y, x[m].y, r[m].y);
        }
      }
      __syncwarp();
      {
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
        dd2 yo;
        pair_solve_open<L>(x, g.cy1, ln, lbase, nc, yo);
        if (emit) {
          if (ln == nc - 1) *reinterpret_cast<double2 *>(cq + 2 * g.nlanes) = make_double2(yo.x, yo.y);
          if (ln == 0) *reinterpret_cast<double2 *>(cq + zoff + 2 * g.nlanes) = make_double2(x[0].x, x[0].y);
        }
        const double k1 = g.cy1.scale;
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          r[m].x = fma(k1, x[m].x, r[m].x);
          r[m].y = fma(k1, x[m].y, r[m].y);
        }
      }
      __syncwarp();  // every lane has read its windows of c (and of a when c == a)
      if (live) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m < n) acc.st(bufC, m + HALO, r[m]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(done + slot);
    }
  }
}

// corrections of the rows next to the slab faces (step 4 above).  One thread per line; rows 0 .. W-1 and n-W .. n-1.
struct ZFixArgs {
  const double *yin, *z0n, *yout;   // [9][nl]: Yout of the previous slab, Z0 of the next slab, this slab's own Yout
  double *sum[3];                   // results of the components in the order of the carries (c1, c2, a)
  const double *a;                  // advecting velocity (the z component) of the slab
  const double *tab;                // [4][n]: A (D1), B (D1), A (D2), B (D2)
  long long nl;
  int n, W;
  double k1, k2;                    // -1/2 / c (D1), xnu / c (D2)
};
__global__ void __launch_bounds__(256) k_zfix(const __grid_constant__ ZFixArgs z) {
  // two neighbouring lines per thread (16-byte accesses), four rows in flight
  const long long l = 2 * (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x);
  if (l >= z.nl) return;
  double2 Y[9], Z[9];
  const double A01 = z.tab[0], A02 = z.tab[2 * z.n];
  X3D_UNROLL
  for (int s = 0; s < 9; ++s) {
    Y[s] = *reinterpret_cast<const double2 *>(z.yin + s * z.nl + l);
    const double2 yo = *reinterpret_cast<const double2 *>(z.yout + s * z.nl + l), z0 = *reinterpret_cast<const double2 *>(z.z0n + s * z.nl + l);
    const double a0 = (s % 3 == 0) ? A02 : A01;
    Z[s] = make_double2(fma(a0, yo.x, z0.x), fma(a0, yo.y, z0.y));
  }
  const int n = z.n, W = z.W;
  const int nrows = (n - W > W) ? 2 * W : n;          // rows 0 .. W-1 and n-W .. n-1 (all rows when the bands meet)
  constexpr int NB = 4;
  for (int r0 = 0; r0 < nrows; r0 += NB) {
    double2 av[NB], sv[NB][3];
    int row[NB];
    X3D_UNROLL
    for (int q = 0; q < NB; ++q) {
      const int r = r0 + q;
      row[q] = (nrows == n || r < W) ? r : n - 2 * W + r;
      if (r < nrows) {
        const long long o = static_cast<long long>(row[q]) * z.nl + l;
        av[q] = *reinterpret_cast<const double2 *>(z.a + o);
        X3D_UNROLL
        for (int c = 0; c < 3; ++c) sv[q][c] = *reinterpret_cast<const double2 *>(z.sum[c] + o);
      }
    }
    X3D_UNROLL
    for (int q = 0; q < NB; ++q) {
      if (r0 + q >= nrows) break;
      const int i = row[q];
      const double a1 = z.tab[i], b1 = z.tab[n + i], a2 = z.tab[2 * n + i], b2 = z.tab[3 * n + i];
      const long long o = static_cast<long long>(i) * z.nl + l;
      X3D_UNROLL
      for (int c = 0; c < 3; ++c) {
        const double d0x = fma(a2, Y[3 * c].x, b2 * Z[3 * c].x), d0y = fma(a2, Y[3 * c].y, b2 * Z[3 * c].y);
        const double d1x = fma(a1, Y[3 * c + 1].x, b1 * Z[3 * c + 1].x), d1y = fma(a1, Y[3 * c + 1].y, b1 * Z[3 * c + 1].y);
        const double d2x = fma(a1, Y[3 * c + 2].x, b1 * Z[3 * c + 2].x), d2y = fma(a1, Y[3 * c + 2].y, b1 * Z[3 * c + 2].y);
        double2 sN = sv[q][c];
        sN.x += z.k2 * d0x + z.k1 * fma(av[q].x, d1x, d2x);
        sN.y += z.k2 * d0y + z.k1 * fma(av[q].y, d1y, d2y);
        *reinterpret_cast<double2 *>(z.sum[c] + o) = sN;
      }
    }
  }
}

}  // namespace x3d
