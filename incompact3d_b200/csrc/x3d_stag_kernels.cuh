// x3d_stag_kernels.cuh -- fused pairs of staggered operators on periodic y / z lines (sm_100a).
//
// divergence (src/navier.f90:321-339) and gradp (:404-417) call, per direction, two staggered operators that either
// share their input or are summed:
//   MODE 0  (two inputs, one output)   t  = opA(uA) + opB(uB)      duydypi2 = deryvp(pgy1) + interyvp(pp1)   :321-325
//                                                                  pp3      = interzvp(duydypi3) + derzvp(upi3) :333-339
//   MODE 1  (one input, two outputs)   tA = opA(u), tB = opB(u)    ppi3, pgz3 = interzpv(pp3), derzpv(pp3)   :404-406
//                                                                  ppi2, pgy2 = interypv(pp2), derypv(pp2)   :413-415
// One kernel per pair moves 3 arrays instead of 5 (MODE 0: the separate sum pass and the intermediate disappear) or 4
// (MODE 1: the shared input is read once).
//
// Structure of k_pair: 128B-swizzled tensor-map TMA tiles of 16 lanes x whole line in a 3-slot ring, warp = lane pair,
// thread = chunk of L rows, a TMA producer warp, mbarriers only.  The ring runs over a sequence of tile uses --
// MODE 0: B (load, read, free), A (load, solved in place, stored); MODE 1: A (load, opA in place, stored), S (no load: the
// slot receives opB's result and is stored).
//
// The periodic systems are solved without coefficient tables (pair_solve_cyclic_ks): the circulant matrix
// tri(alpha, 1, alpha) factors as c (I - rho S-)(I - rho S+), two first-order recurrences with one constant
// multiplier; the recurrence over the chunk ends is a cyclic Kogge-Stone scan with per-lane multipliers
// rho^(rows spanned), which also covers the slowly decaying interpolators (alpha = 0.49: rho = -0.817).
#pragma once
#include "x3d_mom_kernels.cuh"

namespace x3d {

struct StagCyc {
  double rho, esc, gamma, delta, scale;   // as MomGeom::Cyc; scale = 1 / c
  const double *scan;                     // device [10][32]: forward levels 0-4, backward levels 0-4 (per-lane multipliers)
};
struct StagGeom {
  int nbx;
  long long npos;
  int slot_bytes;
  int nbox, br;
  int n, nc;
  StagCyc a, b;
};
struct StagMaps {
  CUtensorMap inA, haloA, inB, haloB, outA, outB;
};

// cyclic solve of one chunk pair; x = right-hand side on entry, z (= c * solution) on return
template <int L>
__device__ __forceinline__ void pair_solve_cyclic_ks(dd2 (&x)[L], const StagCyc &cy, int lane, int nc) {
  const double rho = cy.rho;
  const bool last = lane == nc - 1;
  const double escl = last ? cy.esc : 1.0, gl = last ? cy.gamma : 0.0, dl = last ? cy.delta : 0.0;
  int up[5], dn[5];
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    int s = lane - (1 << lev);
    while (s < 0) s += nc;
    up[lev] = s;
    s = lane + (1 << lev);
    while (s >= nc) s -= nc;
    dn[lev] = s;
  }
  // ---- forward: y(i) = r(i) + rho y(i-1)
  dd2 e = x[0];
  X3D_UNROLL
  for (int m = 1; m < L; ++m) e = fma2(rho, e, x[m]);
  e = escl * e;
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {   // S(c) = e(c) + rho^len(c) S(c-1), inclusive cyclic scan
    const dd2 o = shfl2(e, up[lev]);
    e = fma2(__ldg(cy.scan + lev * 32 + lane), o, e);
  }
  int prev = lane - 1;
  prev += prev < 0 ? nc : 0;
  dd2 t = shfl2(e, prev);
  X3D_UNROLL
  for (int m = 0; m < L; ++m) { t = fma2(rho, t, x[m]); x[m] = t; }
  const dd2 yl = escl * t;
  // ---- backward: z(i) = y(i) + rho z(i+1)
  dd2 b = x[L - 1];
  X3D_UNROLL
  for (int m = L - 2; m >= 0; --m) b = fma2(rho, b, x[m]);
  b = fma2(-gl, yl, b);
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    const dd2 o = shfl2(b, dn[lev]);
    b = fma2(__ldg(cy.scan + (5 + lev) * 32 + lane), o, b);
  }
  int next = lane + 1;
  next -= next >= nc ? nc : 0;
  t = escl * fma2(-dl, yl, shfl2(b, next));
  X3D_UNROLL
  for (int m = L - 1; m >= 0; --m) { t = fma2(rho, t, x[m]); x[m] = t; }
}

// two independent solves interleaved (see pair_solve_cyclic2: four dependent chains per thread instead of two)
template <int L>
__device__ __forceinline__ void pair_solve_cyclic_ks2(dd2 (&x1)[L], const StagCyc &c1, dd2 (&x2)[L], const StagCyc &c2, int lane, int nc) {
  const double r1 = c1.rho, r2 = c2.rho;
  const bool last = lane == nc - 1;
  const double e1s = last ? c1.esc : 1.0, g1 = last ? c1.gamma : 0.0, d1 = last ? c1.delta : 0.0;
  const double e2s = last ? c2.esc : 1.0, g2 = last ? c2.gamma : 0.0, d2 = last ? c2.delta : 0.0;
  int up[5], dn[5];
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    int s = lane - (1 << lev);
    while (s < 0) s += nc;
    up[lev] = s;
    s = lane + (1 << lev);
    while (s >= nc) s -= nc;
    dn[lev] = s;
  }
  dd2 e1 = x1[0], e2 = x2[0];
  X3D_UNROLL
  for (int m = 1; m < L; ++m) { e1 = fma2(r1, e1, x1[m]); e2 = fma2(r2, e2, x2[m]); }
  e1 = e1s * e1; e2 = e2s * e2;
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    const dd2 o1 = shfl2(e1, up[lev]), o2 = shfl2(e2, up[lev]);
    e1 = fma2(__ldg(c1.scan + lev * 32 + lane), o1, e1);
    e2 = fma2(__ldg(c2.scan + lev * 32 + lane), o2, e2);
  }
  int prev = lane - 1;
  prev += prev < 0 ? nc : 0;
  dd2 t1 = shfl2(e1, prev), t2 = shfl2(e2, prev);
  X3D_UNROLL
  for (int m = 0; m < L; ++m) { t1 = fma2(r1, t1, x1[m]); x1[m] = t1; t2 = fma2(r2, t2, x2[m]); x2[m] = t2; }
  const dd2 y1 = e1s * t1, y2 = e2s * t2;
  dd2 b1 = x1[L - 1], b2 = x2[L - 1];
  X3D_UNROLL
  for (int m = L - 2; m >= 0; --m) { b1 = fma2(r1, b1, x1[m]); b2 = fma2(r2, b2, x2[m]); }
  b1 = fma2(-g1, y1, b1); b2 = fma2(-g2, y2, b2);
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) {
    const dd2 o1 = shfl2(b1, dn[lev]), o2 = shfl2(b2, dn[lev]);
    b1 = fma2(__ldg(c1.scan + (5 + lev) * 32 + lane), o1, b1);
    b2 = fma2(__ldg(c2.scan + (5 + lev) * 32 + lane), o2, b2);
  }
  int next = lane + 1;
  next -= next >= nc ? nc : 0;
  t1 = e1s * fma2(-d1, y1, shfl2(b1, next));
  t2 = e2s * fma2(-d2, y2, shfl2(b2, next));
  X3D_UNROLL
  for (int m = L - 1; m >= 0; --m) { t1 = fma2(r1, t1, x1[m]); x1[m] = t1; t2 = fma2(r2, t2, x2[m]); x2[m] = t2; }
}

// 8 consumer warps + one producer warpgroup (one active thread), registers rebalanced with setmaxnreg as in k_mom_pair:
// a ninth full-size warp would put three warps on one SM sub-partition and cap every thread at 168 registers
template <int KA, int KB, int MODE, int L>
__global__ void __launch_bounds__(MOM_THREADS, 1)
    k_stag(const __grid_constant__ DevOp opA, const __grid_constant__ DevOp opB, const __grid_constant__ StagMaps maps,
           const __grid_constant__ StagGeom g) {
  constexpr int NWIN = L + 2 * HALO;
  constexpr int NB = 3;
  constexpr int NTA = (KA == IVP || KA == IPV) ? 4 : 2, NTB = (KB == IVP || KB == IPV) ? 4 : 2;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int slot_bytes = g.slot_bytes;
  // behind the last slot: zeroed padding (the last chunk's window overruns the slot), then the barriers
  double *pad = reinterpret_cast<double *>(smem_raw + NB * slot_bytes);
  unsigned long long *full = reinterpret_cast<unsigned long long *>(pad + 512);
  unsigned long long *done = full + NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = g.n, nc = g.nc;
  for (int idx = threadIdx.x; idx < 512; idx += blockDim.x) pad[idx] = 0.0;
  if (threadIdx.x == 0) {
    X3D_UNROLL
    for (int s = 0; s < NB; ++s) { mbar_init(full + s, 1); mbar_init(done + s, PAIR_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < g.npos ? (g.npos - first + step - 1) / step : 0;
  const long long nuse = 2 * mine;   // tile uses of this CTA: use q = 2 p + s lives in slot q % 3
  // MODE 0: s = 0 -> B (load inB, no store), s = 1 -> A (load inA, store outA)
  // MODE 1: s = 0 -> A (load inA, store outA), s = 1 -> S (no load, store outB)

  if (warp >= PAIR_WARPS) {
    // ---------------- TMA producer ----------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp != PAIR_WARPS || lane != 0) return;
    const unsigned in_bytes = (static_cast<unsigned>(g.nbox) * g.br + 16u) * 128u;
    auto fill = [&](long long q) {   // make slot q % 3 ready for use q: load its input, or just hand it over
      const long long pos = first + (q >> 1) * step;
      const int s = static_cast<int>(q & 1), slot = static_cast<int>(q % NB);
      unsigned char *dst = smem_raw + slot * slot_bytes;
      if (MODE == 1 && s == 1) { mbar_arrive(full + slot); return; }
      const CUtensorMap *tm = (MODE == 0 && s == 0) ? &maps.inB : &maps.inA;
      const CUtensorMap *th = (MODE == 0 && s == 0) ? &maps.haloB : &maps.haloA;
      const int bx = static_cast<int>(pos % g.nbx), by = static_cast<int>(pos / g.nbx);
      mbar_expect_tx(full + slot, in_bytes);
      for (int b = 0; b < g.nbox; ++b) tma_load_3d(dst + (8 + b * g.br) * 128, tm, bx * 16, b * g.br, by, full + slot);
      tma_load_3d(dst, th, bx * 16, n - 8, by, full + slot);
      tma_load_3d(dst + (8 + n) * 128, th, bx * 16, 0, by, full + slot);
    };
    for (long long q = 0; q < NB && q < nuse; ++q) fill(q);
    for (long long q = 0; q < nuse; ++q) {
      const int s = static_cast<int>(q & 1), slot = static_cast<int>(q % NB);
      mbar_wait(done + slot, static_cast<unsigned>((q / NB) & 1));
      const bool stores = MODE == 1 || s == 1;
      if (stores) {
        const long long pos = first + (q >> 1) * step;
        const int bx = static_cast<int>(pos % g.nbx), by = static_cast<int>(pos / g.nbx);
        const CUtensorMap *tm = (MODE == 1 && s == 1) ? &maps.outB : &maps.outA;
        const unsigned char *src = smem_raw + slot * slot_bytes;
        for (int b = 0; b < g.nbox; ++b) tma_store_3d(tm, bx * 16, b * g.br, by, src + (8 + b * g.br) * 128);
        bulk_commit();
      }
      if (q + NB < nuse) {
        if (stores) bulk_wait_read<0>();   // the store has left shared memory: the slot can be refilled
        fill(q + NB);
      }
    }
    bulk_wait_read<0>();
    return;
  }

  // ---------------- consumers: warp = lane pair, thread = chunk ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
  const int jw = warp;
  const int cl = lane < nc ? lane : nc - 1;
  const bool live = lane < nc;
  const int q0 = cl * L;
  TileAcc<false> acc;
  acc.init(jw, q0, 0);
  const double sa = g.a.scale, sb = g.b.scale;
  for (long long p = 0; p < mine; ++p) {
    const long long qa = 2 * p, qb = 2 * p + 1;
    const int slot0 = static_cast<int>(qa % NB), slot1 = static_cast<int>(qb % NB);
    unsigned char *buf0 = smem_raw + slot0 * slot_bytes, *buf1 = smem_raw + slot1 * slot_bytes;
    const unsigned par0 = static_cast<unsigned>((qa / NB) & 1), par1 = static_cast<unsigned>((qb / NB) & 1);
    if constexpr (MODE == 0) {
      dd2 y[L];
      mbar_wait(full + slot0, par0);          // B
      {
        dd2 win[NWIN];
        X3D_UNROLL
        for (int j = 0; j < NWIN; ++j) win[j] = acc.ld(buf0, j);
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          const dd2 v = rhs_interior<KB, NTB, NWIN, dd2>(opB, win, m);
          const bool ok = live && q0 + m < n;
          y[m].x = ok ? v.x : 0.0;
          y[m].y = ok ? v.y : 0.0;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(done + slot0);   // B has been read: its slot is free
      dd2 x[L];
      mbar_wait(full + slot1, par1);          // A
      {
        dd2 win[NWIN];
        X3D_UNROLL
        for (int j = 0; j < NWIN; ++j) win[j] = acc.ld(buf1, j);
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          const dd2 v = rhs_interior<KA, NTA, NWIN, dd2>(opA, win, m);
          const bool ok = live && q0 + m < n;
          x[m].x = ok ? v.x : 0.0;
          x[m].y = ok ? v.y : 0.0;
        }
      }
      pair_solve_cyclic_ks2<L>(x, g.a, y, g.b, lane, nc);
      __syncwarp();  // every lane has read its window of A
      if (live) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m < n) {
            const dd2 xa = sa * x[m], yb = sb * y[m];
            acc.st(buf1, m + HALO, xa + yb);
          }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(done + slot1);
    } else {
      dd2 x[L], y[L];
      mbar_wait(full + slot0, par0);          // A
      {
        dd2 win[NWIN];
        X3D_UNROLL
        for (int j = 0; j < NWIN; ++j) win[j] = acc.ld(buf0, j);
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          const dd2 va = rhs_interior<KA, NTA, NWIN, dd2>(opA, win, m);
          const dd2 vb = rhs_interior<KB, NTB, NWIN, dd2>(opB, win, m);
          const bool ok = live && q0 + m < n;
          x[m].x = ok ? va.x : 0.0;
          x[m].y = ok ? va.y : 0.0;
          y[m].x = ok ? vb.x : 0.0;
          y[m].y = ok ? vb.y : 0.0;
        }
      }
      pair_solve_cyclic_ks2<L>(x, g.a, y, g.b, lane, nc);
      __syncwarp();  // every lane has read its window
      if (live) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m < n) acc.st(buf0, m + HALO, sa * x[m]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(done + slot0);
      mbar_wait(full + slot1, par1);          // S: free slot handed over by the producer
      if (live) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m < n) acc.st(buf1, m + HALO, sb * y[m]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(done + slot1);
    }
  }
}

}  // namespace x3d
