// x3d_ops_kernels.cuh -- the compact-scheme line kernels (sm_100a).
//
// One operator call = RHS stencil + tridiagonal solve along every line of a pencil
// (src/derive.f90:26-61).  HBM traffic is the algorithmic minimum, 8 B read + 8 B
// written per point: a line never goes back to global memory between the RHS, the
// forward sweep, the backward sweep and the periodic correction.
//
// Partitioned Thomas: a line of n points is cut into nc chunks of L rows.  Each
// thread owns one chunk IN REGISTERS, sweeps it with a zero carry, then the exact
// carry is injected through the precomputed products Pf/Pb of the LU multipliers
// (x3d_tables.cu):   t'(i) = t'_local(i) + Pf(i) t'(chunk_start-1)
//                    x(i)  = x_local(i)  + Pb(i) x(chunk_end+1).
// The chunk-boundary values are themselves a first-order recurrence over chunks,
// solved through shared memory (y/z kernel) or warp shuffles (x kernel).
//
//  * k_strided  -- y- and z-direction lines (stride nx or nx*ny).  A block is
//    LX lanes (consecutive i, coalesced 8*LX-byte rows) x nc chunks.
//  * k_contig   -- x-direction lines (contiguous).  One warp per line, lane = chunk;
//    the line is staged through shared memory with an ODD chunk length so that the
//    strided per-lane reads are bank-conflict free; loads/stores to HBM are fully
//    coalesced 256-byte rows.
#pragma once
#include <cuda.h>
#include "x3d_common.cuh"

namespace x3d {

#define X3D_UNROLL _Pragma("unroll")

// interior stencil; win[j] holds input q0 + j - HALO, row m has its centre at j = m + HALO
// two adjacent lanes of a line tile processed by one thread
struct dd2 {
  double x, y;
};
__device__ __forceinline__ dd2 operator+(dd2 a, dd2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ dd2 operator-(dd2 a, dd2 b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ dd2 operator*(double a, dd2 b) { return {a * b.x, a * b.y}; }
__device__ __forceinline__ dd2 fma2(double a, dd2 b, dd2 c) { return {fma(a, b.x, c.x), fma(a, b.y, c.y)}; }

template <int KIND, int NT, int NWIN, class V = double>
__device__ __forceinline__ V rhs_interior(const DevOp &op, const V (&win)[NWIN], int m) {
  const int j = m + HALO;
  V r;
  if constexpr (KIND == D1) {
    r = op.c[0] * (win[j + 1] - win[j - 1]) + op.c[1] * (win[j + 2] - win[j - 2]);
  } else if constexpr (KIND == D2) {
    r = op.c[0] * (win[j + 1] - win[j] - win[j] + win[j - 1]) + op.c[1] * (win[j + 2] - win[j] - win[j] + win[j - 2]);
    if constexpr (NT > 2)
      r = r + (op.c[2] * (win[j + 3] - win[j] - win[j] + win[j - 3]) + op.c[3] * (win[j + 4] - win[j] - win[j] + win[j - 4]));
  } else if constexpr (KIND == FIL) {
    r = op.c0 * win[j] + op.c[0] * (win[j + 1] + win[j - 1]) + op.c[1] * (win[j + 2] + win[j - 2]) +
        op.c[2] * (win[j + 3] + win[j - 3]);
  } else if constexpr (KIND == DVP) {
    r = op.c[0] * (win[j + 1] - win[j]) + op.c[1] * (win[j + 2] - win[j - 1]);
  } else if constexpr (KIND == IVP) {
    r = op.c[0] * (win[j + 1] + win[j]) + op.c[1] * (win[j + 2] + win[j - 1]) + op.c[2] * (win[j + 3] + win[j - 2]);
    if constexpr (NT > 3) r = r + op.c[3] * (win[j + 4] + win[j - 3]);
  } else if constexpr (KIND == DPV) {
    r = op.c[0] * (win[j] - win[j - 1]) + op.c[1] * (win[j + 1] - win[j - 2]);
  } else {  // IPV
    r = op.c[0] * (win[j] + win[j - 1]) + op.c[1] * (win[j + 1] + win[j - 2]) + op.c[2] * (win[j + 2] + win[j - 3]);
    if constexpr (NT > 3) r = r + op.c[3] * (win[j + 3] + win[j - 4]);
  }
  return r;
}

__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// ===========================================================================
// y / z lines
// ===========================================================================
template <int KIND, int NT, int L, int LX, int NCMAX, int MINB>
__global__ void __launch_bounds__(LX *NCMAX, MINB)
    k_strided(const __grid_constant__ DevOp op, const double *__restrict__ u, double *__restrict__ t,
              const double *__restrict__ rows, const double *__restrict__ chunkp, int nc, int n1, long long sin,
              long long sout, long long oin, long long oout) {
  constexpr int NWIN = L + 2 * HALO;
  __shared__ double sE[NCMAX][LX];
  __shared__ double sB[NCMAX][LX];
  __shared__ double sX[2][LX];
  const int lane = threadIdx.x, c = threadIdx.y;
  const long long i = static_cast<long long>(blockIdx.x) * LX + lane;
  const bool active = i < n1;
  const double *up = u + blockIdx.y * oin + (active ? i : 0);
  double *tp = t + blockIdx.y * oout + (active ? i : 0);
  const int q0 = c * L;
  const int n_in = op.n_in, n_out = op.n_out;

  // ---- window of inputs (registers) ----------------------------------------
  double win[NWIN];
  if (c > 0 && q0 + L + HALO <= n_in) {  // interior chunk: no wrap, no bounds
    X3D_UNROLL
    for (int j = 0; j < NWIN; ++j) win[j] = up[(q0 + j - HALO) * sin];
  } else {
    X3D_UNROLL
    for (int j = 0; j < NWIN; ++j) {
      int q = q0 + j - HALO;
      if (op.periodic) { q = q < 0 ? q + n_in : (q >= n_in ? q - n_in : q); }
      const bool ok = q >= 0 && q < n_in;
      win[j] = ok ? up[q * sin] : 0.0;
    }
  }
  // ---- RHS -------------------------------------------------------------------
  double x[L];
  X3D_UNROLL
  for (int m = 0; m < L; ++m) {
    const int row = q0 + m;
    double v = rhs_interior<KIND, NT, NWIN>(op, win, m);
    if (op.nb) {
      if (row < NBROW) {
        v = 0.0;
#pragma unroll 1
        for (int q = 0; q < NBCOL; ++q) v += op.wstart[row][q] * up[q * sin];
      } else if (row >= n_out - NBROW) {
        v = 0.0;
        if (row < n_out) {
#pragma unroll 1
          for (int q = 0; q < NBCOL; ++q) v += op.wend[row - (n_out - NBROW)][q] * up[(n_in - NBCOL + q) * sin];
        }
      }
    }
    x[m] = (row < n_out) ? v : 0.0;
  }
  if (!op.rhs_only) {
    const double *rw = rows + static_cast<long long>(q0) * TRI_W;
    // ---- forward sweep, zero carry -------------------------------------------
    X3D_UNROLL
    for (int m = 1; m < L; ++m) x[m] = fma(-x[m - 1], __ldg(rw + m * TRI_W + T_S), x[m]);
    sE[c][lane] = x[L - 1];
    __syncthreads();
    double cin = 0.0;
    for (int cc = 0; cc < c; ++cc) cin = fma(__ldg(chunkp + cc), cin, sE[cc][lane]);
    // ---- inject carry + backward sweep, zero carry ----------------------------
    {
      double xn = 0.0;
      X3D_UNROLL
      for (int m = L - 1; m >= 0; --m) {
        const double pf = __ldg(rw + m * TRI_W + T_PF);
        const double2 wf = ldg2(rw + m * TRI_W + T_W);
        const double tt = fma(pf, cin, x[m]);
        xn = fma(-wf.y, xn, tt * wf.x);
        x[m] = xn;
      }
    }
    sB[c][lane] = x[0];
    __syncthreads();
    double cb = 0.0;
    for (int cc = nc - 1; cc > c; --cc) cb = fma(__ldg(chunkp + nc + cc), cb, sB[cc][lane]);
    if (op.periodic) {
      // Sherman-Morrison: x -= (x_0 - alpha x_{n-1}) * rs   (src/derive.f90:55-59)
      X3D_UNROLL
      for (int m = 0; m < L; ++m) {
        x[m] = fma(__ldg(rw + m * TRI_W + T_PB), cb, x[m]);
        if (q0 + m == n_out - 1) sX[1][lane] = x[m];
      }
      if (c == 0) sX[0][lane] = x[0];
      __syncthreads();
      const double sf = sX[0][lane] - op.alpha * sX[1][lane];
      X3D_UNROLL
      for (int m = 0; m < L; ++m) x[m] = fma(-sf, __ldg(rw + m * TRI_W + T_RS), x[m]);
    } else {
      X3D_UNROLL
      for (int m = 0; m < L; ++m) x[m] = fma(__ldg(rw + m * TRI_W + T_PB), cb, x[m]);
    }
    if (op.has_post) {
      X3D_UNROLL
      for (int m = 0; m < L; ++m) x[m] *= __ldg(rw + m * TRI_W + T_POST);
    }
  }
  if (active) {
    if (op.store_mode == 0) {
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) tp[(q0 + m) * sout] = x[m];
    } else {
      const double sg = op.store_mode == 1 ? 1.0 : -1.0;
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) tp[(q0 + m) * sout] += sg * x[m];
    }
  }
}

// ===========================================================================
// x lines
// ===========================================================================
// PTX helpers: mbarrier + bulk async copy (TMA, 1-D).  SASS: UBLKCP / SYNCS.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "X3D_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra X3D_WAIT_DONE;\n"
      "bra X3D_WAIT_LOOP;\n"
      "X3D_WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// destination += shared (element-wise f64 add performed by the TMA unit; SASS UBLKRED.ADD.F64)
__device__ __forceinline__ void bulk_red_add_s2g(void *dst, const void *src, unsigned bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared memory layout (doubles):
//   coefficient columns [7][NP] | per-warp: NB line buffers [NBUF] + boundary rows [8] | mbarriers
// Element q of a line sits at buf[HALO + q]; ghosts / padding around it stay finite.
// TMA = true : lines are moved HBM<->shared by 1-D bulk async copies (cp.async.bulk + mbarrier),
//              NB-deep per-warp ring, so the load of line k+1 and the store of line k-1 overlap
//              the arithmetic of line k.  Needs n_in, n_out even (16-byte rows).
// TMA = false: plain coalesced LDG/STG through the same buffers (any n).
template <int KIND, int NT, int L, int WPB, int NB, int MINB, bool TMA>
__global__ void __launch_bounds__(32 * WPB, MINB)
    k_contig(const __grid_constant__ DevOp op, const double *__restrict__ u, double *__restrict__ t,
             const double *__restrict__ rows, const double *__restrict__ scan, int nc, long long nlines, int NP,
             int NBUF, int COEF, int CHEAD, int CHEAD_RS) {
  // NP = rows of the coefficient table: nc * L, or (2 CHEAD + 2) * L when it is compressed (CHEAD > 0: head chunks,
  // one generic chunk, CHEAD + 1 tail chunks; `rows` is then the ready shared-memory image with the Sherman-Morrison
  // column appended under its own head count CHEAD_RS, see compress_tri)
  constexpr int NWIN = L + 2 * HALO;
  extern __shared__ __align__(16) double smem[];
  double *coef = smem;  // [7][NP], COEF doubles reserved (even)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int WSTRIDE = NB * NBUF + 8;
  double *wbase = smem + COEF + warp * WSTRIDE;
  double *sb = wbase + NB * NBUF;  // [8] explicit boundary rows of the current line
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + COEF + WPB * WSTRIDE) + warp * NB;
  const int n_in = op.n_in, n_out = op.n_out;
  if (CHEAD > 0) {
    const int nimg = 7 * NP + (2 * CHEAD_RS + 2) * L;
    for (int idx = threadIdx.x; idx < nimg; idx += blockDim.x) coef[idx] = __ldg(rows + idx);
  } else {
    for (int idx = threadIdx.x; idx < NP * TRI_W; idx += blockDim.x) {
      const int r = idx / TRI_W, col = idx % TRI_W;
      if (col < 7) coef[col * NP + r] = __ldg(rows + idx);
    }
  }
  for (int q = lane; q < WSTRIDE; q += 32) wbase[q] = 0.0;
  if (TMA && lane == 0) {
    X3D_UNROLL
    for (int b = 0; b < NB; ++b) mbar_init(bars + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (TMA) fence_proxy_async();
  __syncthreads();
  const int c = lane;
  const int cl = c < nc ? c : nc - 1;  // idle lanes shadow the last chunk (results discarded)
  const int q0 = cl * L;
  const bool live = c < nc;
  const int tab = CHEAD <= 0 ? cl : (cl < CHEAD ? cl : (cl >= nc - CHEAD - 1 ? CHEAD + 1 + cl - (nc - CHEAD - 1) : CHEAD));
  const int qc = tab * L;
  const int tab_rs = cl < CHEAD_RS ? cl : (cl >= nc - CHEAD_RS - 1 ? CHEAD_RS + 1 + cl - (nc - CHEAD_RS - 1) : CHEAD_RS);
  const double *cS = coef + T_S * NP + qc, *cPF = coef + T_PF * NP + qc, *cW = coef + T_W * NP + qc,
               *cFW = coef + T_FW * NP + qc, *cPB = coef + T_PB * NP + qc,
               *cRS = CHEAD > 0 ? coef + 7 * NP + tab_rs * L : coef + T_RS * NP + qc, *cPO = coef + T_POST * NP + qc;
  const long long first = static_cast<long long>(blockIdx.x) * WPB + warp;
  const long long step = static_cast<long long>(gridDim.x) * WPB;
  const unsigned in_bytes = static_cast<unsigned>(n_in) * 8u, out_bytes = static_cast<unsigned>(n_out) * 8u;
  if (TMA && lane == 0 && first < nlines) {
    mbar_expect_tx(bars + 0, in_bytes);
    bulk_g2s(wbase + HALO, u + first * n_in, in_bytes, bars + 0);
  }
  int it = 0;
  for (long long line = first; line < nlines; line += step, ++it) {
    const int b = it % NB;
    double *buf = wbase + b * NBUF;
    if constexpr (TMA) {
      // prefetch the next line into the next ring slot (its previous store must have left shared memory)
      const long long nxt = line + step;
      if (nxt < nlines) {
        const int bn = (it + 1) % NB;
        if (lane == 0) {
          bulk_wait_read<NB - 2>();
          mbar_expect_tx(bars + bn, in_bytes);
          bulk_g2s(wbase + bn * NBUF + HALO, u + nxt * n_in, in_bytes, bars + bn);
        }
      }
      mbar_wait(bars + b, (it / NB) & 1);
    } else {
      const double *up = u + line * n_in;
      for (int q = lane; q < n_in; q += 32) buf[HALO + q] = up[q];
      __syncwarp();
    }
    if (op.periodic) {  // wrap ghosts
      if (lane < HALO) buf[HALO - 1 - lane] = buf[HALO + n_in - 1 - lane];
      else if (lane < 2 * HALO) buf[HALO + n_in + (lane - HALO)] = buf[HALO + (lane - HALO)];
      __syncwarp();
    } else if (op.nb) {  // explicit closure rows, one lane per row
      if (lane < 2 * NBROW) {
        const bool end = lane >= NBROW;
        const double *wr = end ? op.wend[lane - NBROW] : op.wstart[lane];
        const double *src = buf + HALO + (end ? n_in - NBCOL : 0);
        double v = 0.0;
        X3D_UNROLL
        for (int q = 0; q < NBCOL; ++q) v = fma(wr[q], src[q], v);
        sb[lane] = v;
      }
      __syncwarp();
    }
    double x[L];
    {
      double win[NWIN];
      X3D_UNROLL
      for (int j = 0; j < NWIN; ++j) win[j] = buf[q0 + j];
      X3D_UNROLL
      for (int m = 0; m < L; ++m) {
        const int row = q0 + m;
        double v = rhs_interior<KIND, NT, NWIN>(op, win, m);
        if (op.nb) {
          if (row < NBROW) v = sb[row];
          else if (row >= n_out - NBROW && row < n_out) v = sb[NBROW + row - (n_out - NBROW)];
        }
        x[m] = (live && row < n_out) ? v : 0.0;
      }
    }
    if (!op.rhs_only) {
      X3D_UNROLL
      for (int m = 1; m < L; ++m) x[m] = fma(-x[m - 1], cS[m], x[m]);
      // chunk-end values: v(c) = e(c) + Af(c) v(c-1)  -> Kogge-Stone over lanes
      double v = live ? x[L - 1] : 0.0;
      X3D_UNROLL
      for (int lev = 0; lev < 5; ++lev) {
        const double o = __shfl_up_sync(0xffffffffu, v, 1 << lev);
        v = fma(__ldg(scan + lev * 32 + lane), o, v);
      }
      double cin = __shfl_up_sync(0xffffffffu, v, 1);
      if (lane == 0) cin = 0.0;
      {
        double xn = 0.0;
        X3D_UNROLL
        for (int m = L - 1; m >= 0; --m) {
          const double tt = fma(cPF[m], cin, x[m]);
          xn = fma(-cFW[m], xn, tt * cW[m]);
          x[m] = xn;
        }
      }
      v = live ? x[0] : 0.0;
      X3D_UNROLL
      for (int lev = 0; lev < 5; ++lev) {
        const double o = __shfl_down_sync(0xffffffffu, v, 1 << lev);
        v = fma(__ldg(scan + (5 + lev) * 32 + lane), o, v);
      }
      double cb = __shfl_down_sync(0xffffffffu, v, 1);
      if (lane >= nc - 1) cb = 0.0;
      X3D_UNROLL
      for (int m = 0; m < L; ++m) x[m] = fma(cPB[m], cb, x[m]);
      if (op.periodic) {
        double xl = 0.0;
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m == n_out - 1) xl = x[m];
        const double x0 = __shfl_sync(0xffffffffu, x[0], 0);
        const double xe = __shfl_sync(0xffffffffu, xl, nc - 1);
        const double sf = x0 - op.alpha * xe;
        X3D_UNROLL
        for (int m = 0; m < L; ++m) x[m] = fma(-sf, cRS[m], x[m]);
      }
      if (op.has_post) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m) x[m] *= cPO[m];
      }
    }
    __syncwarp();  // every lane has read its window
    if (live) {
      const double sg = op.store_mode == 2 ? -1.0 : 1.0;
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) buf[HALO + q0 + m] = sg * x[m];
    }
    if constexpr (TMA) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (op.store_mode == 0) bulk_s2g(t + line * n_out, buf + HALO, out_bytes);
        else bulk_red_add_s2g(t + line * n_out, buf + HALO, out_bytes);
      }
    } else {
      __syncwarp();
      double *tp = t + line * n_out;
      if (op.store_mode == 0) for (int q = lane; q < n_out; q += 32) tp[q] = buf[HALO + q];
      else for (int q = lane; q < n_out; q += 32) tp[q] += buf[HALO + q];
      __syncwarp();
    }
  }
  if (TMA && lane == 0) bulk_wait_read<0>();
}


// ===========================================================================
// y / z lines, TMA-tiled (the default for strided lines)
// ===========================================================================
// A tile is LX consecutive lanes x the whole line, moved HBM -> shared memory by tensor-map TMA
// (cp.async.bulk.tensor, SASS UTMALDG) into an NB-deep ring, solved in place and sent back with a
// TMA store (UTMASTG).  The CTA is persistent: while the threads work on tile k the TMA engine is
// loading tile k+1 and draining tile k-1, so HBM never waits for the arithmetic.  Thread (lane, c)
// owns rows [c*L, c*L+L) of one lane in registers (partitioned Thomas as in k_strided); the
// first-order recurrence over the chunk boundaries is a Kogge-Stone scan with warp shuffles, one
// warp per lane, instead of a serial loop over the chunks.
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int c0, int c1, int c2, const void *src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1),
               "r"(c2), "r"(smem_u32(src))
               : "memory");
}
// tensor-map reduce: destination tile += shared tile (SASS UTMAREDG.ADD; f64 from the tensor map)
__device__ __forceinline__ void tma_red_add_3d(const CUtensorMap *tm, int c0, int c1, int c2, const void *src) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0),
               "r"(c1), "r"(c2), "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

struct TileGeom {
  int nbx;                 // lane blocks per outer slab
  long long ntiles;        // nbx * nouter
  int rows_slot;           // rows reserved per ring slot
  int nbox_in, br_in;      // input tile = nbox_in boxes of br_in rows
  int nbox_out, br_out;
};

// distributed solve of one chunk held in x[] (rows q0 .. q0+L-1 of lane `lane`): forward sweep,
// chunk-carry scan, backward sweep, carry scan, Sherman-Morrison correction, trailing multiply.
// sE/sC: [32][LX+1] exchange buffers, sX: [2][LX].  Contains 4 __syncthreads (all threads of the CTA).
template <int L, int LX>
__device__ __forceinline__ void tile_solve(double (&x)[L], const DevOp &op, const double *__restrict__ rw,
                                           const double *__restrict__ scan, int nc, int c, int lane, bool live, int q0,
                                           double (*sE)[LX + 1], double (*sC)[LX + 1], double (*sX)[LX]) {
  const int n_out = op.n_out;
  const int tid = threadIdx.x, wid = tid >> 5, wl = tid & 31, nwarps = blockDim.x >> 5;
  X3D_UNROLL
  for (int m = 1; m < L; ++m) x[m] = fma(-x[m - 1], __ldg(rw + m * TRI_W + T_S), x[m]);
  if (live) sE[c][lane] = x[L - 1];
  __syncthreads();
  for (int l = wid; l < LX; l += nwarps) {  // v(c) = e(c) + Af(c) v(c-1)
    double v = wl < nc ? sE[wl][l] : 0.0;
    X3D_UNROLL
    for (int lev = 0; lev < 5; ++lev) {
      const double o = __shfl_up_sync(0xffffffffu, v, 1 << lev);
      v = fma(__ldg(scan + lev * 32 + wl), o, v);
    }
    double cin = __shfl_up_sync(0xffffffffu, v, 1);
    if (wl == 0) cin = 0.0;
    sC[wl][l] = cin;
  }
  __syncthreads();
  const double cin = sC[c][lane];
  {
    double xn = 0.0;
    X3D_UNROLL
    for (int m = L - 1; m >= 0; --m) {
      const double pf = __ldg(rw + m * TRI_W + T_PF);
      const double2 wf = ldg2(rw + m * TRI_W + T_W);
      const double tt = fma(pf, cin, x[m]);
      xn = fma(-wf.y, xn, tt * wf.x);
      x[m] = xn;
    }
  }
  if (live) {
    sE[c][lane] = x[0];
    if (op.periodic) {
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m == n_out - 1) sX[1][lane] = x[m];  // last chunk: its backward sweep is already exact
    }
  }
  __syncthreads();
  for (int l = wid; l < LX; l += nwarps) {  // v(c) = b(c) + Ab(c) v(c+1)
    double v = wl < nc ? sE[wl][l] : 0.0;
    X3D_UNROLL
    for (int lev = 0; lev < 5; ++lev) {
      const double o = __shfl_down_sync(0xffffffffu, v, 1 << lev);
      v = fma(__ldg(scan + (5 + lev) * 32 + wl), o, v);
    }
    double cb = __shfl_down_sync(0xffffffffu, v, 1);
    if (wl >= nc - 1) cb = 0.0;
    sC[wl][l] = cb;
    if (op.periodic && wl == 0) sX[0][l] = v - op.alpha * sX[1][l];  // x_0 - alpha x_{n-1}  (src/derive.f90:55-59)
  }
  __syncthreads();
  const double cb = sC[c][lane];
  if (op.periodic) {
    const double sf = sX[0][lane];
    X3D_UNROLL
    for (int m = 0; m < L; ++m) {
      const double2 pr = ldg2(rw + m * TRI_W + T_PB);  // (Pb, rs)
      x[m] = fma(-sf, pr.y, fma(pr.x, cb, x[m]));
    }
  } else {
    X3D_UNROLL
    for (int m = 0; m < L; ++m) x[m] = fma(__ldg(rw + m * TRI_W + T_PB), cb, x[m]);
  }
  if (op.has_post) {
    X3D_UNROLL
    for (int m = 0; m < L; ++m) x[m] *= __ldg(rw + m * TRI_W + T_POST);
  }
}

// RHS of the rows of one chunk from a shared-memory tile ([row][LX])
template <int KIND, int NT, int L, int LX>
__device__ __forceinline__ void tile_rhs(double (&x)[L], const DevOp &op, const double *__restrict__ tile, int nc, int c, int lane,
                                         bool live, int q0) {
  constexpr int NWIN = L + 2 * HALO;
  const int n_in = op.n_in, n_out = op.n_out;
  const double *tl = tile + lane;
  double win[NWIN];
  if (c > 0 && q0 + L + HALO <= n_in) {
    X3D_UNROLL
    for (int j = 0; j < NWIN; ++j) win[j] = tl[(q0 + j - HALO) * LX];
  } else {
    X3D_UNROLL
    for (int j = 0; j < NWIN; ++j) {
      int q = q0 + j - HALO;
      if (op.periodic) { q = q < 0 ? q + n_in : (q >= n_in ? q - n_in : q); }
      const bool ok = q >= 0 && q < n_in;
      win[j] = ok ? tl[q * LX] : 0.0;
    }
  }
  X3D_UNROLL
  for (int m = 0; m < L; ++m) {
    const int row = q0 + m;
    double v = rhs_interior<KIND, NT, NWIN>(op, win, m);
    if (op.nb) {
      if (row < NBROW) {
        v = 0.0;
#pragma unroll 1
        for (int q = 0; q < NBCOL; ++q) v += op.wstart[row][q] * tl[q * LX];
      } else if (row >= n_out - NBROW) {
        v = 0.0;
        if (row < n_out) {
#pragma unroll 1
          for (int q = 0; q < NBCOL; ++q) v += op.wend[row - (n_out - NBROW)][q] * tl[(n_in - NBCOL + q) * LX];
        }
      }
    }
    x[m] = (live && row < n_out) ? v : 0.0;
  }
}

template <int KIND, int NT, int L, int LX, int NB, int MINB>
__global__ void __launch_bounds__(LX * 32, MINB)
    k_tile(const __grid_constant__ DevOp op, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
           const double *__restrict__ rows, const double *__restrict__ scan, int nc, const TileGeom g) {
  extern __shared__ __align__(128) double smem[];
  const int slot_doubles = g.rows_slot * LX;
  double(*sE)[LX + 1] = reinterpret_cast<double(*)[LX + 1]>(smem + NB * slot_doubles);
  double(*sC)[LX + 1] = sE + 32;
  double(*sX)[LX] = reinterpret_cast<double(*)[LX]>(sC + 32);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sX + 2);
  const int tid = threadIdx.x;
  const int lane = tid % LX;
  const int craw = tid / LX;
  const bool live = craw < nc;
  const int c = live ? craw : nc - 1;  // padding threads shadow the last chunk (results discarded)
  const int q0 = c * L;
  const int n_out = op.n_out;
  const unsigned in_bytes = static_cast<unsigned>(g.nbox_in) * g.br_in * LX * 8u;
  if (tid == 0) {
    X3D_UNROLL
    for (int b = 0; b < NB; ++b) mbar_init(bars + b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  const long long first = blockIdx.x, step = gridDim.x;
  auto issue_load = [&](long long tile, int slot) {
    const int bx = static_cast<int>(tile % g.nbx), by = static_cast<int>(tile / g.nbx);
    double *dst = smem + slot * slot_doubles;
    mbar_expect_tx(bars + slot, in_bytes);
    for (int b = 0; b < g.nbox_in; ++b) tma_load_3d(dst + b * g.br_in * LX, &tm_in, bx * LX, b * g.br_in, by, bars + slot);
  };
  if (tid == 0) {  // prologue: NB-2 tiles ahead (the in-loop prefetch adds one more)
    X3D_UNROLL
    for (int d = 0; d < NB - 2; ++d)
      if (first + d * step < g.ntiles) issue_load(first + d * step, d);
  }
  const double *rw = rows + static_cast<long long>(q0) * TRI_W;
  int it = 0;
  for (long long tile = first; tile < g.ntiles; tile += step, ++it) {
    const int slot = it % NB;
    double *buf = smem + slot * slot_doubles;
    if (tid == 0) {
      // refill the slot whose store was issued two iterations ago (at most one newer store may still be reading)
      const long long nxt = tile + (NB - 2) * step;
      if (nxt < g.ntiles) {
        bulk_wait_read<1>();
        issue_load(nxt, (it + NB - 2) % NB);
      }
    }
    mbar_wait(bars + slot, (it / NB) & 1);
    double x[L];
    tile_rhs<KIND, NT, L, LX>(x, op, buf, nc, c, lane, live, q0);
    if (!op.rhs_only) tile_solve<L, LX>(x, op, rw, scan, nc, c, lane, live, q0, sE, sC, sX);
    else __syncthreads();  // every window has been read before the tile is overwritten
    if (live) {
      double *tl = buf + lane;
      const double sg = op.store_mode == 2 ? -1.0 : 1.0;
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) tl[(q0 + m) * LX] = sg * x[m];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      const int bx = static_cast<int>(tile % g.nbx), by = static_cast<int>(tile / g.nbx);
      for (int b = 0; b < g.nbox_out; ++b) {
        if (op.store_mode == 0) tma_store_3d(&tm_out, bx * LX, b * g.br_out, by, buf + b * g.br_out * LX);
        else tma_red_add_3d(&tm_out, bx * LX, b * g.br_out, by, buf + b * g.br_out * LX);
      }
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait_read<0>();
}


// ===========================================================================
// y / z lines, warp per lane pair on a 128B-swizzled TMA tile (the default strided kernel)
// ===========================================================================
// Tile = 16 consecutive lanes (one 128-byte row) x the whole line, brought into shared memory by
// tensor-map TMA with SWIZZLE_128B into an NB-deep ring.  Warp w of the 8 consumer warps owns lanes
// (2w, 2w+1) -- the 16-byte chunk w of every row -- and solves both lines like k_contig solves one:
// thread = chunk of L rows (L odd, so that the 32 threads of a 16-byte LDS/STS hit 8 distinct swizzle
// positions per quarter warp: conflict free), carries over the chunk boundaries by Kogge-Stone shuffles.
// No CTA-wide barrier exists after set-up: consumer warps run decoupled from each other; a ninth warp
// is the TMA producer (loads ahead, stores behind), synchronised through mbarriers only.
// Slot layout (rows of 128 bytes): [8 pre-halo][n data rows][8 post-halo]; every TMA box lands on a
// 1024-byte boundary, so the swizzle is chunk ^= (physical row & 7).
struct PairGeom {
  int nbx;
  long long ntiles;
  int slot_rows;           // physical rows per ring slot (multiple of 8)
  int nbox_in, br_in, nbox_out, br_out;
  int NP;                  // rows of the coefficient table (nc * L)
  int halo;                // periodic: wrap ghosts are loaded by two extra 8-row boxes
};
constexpr int PAIR_WARPS = 8;

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ dd2 shfl_up2(dd2 v, int d) { return {__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d)}; }
__device__ __forceinline__ dd2 shfl_down2(dd2 v, int d) { return {__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d)}; }
__device__ __forceinline__ dd2 shfl2(dd2 v, int src) { return {__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)}; }

template <int KIND, int NT, int L, int NB>
__global__ void __launch_bounds__(32 * (PAIR_WARPS + 1), 1)
    k_pair(const __grid_constant__ DevOp op, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_halo,
           const __grid_constant__ CUtensorMap tm_out, const double *__restrict__ rows, const double *__restrict__ scan, int nc, const PairGeom g) {
  constexpr int NWIN = L + 2 * HALO;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int slot_bytes = g.slot_rows * 128;
  double2 *coef = reinterpret_cast<double2 *>(smem_raw + NB * slot_bytes);  // [3][NP]: (s,Pf) (w,fw) (Pb,rs)
  dd2 *sball = reinterpret_cast<dd2 *>(coef + 3 * g.NP);                     // [PAIR_WARPS][8] closure rows
  unsigned long long *full = reinterpret_cast<unsigned long long *>(sball + PAIR_WARPS * 8);
  unsigned long long *done = full + NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_in = op.n_in, n_out = op.n_out;
  for (int idx = threadIdx.x; idx < g.NP * 3; idx += blockDim.x) {
    const int r = idx / 3, col = idx % 3;
    coef[col * g.NP + r] = ldg2(rows + static_cast<long long>(r) * TRI_W + 2 * col);
  }
  if (threadIdx.x == 0) {
    X3D_UNROLL
    for (int b = 0; b < NB; ++b) { mbar_init(full + b, 1); mbar_init(done + b, PAIR_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncthreads();
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < g.ntiles ? (g.ntiles - first + step - 1) / step : 0;  // tiles of this CTA

  if (warp == PAIR_WARPS) {
    // ---------------- TMA producer ----------------
    if (lane != 0) return;
    const unsigned in_bytes = (static_cast<unsigned>(g.nbox_in) * g.br_in + (g.halo ? 16u : 0u)) * 128u;
    auto load = [&](long long k) {
      const long long tile = first + k * step;
      const int bx = static_cast<int>(tile % g.nbx), by = static_cast<int>(tile / g.nbx);
      const int slot = static_cast<int>(k % NB);
      unsigned char *dst = smem_raw + slot * slot_bytes;
      mbar_expect_tx(full + slot, in_bytes);
      for (int b = 0; b < g.nbox_in; ++b) tma_load_3d(dst + (8 + b * g.br_in) * 128, &tm_in, bx * 16, b * g.br_in, by, full + slot);
      if (g.halo) {
        tma_load_3d(dst, &tm_halo, bx * 16, n_in - 8, by, full + slot);              // rows n-8..n-1 -> pre-halo
        tma_load_3d(dst + (8 + n_in) * 128, &tm_halo, bx * 16, 0, by, full + slot);  // rows 0..7    -> post-halo
      }
    };
    for (long long k = 0; k < NB && k < mine; ++k) load(k);
    for (long long k = 0; k < mine; ++k) {
      const int slot = static_cast<int>(k % NB);
      mbar_wait(done + slot, static_cast<unsigned>((k / NB) & 1));
      const long long tile = first + k * step;
      const int bx = static_cast<int>(tile % g.nbx), by = static_cast<int>(tile / g.nbx);
      const unsigned char *src = smem_raw + slot * slot_bytes;
      for (int b = 0; b < g.nbox_out; ++b) {
        if (op.store_mode == 0) tma_store_3d(&tm_out, bx * 16, b * g.br_out, by, src + (8 + b * g.br_out) * 128);
        else tma_red_add_3d(&tm_out, bx * 16, b * g.br_out, by, src + (8 + b * g.br_out) * 128);
      }
      bulk_commit();
      if (k >= 1 && k - 1 + NB < mine) {  // refill the slot of the previous store once it has left shared memory
        bulk_wait_read<1>();
        load(k - 1 + NB);
      }
    }
    bulk_wait_read<0>();
    return;
  }

  // ---------------- consumers: warp = lane pair, thread = chunk ----------------
  const int jw = warp;                    // 16-byte chunk of every row
  const int cl = lane < nc ? lane : nc - 1;
  const bool live = lane < nc;
  const int q0 = cl * L;
  const int base = q0 + 8 - HALO;         // physical row of window entry 0
  int off[8];                             // slot offset (16-byte units) of physical row base+j is off[j & 7] + 8 j
  X3D_UNROLL
  for (int p = 0; p < 8; ++p) off[p] = base * 8 + (jw ^ ((base + p) & 7));
  const double2 *cSP = coef + q0, *cWF = coef + g.NP + q0, *cBR = coef + 2 * g.NP + q0;
  dd2 *sb = sball + warp * 8;
  for (long long k = 0; k < mine; ++k) {
    const int slot = static_cast<int>(k % NB);
    dd2 *buf = reinterpret_cast<dd2 *>(smem_raw + slot * slot_bytes);
    mbar_wait(full + slot, static_cast<unsigned>((k / NB) & 1));
    if (op.nb) {  // explicit closure rows, one lane per row
      if (lane < 2 * NBROW) {
        const bool end = lane >= NBROW;
        const double *wr = end ? op.wend[lane - NBROW] : op.wstart[lane];
        const int r0 = 8 + (end ? n_in - NBCOL : 0);
        dd2 v = {0.0, 0.0};
        X3D_UNROLL
        for (int q = 0; q < NBCOL; ++q) {
          const int pr = r0 + q;
          v = fma2(wr[q], buf[pr * 8 + (jw ^ (pr & 7))], v);
        }
        sb[lane] = v;
      }
      __syncwarp();
    }
    dd2 x[L];
    {
      dd2 win[NWIN];
      X3D_UNROLL
      for (int j = 0; j < NWIN; ++j) win[j] = buf[off[j & 7] + 8 * j];
      X3D_UNROLL
      for (int m = 0; m < L; ++m) {
        const int row = q0 + m;
        dd2 v = rhs_interior<KIND, NT, NWIN, dd2>(op, win, m);
        if (op.nb) {
          if (row < NBROW) v = sb[row];
          else if (row >= n_out - NBROW && row < n_out) v = sb[NBROW + row - (n_out - NBROW)];
        }
        const bool ok = live && row < n_out;
        x[m].x = ok ? v.x : 0.0;
        x[m].y = ok ? v.y : 0.0;
      }
    }
    if (!op.rhs_only) {
      X3D_UNROLL
      for (int m = 1; m < L; ++m) x[m] = fma2(-cSP[m].x, x[m - 1], x[m]);
      dd2 v = x[L - 1];  // chunk-end values: v(c) = e(c) + Af(c) v(c-1)
      X3D_UNROLL
      for (int lev = 0; lev < 5; ++lev) {
        const dd2 o = shfl_up2(v, 1 << lev);
        v = fma2(__ldg(scan + lev * 32 + lane), o, v);
      }
      dd2 cin = shfl_up2(v, 1);
      if (lane == 0) cin = {0.0, 0.0};
      {
        dd2 xn = {0.0, 0.0};
        X3D_UNROLL
        for (int m = L - 1; m >= 0; --m) {
          const double2 wf = cWF[m];
          const dd2 tt = fma2(cSP[m].y, cin, x[m]);
          xn = fma2(-wf.y, xn, wf.x * tt);
          x[m] = xn;
        }
      }
      v = x[0];
      if (!live) v = {0.0, 0.0};
      X3D_UNROLL
      for (int lev = 0; lev < 5; ++lev) {
        const dd2 o = shfl_down2(v, 1 << lev);
        v = fma2(__ldg(scan + (5 + lev) * 32 + lane), o, v);
      }
      dd2 cb = shfl_down2(v, 1);
      if (lane >= nc - 1) cb = {0.0, 0.0};
      if (op.periodic) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m) x[m] = fma2(cBR[m].x, cb, x[m]);
        dd2 xl = {0.0, 0.0};
        X3D_UNROLL
        for (int m = 0; m < L; ++m)
          if (q0 + m == n_out - 1) xl = x[m];
        const dd2 x0 = shfl2(x[0], 0);
        const dd2 xe = shfl2(xl, nc - 1);
        const dd2 sf = {x0.x - op.alpha * xe.x, x0.y - op.alpha * xe.y};  // src/derive.f90:55-59
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          const double rs = cBR[m].y;
          x[m].x = fma(-sf.x, rs, x[m].x);
          x[m].y = fma(-sf.y, rs, x[m].y);
        }
      } else {
        X3D_UNROLL
        for (int m = 0; m < L; ++m) x[m] = fma2(cBR[m].x, cb, x[m]);
      }
      if (op.has_post) {
        X3D_UNROLL
        for (int m = 0; m < L; ++m) {
          const double po = __ldg(rows + static_cast<long long>(q0 + m) * TRI_W + T_POST);
          x[m].x *= po;
          x[m].y *= po;
        }
      }
    }
    __syncwarp();  // every lane has read its window (and the closure rows)
    if (live) {
      const double sg = op.store_mode == 2 ? -1.0 : 1.0;
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) buf[off[(m + HALO) & 7] + 8 * (m + HALO)] = sg * x[m];
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(done + slot);
  }
}

}  // namespace x3d
