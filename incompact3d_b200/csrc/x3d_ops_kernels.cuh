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
#include "x3d_common.cuh"

namespace x3d {

#define X3D_UNROLL _Pragma("unroll")

// interior stencil; win[j] holds input q0 + j - HALO, row m has its centre at j = m + HALO
template <int KIND, int NT, int NWIN>
__device__ __forceinline__ double rhs_interior(const DevOp &op, const double (&win)[NWIN], int m) {
  const int j = m + HALO;
  double r;
  if constexpr (KIND == D1) {
    r = op.c[0] * (win[j + 1] - win[j - 1]) + op.c[1] * (win[j + 2] - win[j - 2]);
  } else if constexpr (KIND == D2) {
    r = op.c[0] * (win[j + 1] - win[j] - win[j] + win[j - 1]) + op.c[1] * (win[j + 2] - win[j] - win[j] + win[j - 2]);
    if constexpr (NT > 2)
      r += op.c[2] * (win[j + 3] - win[j] - win[j] + win[j - 3]) + op.c[3] * (win[j + 4] - win[j] - win[j] + win[j - 4]);
  } else if constexpr (KIND == FIL) {
    r = op.c0 * win[j] + op.c[0] * (win[j + 1] + win[j - 1]) + op.c[1] * (win[j + 2] + win[j - 2]) +
        op.c[2] * (win[j + 3] + win[j - 3]);
  } else if constexpr (KIND == DVP) {
    r = op.c[0] * (win[j + 1] - win[j]) + op.c[1] * (win[j + 2] - win[j - 1]);
  } else if constexpr (KIND == IVP) {
    r = op.c[0] * (win[j + 1] + win[j]) + op.c[1] * (win[j + 2] + win[j - 1]) + op.c[2] * (win[j + 3] + win[j - 2]);
    if constexpr (NT > 3) r += op.c[3] * (win[j + 4] + win[j - 3]);
  } else if constexpr (KIND == DPV) {
    r = op.c[0] * (win[j] - win[j - 1]) + op.c[1] * (win[j + 1] - win[j - 2]);
  } else {  // IPV
    r = op.c[0] * (win[j] + win[j - 1]) + op.c[1] * (win[j + 1] + win[j - 2]) + op.c[2] * (win[j + 2] + win[j - 3]);
    if constexpr (NT > 3) r += op.c[3] * (win[j + 3] + win[j - 4]);
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
    X3D_UNROLL
    for (int m = 0; m < L; ++m)
      if (q0 + m < n_out) tp[(q0 + m) * sout] = x[m];
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
             int NBUF, int COEF) {
  constexpr int NWIN = L + 2 * HALO;
  extern __shared__ __align__(16) double smem[];
  double *coef = smem;  // [7][NP], COEF doubles reserved (even)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int WSTRIDE = NB * NBUF + 8;
  double *wbase = smem + COEF + warp * WSTRIDE;
  double *sb = wbase + NB * NBUF;  // [8] explicit boundary rows of the current line
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + COEF + WPB * WSTRIDE) + warp * NB;
  const int n_in = op.n_in, n_out = op.n_out;
  for (int idx = threadIdx.x; idx < NP * TRI_W; idx += blockDim.x) {
    const int r = idx / TRI_W, col = idx % TRI_W;
    if (col < 7) coef[col * NP + r] = __ldg(rows + idx);
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
  const double *cS = coef + T_S * NP + q0, *cPF = coef + T_PF * NP + q0, *cW = coef + T_W * NP + q0,
               *cFW = coef + T_FW * NP + q0, *cPB = coef + T_PB * NP + q0, *cRS = coef + T_RS * NP + q0,
               *cPO = coef + T_POST * NP + q0;
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
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) buf[HALO + q0 + m] = x[m];
    }
    if constexpr (TMA) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) bulk_s2g(t + line * n_out, buf + HALO, out_bytes);
    } else {
      __syncwarp();
      double *tp = t + line * n_out;
      for (int q = lane; q < n_out; q += 32) tp[q] = buf[HALO + q];
      __syncwarp();
    }
  }
  if (TMA && lane == 0) bulk_wait_read<0>();
}

}  // namespace x3d
