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
template <int KIND, int NT, int L, int LX, int NCMAX>
__global__ void __launch_bounds__(LX *NCMAX)
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
// shared memory layout (doubles):  coefficient columns [7][NP] | per-warp line buffers [WPB][NBUF]
template <int KIND, int NT, int L, int WPB>
__global__ void __launch_bounds__(32 * WPB)
    k_contig(const __grid_constant__ DevOp op, const double *__restrict__ u, double *__restrict__ t,
             const double *__restrict__ rows, const double *__restrict__ scan, int nc, long long nlines, int NP,
             int NBUF) {
  constexpr int NWIN = L + 2 * HALO;
  extern __shared__ double smem[];
  double *coef = smem;                       // [7][NP]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *buf = smem + 7 * NP + warp * NBUF;  // [NBUF], element q at buf[HALO + q]
  const int n_in = op.n_in, n_out = op.n_out;
  // stage the coefficient table (SoA so that lane-strided reads are conflict free)
  for (int idx = threadIdx.x; idx < NP * TRI_W; idx += blockDim.x) {
    const int r = idx / TRI_W, col = idx % TRI_W;
    if (col < 7) coef[col * NP + r] = __ldg(rows + idx);
  }
  __syncthreads();
  const int c = lane;
  const int cl = c < nc ? c : nc - 1;  // idle lanes shadow the last chunk (results discarded)
  const int q0 = cl * L;
  const bool live = c < nc;
  // Kogge-Stone multipliers of this lane
  double mf[5], mb[5];
  X3D_UNROLL
  for (int lev = 0; lev < 5; ++lev) { mf[lev] = __ldg(scan + lev * 32 + lane); mb[lev] = __ldg(scan + (5 + lev) * 32 + lane); }

  if (lane < HALO) buf[lane] = 0.0;
  for (long long line = static_cast<long long>(blockIdx.x) * WPB + warp; line < nlines;
       line += static_cast<long long>(gridDim.x) * WPB) {
    const double *up = u + line * n_in;
    double *tp = t + line * n_out;
    // ---- coalesced load of the line into shared memory ----------------------
    for (int q = lane; q < NBUF - 2 * HALO; q += 32) buf[HALO + q] = q < n_in ? up[q] : 0.0;
    __syncwarp();
    if (op.periodic && lane < 2 * HALO) {  // wrap ghosts
      const int k = lane < HALO ? lane : lane - HALO;  // k = 0..3
      if (lane < HALO) buf[HALO - 1 - k] = up[n_in - 1 - k]; else buf[HALO + n_in + k] = up[k];
    }
    __syncwarp();
    double win[NWIN];
    X3D_UNROLL
    for (int j = 0; j < NWIN; ++j) win[j] = buf[q0 + j];
    double x[L];
    X3D_UNROLL
    for (int m = 0; m < L; ++m) {
      const int row = q0 + m;
      double v = rhs_interior<KIND, NT, NWIN>(op, win, m);
      if (op.nb) {
        if (row < NBROW) {
          v = 0.0;
#pragma unroll 1
          for (int q = 0; q < NBCOL; ++q) v += op.wstart[row][q] * buf[HALO + q];
        } else if (row >= n_out - NBROW) {
          v = 0.0;
          if (row < n_out) {
#pragma unroll 1
            for (int q = 0; q < NBCOL; ++q) v += op.wend[row - (n_out - NBROW)][q] * buf[HALO + n_in - NBCOL + q];
          }
        }
      }
      x[m] = (live && row < n_out) ? v : 0.0;
    }
    if (!op.rhs_only) {
      const double *cS = coef + T_S * NP + q0, *cPF = coef + T_PF * NP + q0, *cW = coef + T_W * NP + q0,
                   *cFW = coef + T_FW * NP + q0, *cPB = coef + T_PB * NP + q0, *cRS = coef + T_RS * NP + q0,
                   *cPO = coef + T_POST * NP + q0;
      X3D_UNROLL
      for (int m = 1; m < L; ++m) x[m] = fma(-x[m - 1], cS[m], x[m]);
      // chunk-end values: v(c) = e(c) + Af(c) v(c-1)  -> Kogge-Stone over lanes
      double v = live ? x[L - 1] : 0.0;
      X3D_UNROLL
      for (int lev = 0; lev < 5; ++lev) {
        const double o = __shfl_up_sync(0xffffffffu, v, 1 << lev);
        v = fma(mf[lev], o, v);
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
        v = fma(mb[lev], o, v);
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
    __syncwarp();
    if (live) {
      X3D_UNROLL
      for (int m = 0; m < L; ++m)
        if (q0 + m < n_out) buf[HALO + q0 + m] = x[m];
    }
    __syncwarp();
    for (int q = lane; q < n_out; q += 32) tp[q] = buf[HALO + q];
    __syncwarp();
  }
}

}  // namespace x3d
