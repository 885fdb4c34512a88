// x3d_fft_kernels.cuh -- hand-written FP64 FFT passes of the periodic Poisson solve (sm_100a).
//
// decomp_2d_fft_3d of the reference (2DECOMP&FFT, call sites src/poisson.f90:330,405) is a 3-D real-to-complex transform done as
// three 1-D passes: real transforms along z (z pencils), complex ones along y and x.  The library passes (cuFFT) reach 0.51 of the
// HBM roofline on the z-strided real transforms, and the spectral array makes extra round trips between them and the spectral
// division.  These kernels do every pass at one read + one write of its arrays:
//   k_fft_z_r2c   real z lines (stride nx ny)      -> nz/2+1 complex planes      16 lanes (128 B of reals) x whole line per CTA
//   k_fft_y       complex y lines (stride nx)       in place, forward or inverse   8 lanes (128 B of complex) x whole line
//   k_fft_x_spec  complex x lines (contiguous)      forward transform, spectral factor of poisson_000 (src/poisson.f90:336-402),
//                                                   inverse transform, in place: one pass instead of three
//   k_fft_z_c2r   the inverse of the first
// Conventions of 2DECOMP&FFT / FFTW: forward sign -1, inverse +1, both unnormalised (the 1/(nx ny nz) of src/poisson.f90:333 is
// part of the spectral factor).
//
// One line of N complex points is held by N/8 threads with 8 points each, in registers.  The transform is a Stockham autosort
// FFT with radix-8 / -4 / -2 stages (N = R1 R2 ..., N in {32 .. 1024}): before EVERY stage a thread holds the points
// j + m N/8 (m = 0..7), which for radix R are 8/R complete butterflies; after the butterflies the points go to their Stockham
// positions in shared memory and come back in the same read pattern.  The first stage reads global memory directly and the last
// one writes it directly (natural order), so a line crosses shared memory stages-1 times.  Shared-memory layouts: strided passes
// keep [row][lane] with the 128-byte lane group of a row contiguous (a quarter-warp access is one row = all 32 banks once:
// conflict-free for any row pattern); the contiguous pass keeps [line][row] with one pad element per 8 rows (the stride-8
// writes of the first stage then hit 32 different banks).  Real transforms use the half-length complex transform of
// (even, odd) pairs plus the usual untangling pass through shared memory.
#pragma once
#include "x3d_common.cuh"

namespace x3d {

// ---- complex helpers: W is the forward table W[m] = exp(-2 pi i m / NW) ----
template <bool INV>
__device__ __forceinline__ double2 cmulw(double2 a, double2 w) {
  if (!INV) return make_double2(fma(a.x, w.x, -a.y * w.y), fma(a.x, w.y, a.y * w.x));
  return make_double2(fma(a.x, w.x, a.y * w.y), fma(a.y, w.x, -a.x * w.y));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ double2 cmuli(double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

template <bool INV>
__device__ __forceinline__ void bfly2(double2 &a, double2 &b) {
  const double2 t = csub(a, b);
  a = cadd(a, b);
  b = t;
}
template <bool INV>
__device__ __forceinline__ void bfly4(double2 &a0, double2 &a1, double2 &a2, double2 &a3) {
  const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmuli<INV>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}
template <bool INV>
__device__ __forceinline__ void bfly8(double2 &a0, double2 &a1, double2 &a2, double2 &a3, double2 &a4, double2 &a5, double2 &a6, double2 &a7) {
  // even and odd inputs through radix-4, then X[k] = E[k] + w8^k O[k], X[k+4] = E[k] - w8^k O[k]
  bfly4<INV>(a0, a2, a4, a6);
  bfly4<INV>(a1, a3, a5, a7);
  constexpr double h = 0.70710678118654752440;
  // w8^1 = (1 - i)/sqrt2 (forward), (1 + i)/sqrt2 (inverse); w8^2 = -i / +i; w8^3 = (-1 - i)/sqrt2 / (-1 + i)/sqrt2
  const double2 o1 = INV ? make_double2(h * (a3.x - a3.y), h * (a3.x + a3.y)) : make_double2(h * (a3.x + a3.y), h * (a3.y - a3.x));
  const double2 o2 = cmuli<INV>(a5);
  const double2 o3 = INV ? make_double2(-h * (a7.x + a7.y), h * (a7.x - a7.y)) : make_double2(h * (a7.y - a7.x), -h * (a7.x + a7.y));
  const double2 e0 = a0, e1 = a2, e2 = a4, e3 = a6, o0 = a1;
  a0 = cadd(e0, o0); a4 = csub(e0, o0);
  a1 = cadd(e1, o1); a5 = csub(e1, o1);
  a2 = cadd(e2, o2); a6 = csub(e2, o2);
  a3 = cadd(e3, o3); a7 = csub(e3, o3);
}

// radices of the stages of an N-point transform (0 = no such stage)
template <int N> struct FftRadix;
template <> struct FftRadix<16>   { static constexpr int R1 = 8, R2 = 2, R3 = 0, R4 = 0; };
template <> struct FftRadix<32>   { static constexpr int R1 = 8, R2 = 4, R3 = 0, R4 = 0; };
template <> struct FftRadix<64>   { static constexpr int R1 = 8, R2 = 8, R3 = 0, R4 = 0; };
template <> struct FftRadix<128>  { static constexpr int R1 = 8, R2 = 8, R3 = 2, R4 = 0; };
template <> struct FftRadix<256>  { static constexpr int R1 = 8, R2 = 8, R3 = 4, R4 = 0; };
template <> struct FftRadix<512>  { static constexpr int R1 = 8, R2 = 8, R3 = 8, R4 = 0; };
template <> struct FftRadix<1024> { static constexpr int R1 = 8, R2 = 8, R3 = 8, R4 = 2; };

// Twiddles.  Two table forms:
//  * strided passes (threads of a quarter-warp share the butterfly index): ONE table W[m] = exp(-2 pi i m / NW), NW = WS N, read at
//    m = t (b mod NS) N / (NS R) WS -- a broadcast within the quarter-warp;
//  * contiguous pass (neighbouring threads = neighbouring butterflies): one table per stage after the first,
//    TW_s[t-1][b] = exp(-2 pi i t (b mod NS) / (NS R)), b = 0 .. N/R-1, so that neighbouring threads read neighbouring entries.  With
//    the single table the second stage read with a stride of 128 bytes: an 8-way bank conflict, 52 % of that kernel's
//    shared-memory wavefronts (ncu, profiles/r2o_fft_ncu_summary.txt).
// Measured and dropped: taking only w from the table and its powers by multiplication (fewer shared-memory reads, but the
// dependent FP64 chains cost more: y pass 0.44 -> 0.51 ms at 512^3).
template <int N> struct FftTw {
  using F = FftRadix<N>;
  static constexpr int n2 = F::R2 ? (F::R2 - 1) * (N / F::R2) : 0;
  static constexpr int n3 = F::R3 ? (F::R3 - 1) * (N / F::R3) : 0;
  static constexpr int n4 = F::R4 ? (F::R4 - 1) * (N / F::R4) : 0;
  static constexpr int off2 = 0, off3 = n2, off4 = n2 + n3, total = n2 + n3 + n4;
};

// One Stockham stage on the 8 points of a thread (points j + m N/8 on entry).  NS = product of the earlier radices.
// WS > 0: W is the single table of a transform of length WS N; WS = 0: W is this stage's own table.  row[m] = where point m goes.
template <int N, int R, int NS, int WS, bool INV>
__device__ __forceinline__ void fft_stage(double2 (&v)[8], int (&row)[8], int j, const double2 *__restrict__ W) {
  constexpr int T = N / 8, Q = 8 / R;
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    const int b = j + u * T;
    const int k = b & (NS - 1);
    if (NS > 1) {
#pragma unroll
      for (int t = 1; t < R; ++t) {
        const double2 w = WS > 0 ? W[t * k * (N / (NS * R)) * WS] : W[(t - 1) * (N / R) + b];
        v[u + t * Q] = cmulw<INV>(v[u + t * Q], w);
      }
    }
    if (R == 8) bfly8<INV>(v[u], v[u + Q], v[u + 2 * Q], v[u + 3 * Q], v[u + 4 * Q], v[u + 5 * Q], v[u + 6 * Q], v[u + 7 * Q]);
    else if (R == 4) bfly4<INV>(v[u], v[u + Q], v[u + 2 * Q], v[u + 3 * Q]);
    else bfly2<INV>(v[u], v[u + Q]);
    const int j0 = (b / NS) * (NS * R) + k;
#pragma unroll
    for (int t = 0; t < R; ++t) row[u + t * Q] = j0 + t * NS;
  }
}

// shared-memory accessors of one line
struct AccStrided {   // [row][LX lanes]: element (row, l) at row * LX + l
  double2 *base;      // already offset by the lane
  int lx;
  __device__ __forceinline__ void st(int r, double2 v) const { base[r * lx] = v; }
  __device__ __forceinline__ double2 ld(int r) const { return base[r * lx]; }
};
struct AccContig {    // [line][row + row/8]
  double2 *base;      // already offset by the line
  __device__ __forceinline__ void st(int r, double2 v) const { base[r + (r >> 3)] = v; }
  __device__ __forceinline__ double2 ld(int r) const { return base[r + (r >> 3)]; }
};

// exchange between two stages: points to their Stockham rows, back in the read pattern j + m T.  Two CTA barriers: the second
// protects the rows against the next exchange of a faster thread.
template <int N, class Acc>
__device__ __forceinline__ void fft_exchange(double2 (&v)[8], const int (&row)[8], int j, const Acc &acc) {
  constexpr int T = N / 8;
#pragma unroll
  for (int m = 0; m < 8; ++m) acc.st(row[m], v[m]);
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = acc.ld(j + m * T);
  __syncthreads();
}

// the whole transform: v holds the points j + m N/8 on entry and the points row[m] (natural order) on exit
template <int N, int WS, bool INV, class Acc>
__device__ __forceinline__ void fft_run(double2 (&v)[8], int (&row)[8], int j, const double2 *__restrict__ W, const Acc &acc) {
  using F = FftRadix<N>;
  using TW = FftTw<N>;
  fft_stage<N, F::R1, 1, WS, INV>(v, row, j, W);
  if constexpr (F::R2 != 0) {
    fft_exchange<N>(v, row, j, acc);
    fft_stage<N, F::R2, F::R1, WS, INV>(v, row, j, W + (WS > 0 ? 0 : TW::off2));
  }
  if constexpr (F::R3 != 0) {
    fft_exchange<N>(v, row, j, acc);
    fft_stage<N, F::R3, F::R1 * F::R2, WS, INV>(v, row, j, W + (WS > 0 ? 0 : TW::off3));
  }
  if constexpr (F::R4 != 0) {
    fft_exchange<N>(v, row, j, acc);
    fft_stage<N, F::R4, F::R1 * F::R2 * F::R3, WS, INV>(v, row, j, W + (WS > 0 ? 0 : TW::off4));
  }
}

// ---- complex lines with a stride (y pass), in place ---------------------------------------------------------------
// data: (lanes, N rows, outer) complex with element strides (1, stride, ostride); a CTA takes 8 lanes x the whole line.
template <int N, bool INV>
__global__ void __launch_bounds__(N, (N <= 512 ? 2 : 1))
    k_fft_strided(double2 *__restrict__ data, long long stride, long long ostride, int lanes, long long ntiles, const double2 *__restrict__ Wg) {
  constexpr int T = N / 8, LX = 8;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2 *W = reinterpret_cast<double2 *>(fft_smem);   // [N]
  double2 *tile = W + N;                                // [N][LX]
  for (int i = threadIdx.x; i < N; i += blockDim.x) W[i] = Wg[i];
  __syncthreads();
  const int l = threadIdx.x & (LX - 1), j = threadIdx.x >> 3;
  const int nbx = (lanes + LX - 1) / LX;
  const AccStrided acc{tile + l, LX};
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int bx = static_cast<int>(t % nbx);
    const long long o = t / nbx;
    const bool ok = bx * LX + l < lanes;
    double2 *p = data + o * ostride + bx * LX + l;
    double2 v[8];
    int row[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = ok ? p[static_cast<long long>(j + m * T) * stride] : make_double2(0.0, 0.0);
    fft_run<N, 1, INV>(v, row, j, W, acc);
    if (ok) {
#pragma unroll
      for (int m = 0; m < 8; ++m) p[static_cast<long long>(row[m]) * stride] = v[m];
    }
  }
}

// lanes per CTA of the real transforms: 16 (128 B of reals, 256 B of complex per row; N threads, two CTAs per SM up to N = 512).
// Measured: 8 lanes (64-byte rows, N/2 threads, three CTAs per SM) take 0.59 instead of 0.47 / 0.53 ms at 512^3.
template <int N> struct FftZLanes { static constexpr int LX = 16, SH = 4, MINB = N <= 512 ? 2 : 1; };
// ---- real z lines -> complex half spectrum ---------------------------------------------------------------------------
// in: (lanes, N rows) real with row stride `plane`; out: (lanes, N/2+1 rows) complex with the same row stride.  M = N/2.
template <int N>
__global__ void __launch_bounds__(FftZLanes<N>::LX * N / 16, FftZLanes<N>::MINB)
    k_fft_z_r2c(const double *__restrict__ in, double2 *__restrict__ out, long long plane, long long lanes, long long ntiles,
                const double2 *__restrict__ Wg) {
  constexpr int M = N / 2, T = M / 8, LX = FftZLanes<N>::LX;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2 *W = reinterpret_cast<double2 *>(fft_smem);   // [N] table of the real length (the M-point transform reads every second entry)
  double2 *U = W;
  double2 *tile = W + N;                                // [M][LX]
  for (int i = threadIdx.x; i < N; i += blockDim.x) W[i] = Wg[i];
  __syncthreads();
  const int l = threadIdx.x & (LX - 1), j = threadIdx.x >> FftZLanes<N>::SH;
  const AccStrided acc{tile + l, LX};
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long lane = t * LX + l;
    const bool ok = lane < lanes;
    const double *p = in + lane;
    double2 v[8];
    int row[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const long long r = 2LL * (j + m * T);
      v[m] = ok ? make_double2(p[r * plane], p[(r + 1) * plane]) : make_double2(0.0, 0.0);
    }
    fft_run<M, 2, false>(v, row, j, W, acc);
#pragma unroll
    for (int m = 0; m < 8; ++m) acc.st(row[m], v[m]);
    __syncthreads();
    // X[k] = E[k] + w^k O[k], E = (Z[k] + conj Z[M-k]) / 2, O = (Z[k] - conj Z[M-k]) / (2 i)
    double2 *q = out + lane;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = j + m * T;
      const double2 a = acc.ld(k), bq = acc.ld((M - k) & (M - 1));
      const double2 b = make_double2(bq.x, -bq.y);
      const double2 E = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
      const double2 D = csub(a, b);
      const double2 O = make_double2(0.5 * D.y, -0.5 * D.x);
      const double2 X = cadd(E, cmulw<false>(O, U[k]));
      if (ok) {
        q[static_cast<long long>(k) * plane] = X;
        if (k == 0) q[static_cast<long long>(M) * plane] = make_double2(E.x - O.x, 0.0);
      }
    }
    __syncthreads();
  }
}

// ---- complex half spectrum -> real z lines (unnormalised inverse) ---------------------------------------------------
template <int N>
__global__ void __launch_bounds__(FftZLanes<N>::LX * N / 16, FftZLanes<N>::MINB)
    k_fft_z_c2r(const double2 *__restrict__ in, double *__restrict__ out, long long plane, long long lanes, long long ntiles,
                const double2 *__restrict__ Wg) {
  constexpr int M = N / 2, T = M / 8, LX = FftZLanes<N>::LX;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2 *W = reinterpret_cast<double2 *>(fft_smem);   // [N] table of the real length
  double2 *U = W;
  double2 *tile = W + N;                                // [M + 1][LX]
  for (int i = threadIdx.x; i < N; i += blockDim.x) W[i] = Wg[i];
  __syncthreads();
  const int l = threadIdx.x & (LX - 1), j = threadIdx.x >> FftZLanes<N>::SH;
  const AccStrided acc{tile + l, LX};
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long lane = t * LX + l;
    const bool ok = lane < lanes;
    const double2 *p = in + lane;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = j + m * T;
      acc.st(k, ok ? p[static_cast<long long>(k) * plane] : make_double2(0.0, 0.0));
    }
    if (j == 0) acc.st(M, ok ? p[static_cast<long long>(M) * plane] : make_double2(0.0, 0.0));
    __syncthreads();
    // Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) conj(w^k)
    double2 v[8];
    int row[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int k = j + m * T;
      const double2 a = acc.ld(k), bq = acc.ld(M - k);
      const double2 b = make_double2(bq.x, -bq.y);
      const double2 Ze = cadd(a, b);
      const double2 Zo = cmulw<true>(csub(a, b), U[k]);
      v[m] = make_double2(Ze.x - Zo.y, Ze.y + Zo.x);
    }
    __syncthreads();
    fft_run<M, 2, true>(v, row, j, W, acc);
    if (ok) {
      double *q = out + lane;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const long long r = 2LL * row[m];
        q[r * plane] = v[m].x;
        q[(r + 1) * plane] = v[m].y;
      }
    }
    __syncthreads();
  }
}

// ---- contiguous complex x lines: forward transform, spectral factor of poisson_000, inverse transform, in place ------------
struct FftSpec {
  int ny, k0;                 // lines are (j, kl) with line = j + ny kl; global spectral plane k = kl + k0
  double neg_inv_norm;        // -1 / (nx ny nz)
  double eps;
  const double *ax, *bx, *ay, *by, *az, *bz, *xk2, *yk2, *zk2, *tx, *ty, *tz;   // tables of waves() / abxyz(), as SpecArgs
};
// lines per CTA: 4 up to N = 512 (256 threads, three CTAs per SM: no register spills at 85 registers, smaller barrier groups), 8 above
template <int N> struct FftXLines { static constexpr int LL = N <= 512 ? 4 : 8, MINB = N <= 512 ? 3 : 1; };
template <int N, bool SPEC>
__global__ void __launch_bounds__(FftXLines<N>::LL * N / 8, FftXLines<N>::MINB)
    k_fft_x_spec(double2 *__restrict__ data, long long nlines, const double2 *__restrict__ Wg, const __grid_constant__ FftSpec sp, int inverse_only) {
  constexpr int T = N / 8, LL = FftXLines<N>::LL, PITCH = N + N / 8 + 1;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2 *W = reinterpret_cast<double2 *>(fft_smem);   // stage tables of the N-point transform
  double2 *tile = W + FftTw<N>::total;                  // [LL][PITCH]
  double *xt = reinterpret_cast<double *>(tile + LL * PITCH);   // [3][N]: x-dependent parts of the spectral factor
  for (int i = threadIdx.x; i < FftTw<N>::total; i += blockDim.x) W[i] = Wg[i];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if constexpr (SPEC) {
      const double f = sp.tx[i];
      xt[i] = sp.xk2[i];
      xt[N + i] = f * f;
      xt[2 * N + i] = sp.ax[i] * sp.ax[i] + sp.bx[i] * sp.bx[i];
    }
  }
  __syncthreads();
  const int j = threadIdx.x % T, l = threadIdx.x / T;
  const AccContig acc{tile + l * PITCH};
  const long long ntiles = (nlines + LL - 1) / LL;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long line = t * LL + l;
    const bool ok = line < nlines;
    double2 *p = data + line * N;
    double2 v[8];
    int row[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = ok ? p[j + m * T] : make_double2(0.0, 0.0);
    if constexpr (SPEC) {
      fft_run<N, 0, false>(v, row, j, W, acc);
      // the last Stockham stage leaves point row[m] = j + m T in slot m: exactly the read pattern of a first stage, so the inverse
      // transform starts from the registers as they are (v[m] is mode i = j + m T); the forward transform's last exchange ended
      // with a barrier, so the inverse one may write shared memory at once
      if (ok) {
        // the (y, z)-dependent parts of the factor (every thread of the line: passing them through shared memory measured slower)
        const int jy = static_cast<int>(line % sp.ny), k = static_cast<int>(line / sp.ny) + sp.k0;
        const double fy = sp.ty[jy], fz = sp.tz[2 * k];
        const double A = (fy * fz) * (fy * fz);
        const double BC = sp.yk2[jy] * (fz * fz) + sp.zk2[2 * k] * (fy * fy);
        const double wzy = sp.neg_inv_norm * ((sp.az[k] * sp.az[k] + sp.bz[k] * sp.bz[k]) * (sp.ay[jy] * sp.ay[jy] + sp.by[jy] * sp.by[jy]));
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int i = j + m * T;
          const double kk = fma(xt[i], A, xt[N + i] * BC);
          // src/poisson.f90:366 and k_spec_000s; kk is a sum of positive terms well inside the normal range, so the division is a
          // reciprocal seed + two Newton steps (correctly rounded but for the last bit) instead of the IEEE division's slow path
          double rc;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(kk));
          rc = fma(fma(-kk, rc, 1.0), rc, rc);
          rc = fma(fma(-kk, rc, 1.0), rc, rc);
          rc = fma(fma(-kk, rc, 1.0), rc, rc);
          const double g = kk < sp.eps ? 0.0 : (wzy * xt[2 * N + i]) * rc;
          v[m].x *= g;
          v[m].y *= g;
        }
      }
      fft_run<N, 0, true>(v, row, j, W, acc);
    } else {
      if (inverse_only) fft_run<N, 0, true>(v, row, j, W, acc);
      else fft_run<N, 0, false>(v, row, j, W, acc);
    }
    if (ok) {
#pragma unroll
      for (int m = 0; m < 8; ++m) p[row[m]] = v[m];
    }
    __syncthreads();
  }
}

}  // namespace x3d
