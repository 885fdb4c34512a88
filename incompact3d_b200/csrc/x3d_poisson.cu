// x3d_poisson.cu -- spectral Poisson solver (src/poisson.f90) on one GPU.
//
//   rhs (z-pencil, pressure mesh)  ->  [even/odd reorder of every non-periodic axis, one gather]
//   -> 1-D r2c FFTs along z (stride nx*ny) -> c2c along y and x            (decomp_2d_fft_3d, :330)
//   -> fused spectral kernels: normalise, half-sample phase rotations / DCT post-processing with
//      the mirrored partner, division by the modified wavenumbers, inverse post-processing
//   -> inverse FFTs -> inverse reorder.
// The 1-D FFT passes are hand written too (x3d_fft_kernels.cuh) when the Poisson mesh has power-of-two extents
// (64 .. 1024): real z transforms, strided y transforms, contiguous x transforms -- for poisson_000 the forward x
// transform, the spectral factor and the inverse x transform are ONE kernel.  Other extents (768, 1536, odd) and
// X3D_FFT=0 use cuFFT plans (library code, like calling cuBLAS) between the hand-written kernels.  kxyz (src/poisson.f90:1733-1738,1787-1800) is NOT stored: it is
// rebuilt per mode from three 1-D tables (squared modified wavenumbers and interpolator
// transfer functions), which removes a 16 B/mode read and 1 GB of HBM at 512^3.
#include <cufft.h>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include "x3d_state.cuh"
#include "x3d_fft.cuh"
#include "x3d_fft_kernels.cuh"

namespace x3d {

int decomp_info_init(Ctx &ctx, int nx, int ny, int nz);
void decomp_info_get(Ctx &ctx, int id, x3d_decomp_info *out);
void transpose_device(Ctx &ctx, int which, const double *d_src, double *d_dst, int id, int elem);
void decomp_shape(Ctx &ctx, int *p_row, int *p_col, int *rank, int *nranks);

#define X3D_CUFFT(call)                                                                         \
  do {                                                                                          \
    cufftResult r_ = (call);                                                                    \
    if (r_ != CUFFT_SUCCESS)                                                                    \
      throw ::x3d::Error(std::string(#call) + ": cuFFT error " + std::to_string((int)r_));      \
  } while (0)

struct PoissonImpl : PoissonState {
  x3d_poisson_params p{};
  int nx = 0, ny = 0, nz = 0, nzh = 0;  // pressure mesh (global)
  int bcx = 0, bcy = 0, bcz = 0;
  // slab decomposition (p_row = 1): physical z-pencil (nx, nyl, nz), spectral y-pencil (nx, ny, nzhl)
  int nranks = 1, id_ph = -1, id_sp = -1;
  int nyl = 0, nzl = 0, nzhl = 0, k0 = 0;
  DevBuf cwz, rwork2;
  cufftHandle plan_r2c = 0, plan_c2r = 0, plan_xy = 0;
  // poisson_000: the x-y transforms and the spectral factor run plane-chunk by plane-chunk (forward FFT, factor, inverse
  // FFT of fft_chunk planes before the next chunk) so that a chunk stays in the 126 MB L2 between the three passes and
  // the spectral array crosses HBM once in each direction instead of three times.  Opt-in (X3D_FFT_CHUNK=k): it measured slower.
  cufftHandle plan_xy_chunk = 0, plan_xy_rem = 0;
  int fft_chunk = 0;
  bool plans = false;
  DevBuf cw, cwb, rwork, tables, fftwork, maps;
  int *d_map[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // [backward][axis]
  int *d_idx = nullptr;  // identity map (for axes that are split across ranks)
  // device table layout (doubles): ax,bx[nx] ay,by[ny] az,bz[nzh] | xk2[nx] yk2[ny] zk2[nzh][2] | tx[nx] ty[ny] tz[nzh][2]
  double *d_ax = nullptr, *d_bx = nullptr, *d_ay = nullptr, *d_by = nullptr, *d_az = nullptr, *d_bz = nullptr;
  double *d_xk2 = nullptr, *d_yk2 = nullptr, *d_zk2 = nullptr, *d_tx = nullptr, *d_ty = nullptr, *d_tz = nullptr;
  // stretched y mesh (matrice_refinement + inversion5_v1/v2): pre-eliminated pentadiagonal systems
  // poisson_000 on power-of-two meshes: the hand-written FFT passes of x3d_fft_kernels.cuh instead of the cuFFT plans (z real
  // transforms, y transforms, and ONE x pass that does forward transform + spectral factor + inverse transform).  X3D_FFT=0: cuFFT.
  bool own_fft = false;
  bool own_spec = false;   // poisson_000: the spectral factor inside the x pass
  bool spec_real = false;   // poisson_000 with identical (re,im) z tables: the spectral step is one real factor per mode
  int istret = 0;
  int pen_nsys = 0, pen_rows = 0;     // istret 1,2: two systems (odd / even modes) of ny/2 rows; istret 3: one of nym rows
  DevBuf pen;                         // [nsys][7][rows][nzh][nx] double2: L1 L2 INV A1 B1, and the last-block terms
  ~PoissonImpl() override {
    if (plans) { cufftDestroy(plan_r2c); cufftDestroy(plan_c2r); cufftDestroy(plan_xy); }
    if (plan_xy_chunk) cufftDestroy(plan_xy_chunk);
    if (plan_xy_rem) cufftDestroy(plan_xy_rem);
  }
};

namespace {

constexpr double EPS = 1.e-16;  // src/poisson.f90:25

struct SpecArgs {
  int nx, ny, nzh, nz;  // nzh = LOCAL number of spectral z planes; global plane = k + k0
  int k0;
  int bcx, bcy, bcz;
  double norm_x, norm_y, norm_z;
  double inv_norm;   // 1 / (nx ny nz): the reference divides three times (src/poisson.f90:333); exact for powers of two
  const double *ax, *bx, *ay, *by, *az, *bz;
  const double *xk2, *yk2, *zk2, *tx, *ty, *tz;
};

// modified wavenumber of mode (i,j,k): (re,im) pair, src/poisson.f90:1733-1738 / :1787-1800
__device__ __forceinline__ double2 kxyz_of(const SpecArgs &a, int i, int j, int kl) {
  const int k = kl + a.k0;
  const double fx = a.tx[i], fy = a.ty[j];
  const double fzr = a.tz[2 * k], fzi = a.tz[2 * k + 1];
  const double xk = a.xk2[i], yk = a.yk2[j];
  const double zkr = a.zk2[2 * k], zki = a.zk2[2 * k + 1];
  const double fxy2 = (fx * fy) * (fx * fy);
  double2 r;
  r.x = xk * ((fy * fzr) * (fy * fzr)) + yk * ((fx * fzr) * (fx * fzr)) + zkr * fxy2;
  r.y = xk * ((fy * fzi) * (fy * fzi)) + yk * ((fx * fzi) * (fx * fzi)) + zki * fxy2;
  return r;
}
__device__ __forceinline__ double2 rot_fwd(double2 c, double a, double b) { return make_double2(c.x * b + c.y * a, c.y * b - c.x * a); }
__device__ __forceinline__ double2 rot_bwd(double2 c, double a, double b) { return make_double2(c.x * b - c.y * a, c.y * b + c.x * a); }
__device__ __forceinline__ double2 neg(double2 c) { return make_double2(-c.x, -c.y); }
__device__ __forceinline__ double2 divide4(double2 c, double2 kk) {  // src/poisson.f90:545-559
  const bool z1 = fabs(kk.x) < EPS, z2 = fabs(kk.y) < EPS;
  return make_double2(z1 ? 0.0 : c.x / (-kk.x), z2 ? 0.0 : c.y / (-kk.y));
}
// half-sample DCT post-/pre-processing with the mirrored partner, src/poisson.f90:508-528,571-591
__device__ __forceinline__ double2 dct_post(double2 c, double2 p, double a, double b) {
  return make_double2(0.5 * (c.x * b + c.y * a + p.x * b - p.y * a), 0.5 * (-c.x * a + c.y * b + p.x * a + p.y * b));
}
__device__ __forceinline__ double2 dct_pre(double2 c, double2 p, double a, double b) {
  return make_double2(c.x * b - c.y * a + p.x * a + p.y * b, c.x * a + c.y * b - p.x * b + p.y * a);
}

// poisson_000, src/poisson.f90:336-402: everything between the two FFTs in one pass
__global__ void k_spec_000(SpecArgs a, double2 *__restrict__ cw) {
  const long long tot = static_cast<long long>(a.nx) * a.ny * a.nzh;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % a.nx);
    const int j = static_cast<int>((idx / a.nx) % a.ny);
    const int kl = static_cast<int>(idx / (static_cast<long long>(a.nx) * a.ny));
    const int k = kl + a.k0;
    double2 c = cw[idx];
    c.x = c.x * a.inv_norm;
    c.y = c.y * a.inv_norm;
    c = rot_fwd(c, a.az[k], a.bz[k]);
    c = rot_fwd(c, a.ay[j], a.by[j]);
    if (j + 1 > a.ny / 2 + 1) c = neg(c);
    c = rot_fwd(c, a.ax[i], a.bx[i]);
    if (i + 1 > a.nx / 2 + 1) c = neg(c);
    const double2 kk = kxyz_of(a, i, j, kl);
    if (kk.x < EPS || kk.y < EPS) c = make_double2(0.0, 0.0);  // :366
    else c = make_double2(c.x / (-kk.x), c.y / (-kk.y));
    c = make_double2(c.x * a.bz[k] - c.y * a.az[k], -c.y * a.bz[k] - c.x * a.az[k]);  // :381-384
    c = make_double2(c.x * a.by[j] + c.y * a.ay[j], c.y * a.by[j] - c.x * a.ay[j]);   // :387-391
    if (j + 1 > a.ny / 2 + 1) c = neg(c);
    c = make_double2(c.x * a.bx[i] + c.y * a.ax[i], -c.y * a.bx[i] + c.x * a.ax[i]);  // :394-398
    if (i + 1 > a.nx / 2 + 1) c = neg(c);
    cw[idx] = c;
  }
}

// poisson_000 when the (re,im) halves of the z tables coincide (always the case for a periodic z axis).  The forward
// half-cell rotations (src/poisson.f90:340-361), the division (:366-377) and the backward rotations (:381-398) then
// compose to ONE real factor per mode: with W = (bz - i az)(by - i ay) sy (bx - i ax) sx the forward chain is c W,
// the backward chain is conj(conj(c')(bz - i az))(by + i ay) sy ... = c' conj(W), and c' = -c W / kxyz, so
//   out = -c |W|^2 / (nx ny nz kxyz),   |W|^2 = (az^2+bz^2)(ay^2+by^2)(ax^2+bx^2)  (= 1 up to rounding; kept).
// One reciprocal-free division per mode, the row-constant parts of kxyz hoisted per (j,k) row, x tables in registers,
// no 64-bit index arithmetic.  Agrees with the statement-by-statement form to a few ulp (tests: 1e-11).
__global__ void __launch_bounds__(256) k_spec_000s(SpecArgs a, double2 *__restrict__ cw, int kbeg, int kcnt) {
  // planes kbeg .. kbeg + kcnt - 1 of the local spectral array (cw points at plane kbeg)
  const int rows = a.ny * kcnt;
  for (int i = threadIdx.x; i < a.nx; i += blockDim.x) {
    const double xk = a.xk2[i], fx = a.tx[i], fx2 = fx * fx;
    const double wx = a.ax[i] * a.ax[i] + a.bx[i] * a.bx[i];
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
      const int j = row % a.ny, kl = row / a.ny, k = kl + kbeg + a.k0;
      const double fy = a.ty[j], fz = a.tz[2 * k];
      const double A = (fy * fz) * (fy * fz);
      const double BC = a.yk2[j] * (fz * fz) + a.zk2[2 * k] * (fy * fy);
      const double wzy = (a.az[k] * a.az[k] + a.bz[k] * a.bz[k]) * (a.ay[j] * a.ay[j] + a.by[j] * a.by[j]);
      const double kk = fma(xk, A, fx2 * BC);
      const double g = kk < EPS ? 0.0 : (-a.inv_norm * (wzy * wx)) / kk;   // :366
      double2 *p = cw + static_cast<long long>(row) * a.nx + i;
      double2 c = *p;
      c.x *= g;
      c.y *= g;
      *p = c;
    }
  }
}

// Generic staged kernels for the non-periodic variants.  Each stage is out-of-place because the
// DCT steps read the mirrored partner.  MODE bits select what a stage does, in this order:
//   NORM, ROTZ_F, ROTY_F(+sign), ROTX_F(+sign), POSTY, POSTX, DIVIDE, ZERO010, PREX, PREY, ROTX_B(+sign), ROTY_B(+sign), ROTZ_B
enum : unsigned { S_NORM = 1, S_ROTZ_F = 2, S_ROTY_F = 4, S_ROTX_F = 8, S_POSTY = 16, S_POSTX = 32, S_DIVIDE = 64,
                  S_ZERO010 = 128, S_PREX = 256, S_PREY = 512, S_ROTX_B = 1024, S_ROTY_B = 2048, S_ROTZ_B = 4096 };

__device__ __forceinline__ double2 pointwise_fwd(const SpecArgs &a, unsigned mode, double2 c, int i, int j, int kl) {
  const int k = kl + a.k0;
  if (mode & S_NORM) { c.x = c.x * a.inv_norm; c.y = c.y * a.inv_norm; }
  if (mode & S_ROTZ_F) c = rot_fwd(c, a.az[k], a.bz[k]);
  if (mode & S_ROTY_F) { c = rot_fwd(c, a.ay[j], a.by[j]); if (j + 1 > a.ny / 2 + 1) c = neg(c); }
  if (mode & S_ROTX_F) { c = rot_fwd(c, a.ax[i], a.bx[i]); if (i + 1 > a.nx / 2 + 1) c = neg(c); }
  return c;
}

__global__ void k_spec_stage(SpecArgs a, unsigned mode, const double2 *__restrict__ in, double2 *__restrict__ out) {
  const long long tot = static_cast<long long>(a.nx) * a.ny * a.nzh;
  const long long sxy = static_cast<long long>(a.nx) * a.ny;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % a.nx);
    const int j = static_cast<int>((idx / a.nx) % a.ny);
    const int k = static_cast<int>(idx / sxy);
    // the pointwise forward part is applied to the value AND to its partner before a POST step
    double2 c = pointwise_fwd(a, mode, in[idx], i, j, k);
    if (mode & S_POSTY) {
      if (j > 0) {
        const int jp = a.ny - j;
        const double2 p = pointwise_fwd(a, mode, in[i + a.nx * static_cast<long long>(jp) + sxy * k], i, jp, k);
        c = dct_post(c, p, a.ay[j], a.by[j]);
      }
    }
    if (mode & S_POSTX) {
      if (i > 0) {
        const int ip = a.nx - i;
        const double2 p = pointwise_fwd(a, mode, in[ip + a.nx * static_cast<long long>(j) + sxy * k], ip, j, k);
        c = dct_post(c, p, a.ax[i], a.bx[i]);
      }
    }
    if (mode & S_DIVIDE) c = divide4(c, kxyz_of(a, i, j, k));
    if (mode & S_ZERO010) { if (i + 1 == a.nx / 2 + 1 && k + a.k0 + 1 == a.nz / 2 + 1) c = make_double2(0.0, 0.0); }  // :902-908
    if (mode & S_PREX) {
      if (i > 0) c = dct_pre(c, in[(a.nx - i) + a.nx * static_cast<long long>(j) + sxy * k], a.ax[i], a.bx[i]);
    }
    if (mode & S_PREY) {
      if (j > 0) c = dct_pre(c, in[i + a.nx * static_cast<long long>(a.ny - j) + sxy * k], a.ay[j], a.by[j]);
    }
    if (mode & S_ROTX_B) { c = rot_bwd(c, a.ax[i], a.bx[i]); if (i + 1 > a.nx / 2 + 1) c = neg(c); }
    if (mode & S_ROTY_B) { c = rot_bwd(c, a.ay[j], a.by[j]); if (j + 1 > a.ny / 2 + 1) c = neg(c); }
    if (mode & S_ROTZ_B) c = rot_bwd(c, a.az[k + a.k0], a.bz[k + a.k0]);
    out[idx] = c;
  }
}

// even/odd reordering of the non-periodic axes, all axes in one gather driven by host-built
// index maps (src/poisson.f90:435-444, 1047-1101 forward; :643-652, 1422-1458 backward)
__global__ void k_reorder(const double *__restrict__ in, double *__restrict__ out, int nx, int ny, int nz,
                          const int *__restrict__ mx, const int *__restrict__ my, const int *__restrict__ mz) {
  const long long tot = static_cast<long long>(nx) * ny * nz;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % nx);
    const int j = static_cast<int>((idx / nx) % ny);
    const int k = static_cast<int>(idx / (static_cast<long long>(nx) * ny));
    out[idx] = in[mx[i] + static_cast<long long>(nx) * (my[j] + static_cast<long long>(ny) * mz[k])];
  }
}

// ---- host tables (src/poisson.f90:1469-1526 abxyz, :1530-1810 waves) -----------------------
struct AxisTables {
  std::vector<double> a, b;    // sin/cos twiddles
  std::vector<double> k2;      // squared modified wavenumber (re[,im])
  std::vector<double> tf;      // interpolator transfer function (re[,im])
  std::vector<double> es;      // exs / eys / ezs of waves() (re[,im])
  std::vector<double> kraw;    // n k'(w), the modified wavenumber before the division by the length (yky when istret /= 0)
};

// one direction; n = pressure-mesh points (nm), nv = velocity nodes, per = periodic,
// half = true for z (only n/2+1 modes; non-periodic z carries two wavenumbers per mode)
AxisTables axis_tables(int nm, int nv, bool per, double len, const x3d_deriv_coeffs &c, bool half) {
  AxisTables T;
  const double pi = std::acos(-1.0), twopi = 2.0 * std::acos(-1.0);
  const double d = len / static_cast<double>(nm);
  const int na = half ? nm / 2 + 1 : nm;
  T.a.resize(half ? na : nm); T.b.resize(half ? na : nm);
  for (int i = 0; i < static_cast<int>(T.a.size()); ++i) {
    const double arg = per ? static_cast<double>(i) * pi / static_cast<double>(nm) : static_cast<double>(i) * pi * 0.5 / static_cast<double>(nm);
    T.a[i] = std::sin(arg); T.b[i] = std::cos(arg);
  }
  auto kmod = [&](double w) {  // modified wavenumber of the staggered derivative, :1569-1570
    double wp = c.aci6 * 2.0 * d * std::sin(w * 0.5) + (c.bci6 * 2.0 * d) * std::sin(3.0 * 0.5 * w);
    return wp / (1.0 + 2.0 * c.alcai6 * std::cos(w));
  };
  auto transfer = [&](double e) {  // (ytt1+ytt)/yt1 with e = exs*dx, :1716-1731
    const double tt = 2.0 * (c.bici6 * std::cos(e * 1.5) + c.cici6 * std::cos(e * 2.5) + c.dici6 * std::cos(e * 3.5));
    const double tt1 = 2.0 * c.aici6 * std::cos(e * 0.5);
    const double t1 = 1.0 + 2.0 * c.ailcai6 * std::cos(e);
    return (tt1 + tt) / t1;
  };
  const int w2 = half ? 2 : 1;
  T.k2.assign(static_cast<size_t>(na) * w2, 0.0);
  T.tf.assign(static_cast<size_t>(na) * w2, 0.0);
  std::vector<double> &es = T.es;
  es.assign(static_cast<size_t>(na) * w2, 0.0);
  T.kraw.assign(static_cast<size_t>(na) * w2, 0.0);
  if (!half) {
    if (per) {
      for (int i = 0; i <= nm / 2; ++i) {
        const double w = twopi * i / nv;
        const double v = nv * kmod(w) / len;
        T.k2[i] = v * v; es[i] = nv * w / len;
        T.kraw[i] = nv * kmod(w);
      }
      for (int i = nm / 2 + 1; i < nm; ++i) { T.k2[i] = T.k2[nm - i]; es[i] = es[nm - i]; T.kraw[i] = T.kraw[nm - i]; }
    } else {
      for (int i = 1; i < nm; ++i) {
        const double w = twopi * 0.5 * i / nm;
        const double v = nm * kmod(w) / len;
        T.k2[i] = v * v; es[i] = nm * w / len;
        T.kraw[i] = nm * kmod(w);
      }
    }
    for (int i = 0; i < nm; ++i) T.tf[i] = transfer(es[i] * d);
  } else {
    for (int k = 0; k < na; ++k) {
      if (per) {
        const double w = twopi * k / nv;
        const double v = nv * kmod(w) / len;
        T.k2[2 * k] = T.k2[2 * k + 1] = v * v;
        es[2 * k] = es[2 * k + 1] = nv * w / len;
      } else {  // :1646-1657: real part = mode k, imaginary part = mode nzm-k
        const double w = pi * k / nm, w1 = pi * (nm - k) / nm;
        const double v = nm * kmod(w) / len, v1 = nm * kmod(w1) / len;
        T.k2[2 * k] = v * v; T.k2[2 * k + 1] = v1 * v1;
        es[2 * k] = nm * w / len; es[2 * k + 1] = nm * w1 / len;
      }
      T.tf[2 * k] = transfer(es[2 * k] * d);
      T.tf[2 * k + 1] = transfer(es[2 * k + 1] * d);
    }
  }
  return T;
}

// source index of output point q: forward  out(i) = in(2(i-1)+1) | in(2n-2i+2)   (1-based, :438-441)
//                                   backward out(2i-1) = in(i), out(2i) = in(n-i+1)   (:646-649)
std::vector<int> reorder_map(int n, bool active, bool backward) {
  std::vector<int> m(n);
  for (int q = 0; q < n; ++q) {
    if (!active) m[q] = q;
    else if (!backward) m[q] = (q < n / 2) ? 2 * q : 2 * n - 2 * q - 1;
    else m[q] = (q % 2 == 0) ? q / 2 : n - 1 - (q - 1) / 2;
  }
  return m;
}

int grid_for(long long n, int sm) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(sm) * 16;
  return static_cast<int>(b < cap ? b : cap);
}

}  // namespace

// ---- stretched y mesh ---------------------------------------------------------------------------------
// matrice_refinement (src/poisson.f90:1814-2249) builds, for every (kx,kz), a pentadiagonal matrix in y whose
// real and imaginary parts are two independent real systems; inversion5_v1/v2 (src/tools.f90:1225-1498)
// eliminate it without pivoting on EVERY solve.  The matrix depends on the mesh only, so here the elimination is
// done once on the host, in the reference's loop order (including its carried-over multipliers when a pivot is
// exactly zero), and the solve becomes a banded forward/backward substitution on the device.
namespace {

constexpr int PEN_PLANES = 7;  // L1 L2 INV A1 B1 | (T, b1) | (PI, A1l)
constexpr double PEN_EPS = 1.e-16;  // src/tools.f90:1240

struct PentaHost {
  int nx, rows, nk;
  std::vector<double> band[5];  // [row][k][i][2]
  size_t idx(int i, int row, int k) const { return ((static_cast<size_t>(row) * nk + k) * nx + i) * 2; }
};

// in-place elimination of one set of systems; fills the device table planes
void penta_eliminate(PentaHost &H, std::vector<double> &planes) {
  const int nx = H.nx, n = H.rows, nk = H.nk;
  const size_t PS = static_cast<size_t>(n) * nk * nx * 2;
  planes.assign(PS * PEN_PLANES, 0.0);
  auto P = [&](int pl, int i, int row, int k, int c) -> double & { return planes[pl * PS + H.idx(i, row, k) + c]; };
  auto A = [&](int b, int i, int row, int k, int c) -> double & { return H.band[b - 1][H.idx(i, row, k) + c]; };  // b = 1..5
  double tmp[2] = {0.0, 0.0};
  for (int m = 0; m < n - 2; ++m)            // tools.f90:1263-1287
    for (int ii = 1; ii <= 2; ++ii) {
      const int mi = m + ii;
      for (int k = 0; k < nk; ++k)
        for (int i = 0; i < nx; ++i)
          for (int c = 0; c < 2; ++c) {
            if (A(3, i, m, k, c) != 0.0) tmp[c] = A(3 - ii, i, mi, k, c) / A(3, i, m, k, c);
            P(ii - 1, i, m, k, c) = tmp[c];
            for (int jc = 4 - ii; jc <= 5 - ii; ++jc) A(jc, i, mi, k, c) -= tmp[c] * A(jc + ii, i, m, k, c);
          }
    }
  for (int k = 0; k < nk; ++k)                // tools.f90:1289-1333
    for (int i = 0; i < nx; ++i)
      for (int c = 0; c < 2; ++c) {
        const double piv = A(3, i, n - 2, k, c);
        const double s = std::fabs(piv) > PEN_EPS ? A(2, i, n - 1, k, c) / piv : 0.0;
        const double b1 = A(3, i, n - 1, k, c) - s * A(4, i, n - 2, k, c);
        P(5, i, 0, k, c) = std::fabs(b1) > PEN_EPS ? s / b1 : 0.0;
        P(5, i, 1, k, c) = b1;
        const double pinv = std::fabs(piv) > PEN_EPS ? 1.0 / piv : 0.0;
        P(6, i, 0, k, c) = pinv;
        P(6, i, 1, k, c) = A(4, i, n - 2, k, c) * pinv;
      }
  for (int row = n - 3; row >= 0; --row)      // tools.f90:1335-1357
    for (int k = 0; k < nk; ++k)
      for (int i = 0; i < nx; ++i)
        for (int c = 0; c < 2; ++c) {
          const double piv = A(3, i, row, k, c);
          const double inv = std::fabs(piv) > PEN_EPS ? 1.0 / piv : 0.0;
          P(2, i, row, k, c) = inv;
          P(3, i, row, k, c) = A(4, i, row, k, c) * inv;
          P(4, i, row, k, c) = A(5, i, row, k, c) * inv;
        }
}

__global__ void k_penta(double2 *__restrict__ e, long long off0, long long srow, long long sk, int rows, int nx, int nk,
                        const double2 *__restrict__ C, int zero_i, int zero_k) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(nx) * nk) return;
  const int i = static_cast<int>(idx % nx), k = static_cast<int>(idx / nx);
  const long long PS = static_cast<long long>(rows) * nk * nx;
  const double2 *c0 = C + static_cast<long long>(k) * nx + i;
  const long long cs = static_cast<long long>(nk) * nx;  // coefficient row stride
  double2 *e0p = e + off0 + i + sk * k;
  auto mul = [](double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); };
  double2 x0 = e0p[0], x1 = e0p[srow];
  for (int m = 0; m < rows - 2; ++m) {
    double2 x2 = e0p[srow * (m + 2)];
    const double2 l1 = c0[0 * PS + cs * m], l2 = c0[1 * PS + cs * m];
    x1.x -= l1.x * x0.x; x1.y -= l1.y * x0.y;
    x2.x -= l2.x * x0.x; x2.y -= l2.y * x0.y;
    e0p[srow * m] = x0;
    x0 = x1; x1 = x2;
  }
  const double2 T = c0[5 * PS], b1 = c0[5 * PS + cs], PI = c0[6 * PS], A1l = c0[6 * PS + cs];
  double2 en, em;
  en.x = fabs(b1.x) > PEN_EPS ? x1.x / b1.x - T.x * x0.x : 0.0;
  en.y = fabs(b1.y) > PEN_EPS ? x1.y / b1.y - T.y * x0.y : 0.0;
  em.x = x0.x * PI.x - A1l.x * en.x;
  em.y = x0.y * PI.y - A1l.y * en.y;
  const bool zero = (i == zero_i && k == zero_k);  // src/poisson.f90:901-908
  const double2 z = make_double2(0.0, 0.0);
  e0p[srow * (rows - 1)] = zero ? z : en;
  e0p[srow * (rows - 2)] = zero ? z : em;
  double2 y2 = en, y1 = em;
  for (int row = rows - 3; row >= 0; --row) {
    const double2 inv = c0[2 * PS + cs * row], a1 = c0[3 * PS + cs * row], bb = c0[4 * PS + cs * row];
    const double2 ev = e0p[srow * row];
    double2 r;
    r.x = ev.x * inv.x - a1.x * y1.x - bb.x * y2.x;
    r.y = ev.y * inv.y - a1.y * y1.y - bb.y * y2.y;
    e0p[srow * row] = zero ? z : r;
    y2 = y1; y1 = r;
  }
  (void)mul;
}

}  // namespace

// builds and pre-eliminates the systems for the local spectral planes [k0, k0+nk)
static void penta_init(Ctx &ctx, PoissonImpl &P, const AxisTables &TX, const AxisTables &TY, const AxisTables &TZ, double dz,
                       const x3d_deriv_coeffs &cz) {
  const int nx = P.nx, ny = P.ny, nk = P.nzhl, k0 = P.k0;
  const int istret = P.istret;
  const double pi = std::acos(-1.0);
  const double alpha = P.p.alpha, beta = P.p.beta;
  if (!(beta > 0.0)) throw Error("x3d_poisson_init: istret != 0 needs beta > 0 and the alpha computed by stretching()");
  // transfer functions and wavenumbers per direction (matrice_refinement :1860-1926)
  std::vector<double> tzr(nk), tzi(nk);
  for (int kl = 0; kl < nk; ++kl) {
    const int k = kl + k0;
    if (P.bcz == 0) { tzr[kl] = TZ.tf[2 * k]; tzi[kl] = TZ.tf[2 * k]; }
    else {  // the bcz=1 branch of matrice_refinement has no dici6 term (:1910-1915)
      auto tr = [&](double e) {
        const double tt = 2.0 * (cz.bici6 * std::cos(e * 1.5) + cz.cici6 * std::cos(e * 2.5));
        const double tt1 = 2.0 * cz.aici6 * std::cos(e * 0.5);
        return (tt1 + tt) / (1.0 + 2.0 * cz.ailcai6 * std::cos(e));
      };
      tzr[kl] = tr(TZ.es[2 * k] * dz);
      tzi[kl] = tr(TZ.es[2 * k + 1] * dz);
    }
  }
  // cw(i,jy,k) = transx(i) * (yky(jy) * transz(k)) per component; yky has equal parts (jy 0-based pressure-mesh mode)
  auto cw = [&](int i, int jy, int kl, int c) { return TX.tf[i] * (TY.kraw[jy] * (c == 0 ? tzr[kl] : tzi[kl])); };
  auto xk2 = [&](int i) { return TX.k2[i]; };
  auto zk2 = [&](int kl, int c) { return TZ.k2[2 * (kl + k0) + c]; };
  auto base = [&](int i, int jy, int kl, int c) {
    const double tz2 = (c == 0 ? tzr[kl] * tzr[kl] : tzi[kl] * tzi[kl]);
    const double ty2 = TY.tf[jy] * TY.tf[jy], tx2 = TX.tf[i] * TX.tf[i];
    return xk2(i) * ty2 * tz2 + zk2(kl, c) * ty2 * tx2;
  };
  const double xa0 = alpha / pi + 0.5 / beta / pi;
  std::vector<double> planes;
  if (istret == 1 || istret == 2) {
    const double xa1 = (istret == 1) ? +1.0 / 4.0 / beta / pi : -1.0 / 4.0 / beta / pi;
    const double xa0_2 = xa0 * xa0, xa1_2 = xa1 * xa1, xa01 = xa0 * xa1, xa0p1_2 = (xa0 + xa1) * (xa0 + xa1);
    const int n = ny / 2;  // ny is even (checked by poisson_init)
    if (n < 4) throw Error("x3d_poisson_init: stretched mesh needs ny/2 >= 4");
    P.pen_nsys = 2; P.pen_rows = n;
    const size_t PS = static_cast<size_t>(n) * nk * nx * 2 * PEN_PLANES;
    P.pen.reserve(2 * PS * sizeof(double));
    for (int sys = 0; sys < 2; ++sys) {  // sys 0: a (modes 2j-1), sys 1: a2 (modes 2j)
      PentaHost H{nx, n, nk, {}};
      for (auto &b : H.band) b.assign(static_cast<size_t>(n) * nk * nx * 2, 0.0);
      auto W = [&](int i, int j, int kl, int c) { return cw(i, 2 * j + sys, kl, c); };  // j 0-based row
      for (int kl = 0; kl < nk; ++kl)
        for (int i = 0; i < nx; ++i)
          for (int c = 0; c < 2; ++c) {
            auto A = [&](int b, int row) -> double & { return H.band[b - 1][H.idx(i, row, kl) + c]; };
            for (int j = 0; j < n; ++j) {
              const double w = W(i, j, kl, c);
              // modes: row j of a is pressure mode 2j (0-based), of a2 mode 2j+1; the last row of a uses
              // transy(ny-2) (1-based, velocity ny), i.e. the same mode 2(n-1) (:1993), of a2 transy(ny-1)
              const int jy = 2 * j + sys;
              double d = base(i, jy, kl, c);
              if (j > 0 && j < n - 1) d += xa0_2 * w * w + xa1_2 * w * (W(i, j - 1, kl, c) + W(i, j + 1, kl, c));   // :1957-1975
              else if (j == 0) d += (sys == 0 ? xa0_2 : xa0_2 - xa1_2) * w * w + xa1_2 * w * W(i, 1, kl, c);         // :1980-1987,2002-2009
              else d += (sys == 0 ? xa0_2 : xa0p1_2) * w * w + xa1_2 * w * W(i, n - 2, kl, c);                       // :1991-1998,2013-2020
              A(3, j) = -d;
            }
            for (int j = 1; j < n - 1; ++j) A(4, j) = xa01 * (W(i, j + 1, kl, c) * (W(i, j, kl, c) + W(i, j + 1, kl, c)));  // :2027-2036
            if (sys == 0) {
              A(4, 0) = 2.0 * xa01 * (W(i, 0, kl, c) * W(i, 1, kl, c) + W(i, 1, kl, c) * W(i, 1, kl, c));  // :2040
            } else {
              A(4, 0) = (xa0 - xa1) * xa1 * (W(i, 0, kl, c) * W(i, 1, kl, c)) + xa0 * xa1 * (W(i, 1, kl, c) * W(i, 1, kl, c));        // :2043
              A(4, n - 2) = xa0 * xa1 * W(i, n - 2, kl, c) * W(i, n - 1, kl, c) + (xa0 + xa1) * xa1 * (W(i, n - 1, kl, c) * W(i, n - 1, kl, c));  // :2046
              A(4, n - 1) = 0.0;
            }
            for (int j = 0; j < n - 2; ++j) A(5, j) = xa1_2 * (-W(i, j + 1, kl, c) * W(i, j + 2, kl, c));  // :2058-2063
            if (sys == 0) A(5, 0) = 2.0 * A(5, 0);                                                          // :2067
            A(5, n - 2) = 0.0; A(5, n - 1) = 0.0;
            for (int j = 1; j < n; ++j) A(2, j) = xa01 * (W(i, j - 1, kl, c) * (W(i, j, kl, c) + W(i, j - 1, kl, c)));  // :2078-2085
            A(2, 0) = 0.0;
            if (sys == 1) {
              A(2, 1) = xa0 * xa1 * (W(i, 1, kl, c) * W(i, 0, kl, c)) + (xa0 + xa1) * xa1 * (W(i, 0, kl, c) * W(i, 0, kl, c));          // :2090
              A(2, n - 1) = (xa0 + xa1) * xa1 * (W(i, n - 1, kl, c) * W(i, n - 2, kl, c)) + xa0 * xa1 * (W(i, n - 2, kl, c) * W(i, n - 2, kl, c));  // :2096
            }
            for (int j = 2; j < n; ++j) A(1, j) = xa1_2 * (-W(i, j - 1, kl, c) * W(i, j - 2, kl, c));  // :2108-2113
            A(1, 0) = 0.0; A(1, 1) = 0.0;
          }
      if (sys == 0)  // not to have a singular matrix, :2120-2129
        for (int kl = 0; kl < nk; ++kl)
          for (int i = 0; i < nx; ++i)
            if (TX.k2[i] == 0.0 && TZ.k2[2 * (kl + k0)] == 0.0)
              for (int c = 0; c < 2; ++c) {
                H.band[2][H.idx(i, 0, kl) + c] = 1.0; H.band[3][H.idx(i, 0, kl) + c] = 0.0; H.band[4][H.idx(i, 0, kl) + c] = 0.0;
              }
      penta_eliminate(H, planes);
      X3D_CUDA(cudaMemcpyAsync(static_cast<double *>(P.pen.p) + sys * PS, planes.data(), PS * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
      X3D_CUDA(cudaStreamSynchronize(ctx.stream));
    }
  } else {  // istret = 3, :2131-2246
    const double xa1 = -1.0 / 4.0 / beta / pi;
    const double xa0_2 = xa0 * xa0, xa1_2 = xa1 * xa1, xa01 = xa0 * xa1;
    const int n = ny;
    if (n < 4) throw Error("x3d_poisson_init: stretched mesh needs nym >= 4");
    P.pen_nsys = 1; P.pen_rows = n;
    const size_t PS = static_cast<size_t>(n) * nk * nx * 2 * PEN_PLANES;
    P.pen.reserve(PS * sizeof(double));
    PentaHost H{nx, n, nk, {}};
    for (auto &b : H.band) b.assign(static_cast<size_t>(n) * nk * nx * 2, 0.0);
    for (int kl = 0; kl < nk; ++kl)
      for (int i = 0; i < nx; ++i)
        for (int c = 0; c < 2; ++c) {
          auto A = [&](int b, int row) -> double & { return H.band[b - 1][H.idx(i, row, kl) + c]; };
          auto W = [&](int j) { return cw(i, j, kl, c); };
          for (int j = 0; j < n; ++j) {
            const double w = W(j);
            double d = base(i, j, kl, c) + xa0_2 * w * w;
            if (j == 0) d += xa1_2 * w * W(1);
            else if (j == n - 1) d += xa1_2 * w * W(n - 2);
            else d += xa1_2 * w * (W(j - 1) + W(j + 1));
            A(3, j) = -d;
          }
          for (int j = 0; j < n - 1; ++j) A(4, j) = xa01 * (W(j + 1) * (W(j) + W(j + 1)));
          for (int j = 0; j < n - 2; ++j) A(5, j) = -xa1_2 * (W(j + 1) * W(j + 2));
          for (int j = 1; j < n; ++j) A(2, j) = xa01 * (W(j - 1) * (W(j) + W(j - 1)));
          for (int j = 2; j < n; ++j) A(1, j) = -xa1_2 * (W(j - 1) * W(j - 2));
        }
    if (k0 == 0)  // :2240-2245: element (1,1,1) of the rank that owns it
      for (int c = 0; c < 2; ++c) { H.band[2][H.idx(0, 0, 0) + c] = 1.0; H.band[3][H.idx(0, 0, 0) + c] = 0.0; H.band[4][H.idx(0, 0, 0) + c] = 0.0; }
    penta_eliminate(H, planes);
    X3D_CUDA(cudaMemcpyAsync(P.pen.p, planes.data(), PS * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  }
}

// in-place solve of the spectral y-pencil cw (nx, ny, nk)
static void penta_solve(Ctx &ctx, PoissonImpl &P, double2 *cw, bool zero010) {
  const int nx = P.nx, ny = P.ny, nk = P.nzhl, rows = P.pen_rows;
  const long long tot = static_cast<long long>(nx) * nk;
  if (tot == 0) return;
  const size_t PS = static_cast<size_t>(rows) * nk * nx * PEN_PLANES;  // double2 per system
  int zi = -1, zk = -1;
  if (zero010) { zi = nx / 2; zk = P.nz / 2 - P.k0; }
  ProfScope ps(ctx, "poisson_penta(k_penta)");
  for (int sys = 0; sys < P.pen_nsys; ++sys) {
    const long long off0 = (P.pen_nsys == 2) ? static_cast<long long>(sys) * nx : 0;
    const long long srow = (P.pen_nsys == 2) ? 2LL * nx : nx;
    k_penta<<<static_cast<unsigned>((tot + 127) / 128), 128, 0, ctx.stream>>>(cw, off0, srow, static_cast<long long>(nx) * ny, rows, nx, nk,
                                                                              static_cast<const double2 *>(P.pen.p) + sys * PS, zi, zk);
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  }
}

void poisson_init(Ctx &ctx, const x3d_poisson_params &p) {
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (p.istret != 0 && p.bcy == 0) throw Error("x3d_poisson_init: a stretched y mesh needs non-periodic y (src/poisson.f90:225)");
  if (p.istret < 0 || p.istret > 3) throw Error("x3d_poisson_init: istret must be 0..3");
  for (int a = 0; a < 3; ++a)
    if (!ctx.have_dc[a]) throw Error("x3d_poisson_init: call x3d_set_deriv_coeffs for the three axes first (waves() reads derivX/Y/Z)");
  auto P = std::make_unique<PoissonImpl>();
  P->p = p;
  P->istret = p.istret;
  P->bcx = p.bcx; P->bcy = p.bcy; P->bcz = p.bcz;
  const bool ok = (p.bcx == 0 && p.bcy == 0 && p.bcz == 0) || (p.bcx == 1 && p.bcy == 0 && p.bcz == 0) ||
                  (p.bcx == 0 && p.bcy == 1 && p.bcz == 0) || (p.bcx == 1 && p.bcy == 1);
  if (!ok) throw Error("boundary condition not supported (src/poisson.f90:107)");
  P->nx = p.bcx ? p.nx - 1 : p.nx; P->ny = p.bcy ? p.ny - 1 : p.ny; P->nz = p.bcz ? p.nz - 1 : p.nz;
  P->nzh = P->nz / 2 + 1;
  const int nx = P->nx, ny = P->ny, nz = P->nz, nzh = P->nzh;
  if ((p.bcx && nx % 2) || (p.bcy && ny % 2) || (p.bcz && nz % 2)) throw Error("non-periodic pressure mesh extents must be even");
  AxisTables TX = axis_tables(nx, p.nx, p.bcx == 0, p.xlx, ctx.dc[0], false);
  AxisTables TY = axis_tables(ny, p.ny, p.bcy == 0, p.yly, ctx.dc[1], false);
  AxisTables TZ = axis_tables(nz, p.nz, p.bcz == 0, p.zlz, ctx.dc[2], true);
  std::vector<double> h;
  auto put = [&](const std::vector<double> &v) { size_t o = h.size(); h.insert(h.end(), v.begin(), v.end()); return o; };
  const size_t o_ax = put(TX.a), o_bx = put(TX.b), o_ay = put(TY.a), o_by = put(TY.b), o_az = put(TZ.a), o_bz = put(TZ.b);
  const size_t o_xk = put(TX.k2), o_yk = put(TY.k2), o_zk = put(TZ.k2), o_tx = put(TX.tf), o_ty = put(TY.tf), o_tz = put(TZ.tf);
  {
    bool same = p.bcx == 0 && p.bcy == 0 && p.bcz == 0;
    for (size_t q = 0; same && q + 1 < TZ.k2.size(); q += 2) same = TZ.k2[q] == TZ.k2[q + 1] && TZ.tf[q] == TZ.tf[q + 1];
    if (const char *e = getenv("X3D_SPEC_REAL")) same = same && atoi(e) != 0;
    P->spec_real = same;
  }
  P->tables.reserve(h.size() * sizeof(double));
  X3D_CUDA(cudaMemcpyAsync(P->tables.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  double *base = static_cast<double *>(P->tables.p);
  P->d_ax = base + o_ax; P->d_bx = base + o_bx; P->d_ay = base + o_ay; P->d_by = base + o_by; P->d_az = base + o_az; P->d_bz = base + o_bz;
  P->d_xk2 = base + o_xk; P->d_yk2 = base + o_yk; P->d_zk2 = base + o_zk; P->d_tx = base + o_tx; P->d_ty = base + o_ty; P->d_tz = base + o_tz;
  {  // reorder index maps
    std::vector<int> hm;
    size_t off[2][3];
    const int nn[3] = {nx, ny, nz};
    const int bcs[3] = {p.bcx, p.bcy, p.bcz};
    for (int bw = 0; bw < 2; ++bw)
      for (int a = 0; a < 3; ++a) {
        std::vector<int> m = reorder_map(nn[a], bcs[a] != 0, bw == 1);
        off[bw][a] = hm.size();
        hm.insert(hm.end(), m.begin(), m.end());
      }
    const size_t o_idx = hm.size();
    for (int q = 0; q < std::max(nx, std::max(ny, nz)); ++q) hm.push_back(q);
    P->maps.reserve(hm.size() * sizeof(int));
    X3D_CUDA(cudaMemcpyAsync(P->maps.p, hm.data(), hm.size() * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
    X3D_CUDA(cudaStreamSynchronize(ctx.stream));
    for (int bw = 0; bw < 2; ++bw)
      for (int a = 0; a < 3; ++a) P->d_map[bw][a] = static_cast<int *>(P->maps.p) + off[bw][a];
    P->d_idx = static_cast<int *>(P->maps.p) + o_idx;
  }
  // decomposition: one rank, or slabs (p_row = 1) when x3d_decomp_init was called with several ranks
  P->nranks = 1;
  P->nyl = ny; P->nzl = nz; P->nzhl = nzh; P->k0 = 0;
  if (ctx.decomp) {
    int pr, pc, rk, nr;
    decomp_shape(ctx, &pr, &pc, &rk, &nr);
    if (nr > 1) {
      if (pr != 1) throw Error("x3d_poisson_init: the distributed solver uses slabs (p_row = 1)");
      P->nranks = nr;
      P->id_ph = decomp_info_init(ctx, nx, ny, nz);
      P->id_sp = decomp_info_init(ctx, nx, ny, nzh);
      x3d_decomp_info ph{}, sp{};
      decomp_info_get(ctx, P->id_ph, &ph);
      decomp_info_get(ctx, P->id_sp, &sp);
      P->nyl = ph.zsz[1]; P->nzl = ph.ysz[2];
      P->nzhl = sp.ysz[2]; P->k0 = sp.yst[2] - 1;
    }
  }
  // FFT plans (local extents)
  const int nyl = P->nyl, nzhl = P->nzhl;
  int nzv[1] = {nz};
  int inembed[1] = {nz}, onembed[1] = {nzh};
  const int nxyl = nx * nyl;
  size_t ws = 0, wmax = 0;
  {
    const char *e = getenv("X3D_FFT");
    // nx, ny, nz are the extents of the Poisson mesh (nxm ...: 64^3 for the 65^3 free-slip TGV, 256 x 128 x 128 for the channel)
    P->own_fft = !(e && atoi(e) == 0) && fft_real_ok(nz) && fft_complex_ok(ny) && fft_complex_ok(nx) && !getenv("X3D_FFT_CHUNK");
    P->own_spec = P->own_fft && !(p.bcx || p.bcy || p.bcz) && P->spec_real;
  }
  X3D_CUFFT(cufftCreate(&P->plan_r2c)); X3D_CUFFT(cufftCreate(&P->plan_c2r)); X3D_CUFFT(cufftCreate(&P->plan_xy));
  P->plans = true;
  X3D_CUFFT(cufftSetAutoAllocation(P->plan_r2c, 0)); X3D_CUFFT(cufftSetAutoAllocation(P->plan_c2r, 0)); X3D_CUFFT(cufftSetAutoAllocation(P->plan_xy, 0));
  if (nxyl > 0 && !P->own_fft) {
    X3D_CUFFT(cufftMakePlanMany(P->plan_r2c, 1, nzv, inembed, nxyl, 1, onembed, nxyl, 1, CUFFT_D2Z, nxyl, &ws)); wmax = std::max(wmax, ws);
    X3D_CUFFT(cufftMakePlanMany(P->plan_c2r, 1, nzv, onembed, nxyl, 1, inembed, nxyl, 1, CUFFT_Z2D, nxyl, &ws)); wmax = std::max(wmax, ws);
  }
  int nyx[2] = {ny, nx};
  if (nzhl > 0 && !P->own_fft) { X3D_CUFFT(cufftMakePlanMany(P->plan_xy, 2, nyx, nullptr, 1, nx * ny, nullptr, 1, nx * ny, CUFFT_Z2Z, nzhl, &ws)); wmax = std::max(wmax, ws); }
  if (!(p.bcx || p.bcy || p.bcz) && P->spec_real && nzhl > 0) {
    // Off by default: measured at 512^3 (profiles/r2g_fft_chunk_sweep.txt) the three whole-array passes take 1.83 ms, the
    // chunked form 2.0-4.7 ms for chunks of 64 ... 2 planes -- cuFFT's 2-D transforms lose more on small batches than the
    // L2 residency saves.  X3D_FFT_CHUNK=k turns it on.
    int ch = 0;
    const long long plane_bytes = static_cast<long long>(nx) * ny * 16;
    if (const char *e = getenv("X3D_FFT_CHUNK")) ch = atoi(e);
    if (ch > 0 && ch < nzhl && plane_bytes <= (48ll << 20)) {
      P->fft_chunk = ch;
      X3D_CUFFT(cufftCreate(&P->plan_xy_chunk));
      X3D_CUFFT(cufftSetAutoAllocation(P->plan_xy_chunk, 0));
      X3D_CUFFT(cufftMakePlanMany(P->plan_xy_chunk, 2, nyx, nullptr, 1, nx * ny, nullptr, 1, nx * ny, CUFFT_Z2Z, ch, &ws)); wmax = std::max(wmax, ws);
      if (nzhl % ch) {
        X3D_CUFFT(cufftCreate(&P->plan_xy_rem));
        X3D_CUFFT(cufftSetAutoAllocation(P->plan_xy_rem, 0));
        X3D_CUFFT(cufftMakePlanMany(P->plan_xy_rem, 2, nyx, nullptr, 1, nx * ny, nullptr, 1, nx * ny, CUFFT_Z2Z, nzhl % ch, &ws)); wmax = std::max(wmax, ws);
      }
    }
  }
  P->fftwork.reserve(wmax ? wmax : 16);
  for (cufftHandle h : {P->plan_xy_chunk, P->plan_xy_rem})
    if (h) { X3D_CUFFT(cufftSetWorkArea(h, P->fftwork.p)); X3D_CUFFT(cufftSetStream(h, ctx.stream)); }
  if (!P->own_fft) {
    X3D_CUFFT(cufftSetWorkArea(P->plan_r2c, P->fftwork.p)); X3D_CUFFT(cufftSetWorkArea(P->plan_c2r, P->fftwork.p)); X3D_CUFFT(cufftSetWorkArea(P->plan_xy, P->fftwork.p));
    X3D_CUFFT(cufftSetStream(P->plan_r2c, ctx.stream)); X3D_CUFFT(cufftSetStream(P->plan_c2r, ctx.stream)); X3D_CUFFT(cufftSetStream(P->plan_xy, ctx.stream));
  }
  const size_t nsp_y = static_cast<size_t>(nx) * ny * std::max(nzhl, 1);       // spectral y-pencil
  const size_t nsp_z = static_cast<size_t>(nx) * std::max(nyl, 1) * nzh;       // spectral z-pencil
  const size_t nr_z = static_cast<size_t>(nx) * std::max(nyl, 1) * nz;         // physical z-pencil
  const size_t nr_y = static_cast<size_t>(nx) * ny * std::max(P->nzl, 1);      // physical y-pencil
  P->cw.reserve(nsp_y * 16);
  if (P->nranks > 1) P->cwz.reserve(nsp_z * 16);
  if (p.bcx || p.bcy) P->cwb.reserve(nsp_y * 16);
  if (p.bcx || p.bcy || p.bcz) P->rwork.reserve(std::max(nr_z, nr_y) * 8);
  if (P->nranks > 1 && (p.bcx || p.bcy)) P->rwork2.reserve(std::max(nr_z, nr_y) * 8);
  if (P->istret != 0) penta_init(ctx, *P, TX, TY, TZ, p.zlz / static_cast<double>(nz), ctx.dc[2]);
  ctx.poisson = std::move(P);
}

static void reorder_launch(Ctx &ctx, const double *in, double *out, int d0, int d1, int d2, const int *mx, const int *my, const int *mz) {
  const long long n = static_cast<long long>(d0) * d1 * d2;
  if (n == 0) return;
  ProfScope ps(ctx, "poisson_reorder(k_reorder)");
  k_reorder<<<grid_for(n, ctx.sm_count), 256, 0, ctx.stream>>>(in, out, d0, d1, d2, mx, my, mz);
  X3D_CUDA(cudaGetLastError()); ctx.launches++;
}

// rhs: z-pencil of the pressure mesh (nx, nyl, nz), in place
void poisson_solve_device(Ctx &ctx, double *d_rhs) {
  auto *P = dynamic_cast<PoissonImpl *>(ctx.poisson.get());
  if (!P) throw Error("x3d_poisson: x3d_poisson_init has not been called");
  const int nx = P->nx, ny = P->ny, nz = P->nz, nzh = P->nzh, nyl = P->nyl, nzl = P->nzl, nzhl = P->nzhl;
  const bool multi = P->nranks > 1;
  const long long nsp = static_cast<long long>(nx) * ny * nzhl;
  SpecArgs a{nx, ny, nzhl, nz, P->k0, P->bcx, P->bcy, P->bcz, static_cast<double>(nx), static_cast<double>(ny), static_cast<double>(nz),
             1.0 / (static_cast<double>(nx) * static_cast<double>(ny) * static_cast<double>(nz)), P->d_ax, P->d_bx, P->d_ay, P->d_by, P->d_az, P->d_bz, P->d_xk2, P->d_yk2, P->d_zk2, P->d_tx, P->d_ty, P->d_tz};
  double2 *cw = static_cast<double2 *>(P->cw.p), *cwb = static_cast<double2 *>(P->cwb.p);
  double2 *cwz = multi ? static_cast<double2 *>(P->cwz.p) : cw;   // spectral z-pencil (aliases the y-pencil on one rank)
  double *rw = static_cast<double *>(P->rwork.p), *rw2 = static_cast<double *>(P->rwork2.p);
  const int gs = grid_for(nsp, ctx.sm_count);
  const bool any = P->bcx || P->bcy || P->bcz;
  // identity maps are the d_map entries of inactive axes; local extents along split axes use a prefix of them
  const int *ident_y = P->d_map[0][1], *ident_z = P->d_map[0][2];
  double *fft_in = d_rhs;
  if (any && !multi) {
    reorder_launch(ctx, d_rhs, rw, nx, ny, nz, P->d_map[0][0], P->d_map[0][1], P->d_map[0][2]);
    fft_in = rw;
  } else if (any) {
    // z is complete in the z-pencil; x and y are complete in the y-pencil (src/poisson.f90:1047-1101)
    const double *cur = d_rhs;
    if (P->bcz) { reorder_launch(ctx, cur, rw, nx, nyl, nz, P->d_idx, P->d_idx, P->d_map[0][2]); cur = rw; }
    if (P->bcx || P->bcy) {
      transpose_device(ctx, 2, cur, rw2, P->id_ph, 1);                                   // z -> y
      reorder_launch(ctx, rw2, rw, nx, ny, nzl, P->d_map[0][0], P->d_map[0][1], P->d_idx);
      transpose_device(ctx, 1, rw, rw2, P->id_ph, 1);                                    // y -> z
      cur = rw2;
    }
    fft_in = const_cast<double *>(cur);
    (void)ident_y; (void)ident_z;
  }
  if (P->own_spec) {
    // poisson_000 with the hand-written passes: z r2c | (transpose) | y forward | x forward + spectral factor + x inverse |
    // y inverse | (transpose) | z c2r -- the spectral array crosses HBM five times instead of seven
    const long long lanes_z = static_cast<long long>(nx) * nyl;   // (this block returns)
    if (lanes_z > 0) {
      ProfScope ps(ctx, "fft_z_r2c(k_fft_z_r2c)");
      fft_z_r2c(ctx, fft_in, cwz, nz, lanes_z, lanes_z);
    }
    if (multi) transpose_device(ctx, 2, reinterpret_cast<double *>(cwz), reinterpret_cast<double *>(cw), P->id_sp, 2);  // z -> y
    if (nzhl > 0) {
      FftSpec sp{ny, P->k0, -a.inv_norm, EPS, a.ax, a.bx, a.ay, a.by, a.az, a.bz, a.xk2, a.yk2, a.zk2, a.tx, a.ty, a.tz};
      { ProfScope ps(ctx, "fft_y(k_fft_strided)"); fft_strided(ctx, cw, ny, nx, static_cast<long long>(nx) * ny, nx, nzhl, false); }
      { ProfScope ps(ctx, "fft_x_fwd+spectral+fft_x_inv(k_fft_x_spec)"); fft_x_spec(ctx, cw, nx, static_cast<long long>(ny) * nzhl, &sp, 0); }
      { ProfScope ps(ctx, "fft_y(k_fft_strided)"); fft_strided(ctx, cw, ny, nx, static_cast<long long>(nx) * ny, nx, nzhl, true); }
    }
    if (multi) transpose_device(ctx, 1, reinterpret_cast<double *>(cw), reinterpret_cast<double *>(cwz), P->id_sp, 2);  // y -> z
    if (lanes_z > 0) {
      ProfScope ps(ctx, "fft_z_c2r(k_fft_z_c2r)");
      fft_z_c2r(ctx, cwz, d_rhs, nz, lanes_z, lanes_z);
    }
    return;
  }
  const long long lanes_z = static_cast<long long>(nx) * nyl;
  // the x-y transforms of the variants whose spectral step is not the single real factor: hand-written y and x passes, or the 2-D plan
  auto fft_xy = [&](bool inverse) {
    if (nzhl <= 0) return;
    if (P->own_fft) {
      ProfScope ps(ctx, inverse ? "fft_x_inv+fft_y_inv(k_fft_x_spec, k_fft_strided)" : "fft_y_fwd+fft_x_fwd(k_fft_strided, k_fft_x_spec)");
      if (!inverse) fft_strided(ctx, cw, ny, nx, static_cast<long long>(nx) * ny, nx, nzhl, false);
      fft_x_spec(ctx, cw, nx, static_cast<long long>(ny) * nzhl, nullptr, inverse ? 1 : 0);
      if (inverse) fft_strided(ctx, cw, ny, nx, static_cast<long long>(nx) * ny, nx, nzhl, true);
    } else {
      ProfScope ps(ctx, "fft_xy_c2c(cuFFT)");
      X3D_CUFFT(cufftExecZ2Z(P->plan_xy, reinterpret_cast<cufftDoubleComplex *>(cw), reinterpret_cast<cufftDoubleComplex *>(cw), inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
    }
  };
  if (lanes_z > 0 && P->own_fft) {
    ProfScope ps(ctx, "fft_z_r2c(k_fft_z_r2c)");
    fft_z_r2c(ctx, fft_in, cwz, nz, lanes_z, lanes_z);
  } else if (lanes_z > 0) {
    ProfScope ps(ctx, "fft_z_r2c(cuFFT)");
    X3D_CUFFT(cufftExecD2Z(P->plan_r2c, fft_in, reinterpret_cast<cufftDoubleComplex *>(cwz)));
  }
  if (multi) transpose_device(ctx, 2, reinterpret_cast<double *>(cwz), reinterpret_cast<double *>(cw), P->id_sp, 2);  // z -> y
  const bool chunked = !any && P->fft_chunk > 0 && nzhl > 0;
  if (chunked) {
    // forward x-y FFT, spectral factor, inverse x-y FFT, one L2-resident chunk of planes after the other
    ProfScope ps(ctx, "fft_xy_fwd+spectral+fft_xy_inv(cuFFT + k_spec, L2-resident plane chunks)");
    const long long plane = static_cast<long long>(nx) * ny;
    for (int k0 = 0; k0 < nzhl; k0 += P->fft_chunk) {
      const int cnt = std::min(P->fft_chunk, nzhl - k0);
      cufftHandle h = cnt == P->fft_chunk ? P->plan_xy_chunk : P->plan_xy_rem;
      cufftDoubleComplex *pc = reinterpret_cast<cufftDoubleComplex *>(cw + k0 * plane);
      X3D_CUFFT(cufftExecZ2Z(h, pc, pc, CUFFT_FORWARD));
      k_spec_000s<<<std::min<long long>(static_cast<long long>(ny) * cnt, 16LL * ctx.sm_count), 256, 0, ctx.stream>>>(a, cw + k0 * plane, k0, cnt);
      X3D_CUDA(cudaGetLastError()); ctx.launches++;
      X3D_CUFFT(cufftExecZ2Z(h, pc, pc, CUFFT_INVERSE));
    }
  }
  if (!chunked) fft_xy(false);
  auto stage = [&](unsigned mode, const double2 *in, double2 *out) {
    if (nsp == 0) return;
    ProfScope ps(ctx, "poisson_spectral(k_spec)");
    k_spec_stage<<<gs, 256, 0, ctx.stream>>>(a, mode, in, out);
    X3D_CUDA(cudaGetLastError()); ctx.launches++;
  };
  if (chunked) {
    // done above
  } else if (!any) {
    if (nsp > 0) {
      ProfScope ps(ctx, "poisson_spectral(k_spec)");
      if (P->spec_real) k_spec_000s<<<std::min<long long>(static_cast<long long>(ny) * nzhl, 16LL * ctx.sm_count), 256, 0, ctx.stream>>>(a, cw, 0, nzhl);
      else k_spec_000<<<gs, 256, 0, ctx.stream>>>(a, cw);
      X3D_CUDA(cudaGetLastError()); ctx.launches++;
    }
  } else if (P->bcx == 1 && P->bcy == 0) {  // poisson_100, :472-635
    stage(S_NORM | S_ROTZ_F | S_ROTY_F | S_POSTX | S_DIVIDE, cw, cwb);
    stage(S_PREX | S_ROTY_B | S_ROTZ_B, cwb, cw);
  } else if (P->bcx == 0 && P->bcy == 1) {  // poisson_010, :724-991
    if (P->istret == 0) {
      stage(S_NORM | S_ROTZ_F | S_ROTX_F | S_POSTY | S_DIVIDE | S_ZERO010, cw, cwb);
    } else {  // :822-893: pentadiagonal systems in y instead of the division
      stage(S_NORM | S_ROTZ_F | S_ROTX_F | S_POSTY, cw, cwb);
      penta_solve(ctx, *P, cwb, true);
    }
    stage(S_PREY | S_ROTX_B | S_ROTZ_B, cwb, cw);
  } else {  // poisson_11x, :1118-1407
    stage(S_NORM | S_ROTZ_F | S_POSTY, cw, cwb);
    if (P->istret == 0) {
      stage(S_POSTX | S_DIVIDE, cwb, cw);
    } else {  // :1232-1330
      stage(S_POSTX, cwb, cw);
      penta_solve(ctx, *P, cw, false);
    }
    stage(S_PREX, cw, cwb);
    stage(S_PREY | S_ROTZ_B, cwb, cw);
  }
  if (!chunked) fft_xy(true);
  if (multi) transpose_device(ctx, 1, reinterpret_cast<double *>(cw), reinterpret_cast<double *>(cwz), P->id_sp, 2);  // y -> z
  double *fft_out = any ? rw : d_rhs;
  if (lanes_z > 0 && P->own_fft) {
    ProfScope ps(ctx, "fft_z_c2r(k_fft_z_c2r)");
    fft_z_c2r(ctx, cwz, fft_out, nz, lanes_z, lanes_z);
  } else if (lanes_z > 0) {
    ProfScope ps(ctx, "fft_z_c2r(cuFFT)");
    X3D_CUFFT(cufftExecZ2D(P->plan_c2r, reinterpret_cast<cufftDoubleComplex *>(cwz), fft_out));
  }
  if (any && !multi) {
    reorder_launch(ctx, rw, d_rhs, nx, ny, nz, P->d_map[1][0], P->d_map[1][1], P->d_map[1][2]);
  } else if (any) {  // src/poisson.f90:1422-1460
    double *cur = rw;
    if (P->bcz) { reorder_launch(ctx, cur, (P->bcx || P->bcy) ? rw2 : d_rhs, nx, nyl, nz, P->d_idx, P->d_idx, P->d_map[1][2]); cur = rw2; }
    if (P->bcx || P->bcy) {
      double *o1 = (cur == rw) ? rw2 : rw;
      transpose_device(ctx, 2, cur, o1, P->id_ph, 1);
      reorder_launch(ctx, o1, cur, nx, ny, nzl, P->d_map[1][0], P->d_map[1][1], P->d_idx);
      transpose_device(ctx, 1, cur, d_rhs, P->id_ph, 1);
    }
  }
}

// local extents of the z-pencil the solver works on
void poisson_dims(Ctx &ctx, int d[3]) {
  auto *P = dynamic_cast<PoissonImpl *>(ctx.poisson.get());
  if (!P) throw Error("x3d_poisson: not initialised");
  d[0] = P->nx; d[1] = P->nyl; d[2] = P->nz;
}

}  // namespace x3d
