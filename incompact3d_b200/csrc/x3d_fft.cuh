// x3d_fft.cuh -- hand-written FFT passes of the periodic Poisson solve, host interface (see x3d_fft_kernels.cuh)
#pragma once
#include "x3d_ctx.cuh"

namespace x3d {

struct FftSpec;
bool fft_complex_ok(int n);   // line lengths the kernels are instantiated for
bool fft_real_ok(int n);
// complex lines of n points, in place: element (lane, row, outer) at lane + row * stride + outer * ostride
void fft_strided(Ctx &ctx, double2 *data, int n, long long stride, long long ostride, int lanes, long long nouter, bool inverse);
// real lines of n points (row stride `plane`, `lanes` lines) <-> n/2+1 complex rows with the same row stride; unnormalised
void fft_z_r2c(Ctx &ctx, const double *in, double2 *out, int n, long long plane, long long lanes);
void fft_z_c2r(Ctx &ctx, const double2 *in, double *out, int n, long long plane, long long lanes);
// contiguous complex lines of n points, in place: sp != nullptr: forward transform, spectral factor of poisson_000, inverse
// transform; sp == nullptr: one forward (inverse_only = 0) or inverse transform
void fft_x_spec(Ctx &ctx, double2 *data, int n, long long nlines, const FftSpec *sp, int inverse_only);
void fft_release(Ctx *ctx);

}  // namespace x3d
