// x3d_fft.cu -- host side of the hand-written FFT passes (x3d_fft_kernels.cuh): twiddle tables, launch.
#include <cmath>
#include <map>
#include <memory>
#include <vector>
#include "x3d_ctx.cuh"
#include "x3d_fft.cuh"
#include "x3d_fft_kernels.cuh"

namespace x3d {

namespace {
struct TwCache {
  std::map<int, double *> m;
  ~TwCache() { for (auto &kv : m) cudaFree(kv.second); }
};
std::map<Ctx *, std::unique_ptr<TwCache>> g_tw;   // per context: device memory belongs to its device

// stages = false: the single table W[m] = exp(-2 pi i m / n); stages = true: the per-stage tables of an n-point transform
// (FftTw<n>: TW_s[t-1][b] = exp(-2 pi i t (b mod NS) / (NS R)), stages 2 ..).  Angles in long double.
template <int N>
void build_stage_tables(std::vector<double> &h) {
  using F = FftRadix<N>;
  const long double pi = 3.141592653589793238462643383279502884L;
  const int R[4] = {F::R1, F::R2, F::R3, F::R4};
  int NS = R[0];
  for (int s = 1; s < 4 && R[s]; ++s) {
    for (int t = 1; t < R[s]; ++t)
      for (int b = 0; b < N / R[s]; ++b) {
        const long double a = -2.0L * pi * static_cast<long double>(t) * (b % NS) / (static_cast<long double>(NS) * R[s]);
        h.push_back(static_cast<double>(cosl(a)));
        h.push_back(static_cast<double>(sinl(a)));
      }
    NS *= R[s];
  }
}

const double2 *twiddles(Ctx &ctx, int n, bool stages) {
  auto &slot = g_tw[&ctx];
  if (!slot) slot = std::make_unique<TwCache>();
  const int key = stages ? -n : n;
  auto it = slot->m.find(key);
  if (it != slot->m.end()) return reinterpret_cast<const double2 *>(it->second);
  std::vector<double> h;
  if (stages) {
    switch (n) {
      case 64: build_stage_tables<64>(h); break;
      case 128: build_stage_tables<128>(h); break;
      case 256: build_stage_tables<256>(h); break;
      case 512: build_stage_tables<512>(h); break;
      case 1024: build_stage_tables<1024>(h); break;
      default: throw Error("hand-written FFT: unsupported length");
    }
  } else {
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int m = 0; m < n; ++m) {
      const long double a = -2.0L * pi * static_cast<long double>(m) / static_cast<long double>(n);
      h.push_back(static_cast<double>(cosl(a)));
      h.push_back(static_cast<double>(sinl(a)));
    }
  }
  double *d = nullptr;
  X3D_CUDA(cudaMalloc(&d, h.size() * sizeof(double)));
  X3D_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  slot->m[key] = d;
  return reinterpret_cast<const double2 *>(d);
}

template <class K, class... A>
void launch(Ctx &ctx, K kern, int threads, size_t smem, long long ntiles, int per_sm, A... args) {
  X3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  long long blocks = static_cast<long long>(ctx.sm_count) * per_sm;
  if (blocks > ntiles) blocks = ntiles;
  if (blocks < 1) return;
  kern<<<static_cast<unsigned>(blocks), threads, smem, ctx.stream>>>(args...);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}
}  // namespace

void fft_release(Ctx *ctx) { g_tw.erase(ctx); }

bool fft_complex_ok(int n) { return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; }
bool fft_real_ok(int n) { return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024; }

#define X3D_FFT_SWITCH(n, CALL)                                    \
  switch (n) {                                                     \
    case 64: CALL(64); break;                                      \
    case 128: CALL(128); break;                                    \
    case 256: CALL(256); break;                                    \
    case 512: CALL(512); break;                                    \
    case 1024: CALL(1024); break;                                  \
    default: throw Error("hand-written FFT: unsupported length");  \
  }

static size_t tw_bytes(int n) {   // sum over the stages after the first of (R - 1) n / R entries: below 3 n for up to four stages
  return static_cast<size_t>(3) * n * 16;
}

void fft_strided(Ctx &ctx, double2 *data, int n, long long stride, long long ostride, int lanes, long long nouter, bool inverse) {
  const double2 *W = twiddles(ctx, n, false);
  const long long ntiles = static_cast<long long>((lanes + 7) / 8) * nouter;
  const size_t smem = static_cast<size_t>(n) * 16 * 9;
#define CALL(N)                                                                                                             \
  if (inverse) launch(ctx, k_fft_strided<N, true>, N, smem, ntiles, N <= 512 ? 2 : 1, data, stride, ostride, lanes, ntiles, W); \
  else launch(ctx, k_fft_strided<N, false>, N, smem, ntiles, N <= 512 ? 2 : 1, data, stride, ostride, lanes, ntiles, W)
  X3D_FFT_SWITCH(n, CALL)
#undef CALL
}

void fft_z_r2c(Ctx &ctx, const double *in, double2 *out, int n, long long plane, long long lanes) {
  const double2 *W = twiddles(ctx, n, false);
  const int lx = 16;   // FftZLanes
  const long long ntiles = (lanes + lx - 1) / lx;
  const size_t smem = static_cast<size_t>(n) * 16 + static_cast<size_t>(n / 2 + 1) * lx * 16;
#define CALL(N) launch(ctx, k_fft_z_r2c<N>, FftZLanes<N>::LX * N / 16, smem, ntiles, FftZLanes<N>::MINB, in, out, plane, lanes, ntiles, W)
  X3D_FFT_SWITCH(n, CALL)
#undef CALL
}

void fft_z_c2r(Ctx &ctx, const double2 *in, double *out, int n, long long plane, long long lanes) {
  const double2 *W = twiddles(ctx, n, false);
  const int lx = 16;   // FftZLanes
  const long long ntiles = (lanes + lx - 1) / lx;
  const size_t smem = static_cast<size_t>(n) * 16 + static_cast<size_t>(n / 2 + 1) * lx * 16;
#define CALL(N) launch(ctx, k_fft_z_c2r<N>, FftZLanes<N>::LX * N / 16, smem, ntiles, FftZLanes<N>::MINB, in, out, plane, lanes, ntiles, W)
  X3D_FFT_SWITCH(n, CALL)
#undef CALL
}

void fft_x_spec(Ctx &ctx, double2 *data, int n, long long nlines, const FftSpec *sp, int inverse_only) {
  const double2 *W = twiddles(ctx, n, true);
  const int ll = n <= 512 ? 4 : 8;   // FftXLines
  const long long ntiles = (nlines + ll - 1) / ll;
  const size_t smem = tw_bytes(n) + static_cast<size_t>(ll) * (n + n / 8 + 1) * 16 + (sp ? static_cast<size_t>(3) * n * 8 : 0);
  FftSpec none{};
#define CALL(N)                                                                                                                              \
  if (sp) launch(ctx, k_fft_x_spec<N, true>, FftXLines<N>::LL * N / 8, smem, ntiles, FftXLines<N>::MINB, data, nlines, W, *sp, 0);            \
  else launch(ctx, k_fft_x_spec<N, false>, FftXLines<N>::LL * N / 8, smem, ntiles, FftXLines<N>::MINB, data, nlines, W, none, inverse_only)
  X3D_FFT_SWITCH(n, CALL)
#undef CALL
}

}  // namespace x3d
