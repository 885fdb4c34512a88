// fft.cpp -- plain mixed-radix complex FFT for the oracle (TEST INFRASTRUCTURE).
//
// 2DECOMP&FFT v2.0.4 (external, un-vendored: cmake/decomp2d/downloadBuild2decomp.cmake.in:9-15)
// defines decomp_2d_fft_3d as the mathematical, unnormalised DFT with forward sign -1
// (DECOMP_2D_FFT_FORWARD = -1) applied axis by axis: physical-in-Z r2c along z, then c2c
// along y, then c2c along x; backward in the opposite order with sign +1 and a c2r last.
// Any exact DFT reproduces it to round-off; this is a recursive decimation-in-time for
// factors 2,3,5 with a direct O(p^2) butterfly for other primes.
#include <cmath>
#include <map>
#include <mutex>
#include "x3d_oracle.hpp"

namespace x3do {

namespace {
struct Plan {
  int n;
  std::vector<cplx> tw;  // tw[k] = exp(-2 pi i k / n)
};
std::map<int, Plan> g_plans;
std::mutex g_mu;

const Plan &plan(int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_plans.find(n);
  if (it != g_plans.end()) return it->second;
  Plan p;
  p.n = n;
  p.tw.resize(n);
  const double pi = std::acos(-1.0);
  for (int k = 0; k < n; ++k) p.tw[k] = cplx(std::cos(2.0 * pi * k / n), -std::sin(2.0 * pi * k / n));
  return g_plans.emplace(n, std::move(p)).first->second;
}

int smallest_factor(int n) {
  for (int p : {2, 3, 5}) if (n % p == 0) return p;
  for (int p = 7; p * p <= n; p += 2) if (n % p == 0) return p;
  return n;
}

// out[0..n) = DFT of in[0], in[s], in[2s], ...; N = top-level length (twiddle table), conj selects sign
void rec(const cplx *in, cplx *out, int n, std::ptrdiff_t s, const Plan &P, bool inverse, cplx *scratch) {
  if (n == 1) { out[0] = in[0]; return; }
  const int p = smallest_factor(n), m = n / p;
  for (int r = 0; r < p; ++r) rec(in + r * s, out + r * m, m, s * p, P, inverse, scratch);
  const int N = P.n, step = N / n;
  auto W = [&](long e) {  // exp(-+ 2 pi i e / n)
    const cplx w = P.tw[(e % n) * step];
    return inverse ? std::conj(w) : w;
  };
  cplx tmp[64];
  cplx *t = p <= 64 ? tmp : scratch;
  for (int k = 0; k < m; ++k) {
    for (int r = 0; r < p; ++r) t[r] = out[r * m + k] * W(static_cast<long>(r) * k);
    for (int q = 0; q < p; ++q) {
      cplx acc = t[0];
      for (int r = 1; r < p; ++r) acc += t[r] * W(static_cast<long>(r) * q * m);
      scratch[p + q] = acc;
    }
    for (int q = 0; q < p; ++q) out[q * m + k] = scratch[p + q];
  }
}
}  // namespace

void fft_line(cplx *x, int n, std::ptrdiff_t stride, int sign) {
  const Plan &P = plan(n);
  std::vector<cplx> in(n), out(n), scratch(2 * n + 2);
  for (int i = 0; i < n; ++i) in[i] = x[i * stride];
  rec(in.data(), out.data(), n, 1, P, sign > 0, scratch.data());
  for (int i = 0; i < n; ++i) x[i * stride] = out[i];
}

}  // namespace x3do
