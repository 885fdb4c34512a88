// x3d_common.cuh -- shared declarations of the B200 hot-path library.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/x3d_b200.h"

namespace x3d {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define X3D_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      throw ::x3d::Error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" +        \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")");                  \
  } while (0)

// ---- compact line operator, device view ---------------------------------------
// Interior stencils (the association of the reference is kept where it is free):
//   D1   a(u[i+1]-u[i-1]) + b(u[i+2]-u[i-2])                       derive.f90:35-36
//   D2   sum_k c_k (u[i+k]-u[i]-u[i]+u[i-k]), k=1..4               derive.f90:1412-1419
//   FIL  c0 u[i] + sum_k c_k (u[i+k]+u[i-k]), k=1..3               filters.f90:253-255
//   DVP  a(u[i+1]-u[i]) + b(u[i+2]-u[i-1])                         derive.f90:3826-3827
//   IVP  sum_k c_k (u[i+k]+u[i-k+1]), k=1..4                       derive.f90:3947-3950
//   DPV  a(p[i]-p[i-1]) + b(p[i+1]-p[i-2])                         derive.f90:4068-4069
//   IPV  sum_k c_k (p[i+k-1]+p[i-k]), k=1..4                       derive.f90:4168-4171
enum Kind : int { D1 = 0, D2 = 1, FIL = 2, DVP = 3, IVP = 4, DPV = 5, IPV = 6 };

constexpr int HALO = 4;     // widest reach of any stencil
constexpr int NBROW = 4;    // explicit boundary rows kept per end
constexpr int NBCOL = 9;    // inputs an explicit row may touch
constexpr int TRI_W = 8;    // doubles per row of the packed tridiagonal table

// one row of the packed table: s, Pf | w, f*w | Pb, rs | post, (pad)
enum TriCol : int { T_S = 0, T_PF = 1, T_W = 2, T_FW = 3, T_PB = 4, T_RS = 5, T_POST = 6 };

struct DevOp {
  int kind;
  int n_in, n_out;
  int periodic;         // wrap ghosts + Sherman-Morrison correction
  int nb;               // explicit rows at each end (0 when periodic)
  int rhs_only;         // deryy with iimplicit>=1: no solve
  int has_post;
  int untouched;        // unsupported npaire: reference leaves t untouched (only post applied)
  int store_mode;       // 0: t = result | 1: t += result | 2: t -= result  (TMA reduce-add; fuses the elementwise
                        //    sums of divergence, navier.f90:325,339, and cor_vel, :242-244, into the operator)
  double c0, c[4];      // interior stencil coefficients
  double alpha;         // Sherman-Morrison alpha (alfai, alsai, alcai6, ...)
  double wstart[NBROW][NBCOL];  // row r = sum_q wstart[r][q] * u[q]
  double wend[NBROW][NBCOL];    // row n_out-NBROW+r = sum_q wend[r][q] * u[n_in-NBCOL+q]
};

// device-resident tables of one (LU arrays, chunk length) pair
struct TriTable {
  int n = 0, L = 0, nc = 0;
  double *d_rows = nullptr;   // [nc*L][TRI_W]
  double *d_scan = nullptr;   // [10][32] Kogge-Stone multipliers (x kernels), [2][nc] chunk products
  double *d_chunk = nullptr;  // [2][nc]: Af(c) forward chunk product, Ab(c) backward chunk product
  std::vector<double> h_rows, h_scan;  // host copies (compressed tables of the fused kernels are derived from them)
  // compressed rows for long lines (built on first use, x3d_tables.cu: compress_tri): the LU rows reach their
  // floating-point fixed point a few dozen rows from the boundaries, so chunks c_head .. nc-c_head-2 share one table.
  // c_head = 0: not built yet, -1: the table does not compress
  // d_rows_c is the shared-memory image of k_contig: 7 columns of (2 c_head + 2) L entries (c_head head chunks, one
  // generic chunk, c_head + 1 tail chunks), then the Sherman-Morrison column once more with its own head count c_head_rs
  // (it decays more slowly than the LU rows converge: 0.82 per row for the alpha = 0.49 interpolators)
  mutable double *d_rows_c = nullptr;
  mutable int c_head = 0, c_head_rs = 0;
  ~TriTable();
};
// builds T.d_rows_c; returns the number of head chunks, or -1
int compress_tri(const TriTable &T);

struct Ctx;

// host description of one reference operator call
struct OpCall {
  Kind kind;
  int axis;              // 0,1,2
  int ncl1, ncln;        // routine variant
  bool periodic;         // logical nclx/y/z for staggered operators
  int npaire;
  int n, nm;             // velocity nodes / pressure points along the line
  int dims_in[3];
  const double *f, *s, *w;   // host LU arrays (caller-owned)
  const double *post;        // host ppy/ppyi or nullptr
  bool rhs_only;
  double lind;               // the operators' last argument: wall value of the iibm = 3 pre-pass
};

void build_devop(const Ctx &ctx, const OpCall &call, DevOp &op);
int op_n_in(const OpCall &c);
int op_n_out(const OpCall &c);

// chunk length policy
int pick_L_strided(int n, int variant);
int pick_L_contig(int n);

}  // namespace x3d
