// x3d_ops.cu -- host side of the compact operators: context, pointer classification,
// staging of host fields (drop-in mode), geometry and kernel dispatch.
#include <cstring>
#include <mutex>
#include "x3d_ctx.cuh"
#include "x3d_ops_inst.cuh"
#include "x3d_state.cuh"

namespace x3d {

void launch_kind_D1(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);
void launch_kind_D2(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);
void launch_kind_FIL(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);
void launch_kind_DVP(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);
void launch_kind_IVP(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);
void launch_kind_DPV(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);
void launch_kind_IPV(Ctx &, const DevOp &, const LineGeom &, const TriTable &, const double *, double *);

Ctx::Ctx() {}
void stag_release(Ctx *ctx);
void fft_release(Ctx *ctx);
Ctx::~Ctx() {
  stag_release(this);
  fft_release(this);
  tri_cache.clear();
  if (stream) cudaStreamDestroy(stream);
}

// ---- registry of library-owned device allocations -------------------------------------------------
static std::mutex g_alloc_mu;
struct AllocRec { size_t size; unsigned long long gen; };
static std::map<uintptr_t, AllocRec> g_allocs;
static unsigned long long g_alloc_gen = 0;   // every allocation gets its own number: a reused address is a new buffer
void register_alloc(void *p, size_t n) {
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  g_allocs[reinterpret_cast<uintptr_t>(p)] = AllocRec{n, ++g_alloc_gen};
}
void unregister_alloc(void *p) {
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  g_allocs.erase(reinterpret_cast<uintptr_t>(p));
}
bool find_alloc(const void *q, void **base, size_t *size, unsigned long long *gen) {
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  const uintptr_t a = reinterpret_cast<uintptr_t>(q);
  auto it = g_allocs.upper_bound(a);
  if (it == g_allocs.begin()) return false;
  --it;
  if (a >= it->first + it->second.size) return false;
  *base = reinterpret_cast<void *>(it->first);
  *size = it->second.size;
  if (gen) *gen = it->second.gen;
  return true;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a{};
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int pick_L_strided(int n, int variant) {
  if (n > 1024) return n <= 1536 ? 48 : 64;                      // long lines (e.g. 1536^3 on 8 GPUs): 32 chunks of 48 / 64 rows
  if (variant >= 4) return n <= 256 ? 8 : (n <= 512 ? 16 : 32);  // TMA tile kernels: up to 32 chunks per line
  if (n <= 128) return 8;
  if (n <= 256) return 16;
  if (n <= 512 && variant >= 2) return 16;
  return 32;
}
int pick_L_contig(int n) {
  const int cand[7] = {5, 9, 17, 25, 33, 49, 65};
  for (int L : cand)
    if (32 * L >= n) return L;
  return -1;
}

__global__ void k_scale_lines(double *t, const double *post, long long n_lane, long long n_line, long long n_outer) {
  // t[(o*n_line + q)*n_lane + i] *= post[q]
  const long long tot = n_lane * n_line * n_outer;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long q = (idx / n_lane) % n_line;
    t[idx] *= post[q];
  }
}

static LineGeom make_geom(const OpCall &call, int n_in, int n_out) {
  LineGeom g{};
  g.axis = call.axis;
  const long long d0 = call.dims_in[0], d1 = call.dims_in[1], d2 = call.dims_in[2];
  if (call.axis == 0) {
    g.nlines = d1 * d2;
  } else if (call.axis == 1) {
    g.n1 = static_cast<int>(d0);
    g.nouter = d2;
    g.sin = g.sout = d0;
    g.oin = d0 * n_in;
    g.oout = d0 * n_out;
  } else {
    if (d0 * d1 > 2147483647LL) throw Error("z-pencil plane too large");
    g.n1 = static_cast<int>(d0 * d1);
    g.nouter = 1;
    g.sin = g.sout = d0 * d1;
    g.oin = g.oout = 0;
  }
  return g;
}

void launch_line_op(Ctx &ctx, const DevOp &op, const OpCall &call, const double *d_u, double *d_t) {
  const int n_in = op.n_in, n_out = op.n_out;
  LineGeom g = make_geom(call, n_in, n_out);
  int L = -1;
  if (call.axis != 0 && ctx.strided_variant >= 6) {  // warp-per-lane-pair TMA kernel when the call qualifies
    PairGeom pg;
    size_t smem;
    const int Lp = pick_L_contig(n_out);
    if (Lp > 0 && Lp <= 17 && pair_plan(op, g, Lp, d_u, d_t, pg, smem)) { L = Lp; g.pair = true; }
  }
  if (L < 0) L = call.axis == 0 ? pick_L_contig(n_out) : pick_L_strided(n_out, ctx.strided_variant);
  if (L < 0) throw Error("x-direction line too long for the warp-per-line kernel (n <= 2080)");
  const TriTable &T = get_tri(ctx, call.f, call.s, call.w, n_out, L, op.periodic != 0, op.alpha, call.post);
  // operators that accumulate into their destination move 24 B per point, not 16: they get their own class
  // the class names carry the kernel that runs: k_pair for y / z lines that qualify, the k_tile / k_strided fallbacks else
  static const char *names[2][2][3] = {
      {{"compact_x(k_contig)", "compact_y(k_strided)", "compact_z(k_strided)"},
       {"accumulate_x(k_contig, TMA reduce-add)", "accumulate_y(k_strided)", "accumulate_z(k_strided)"}},
      {{"compact_x(k_contig)", "compact_y(k_pair)", "compact_z(k_pair)"},
       {"accumulate_x(k_contig, TMA reduce-add)", "accumulate_y(k_pair, TMA reduce-add)", "accumulate_z(k_pair, TMA reduce-add)"}}};
  ProfScope ps(ctx, names[g.pair ? 1 : 0][op.store_mode != 0][call.axis]);
  switch (op.kind) {
    case D1: launch_kind_D1(ctx, op, g, T, d_u, d_t); break;
    case D2: launch_kind_D2(ctx, op, g, T, d_u, d_t); break;
    case FIL: launch_kind_FIL(ctx, op, g, T, d_u, d_t); break;
    case DVP: launch_kind_DVP(ctx, op, g, T, d_u, d_t); break;
    case IVP: launch_kind_IVP(ctx, op, g, T, d_u, d_t); break;
    case DPV: launch_kind_DPV(ctx, op, g, T, d_u, d_t); break;
    case IPV: launch_kind_IPV(ctx, op, g, T, d_u, d_t); break;
    default: throw Error("bad operator kind");
  }
}

void run_op(Ctx &ctx, OpCall &call, const double *u, double *t) {
  X3D_CUDA(cudaSetDevice(ctx.device));
  const int n_in = op_n_in(call), n_out = op_n_out(call);
  if (call.dims_in[call.axis] != n_in) throw Error("operator: line extent does not match the array shape");
  long long cnt_in = 1, cnt_out = 1;
  for (int d = 0; d < 3; ++d) {
    cnt_in *= call.dims_in[d];
    cnt_out *= (d == call.axis ? n_out : call.dims_in[d]);
  }
  if (cnt_in == 0 || cnt_out == 0) return;
  const bool u_dev = is_device_ptr(u), t_dev = is_device_ptr(t);
  const double *d_u = u;
  double *d_t = t;
  if (!u_dev) {
    ctx.stage_in.reserve(cnt_in * sizeof(double));
    X3D_CUDA(cudaMemcpyAsync(ctx.stage_in.p, u, cnt_in * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    d_u = static_cast<const double *>(ctx.stage_in.p);
  }
  bool t_needs_upload = false;
  if (!t_dev) {
    ctx.stage_out.reserve(cnt_out * sizeof(double));
    d_t = static_cast<double *>(ctx.stage_out.p);
  }
  // n == 1 shortcuts of the z operators (derive.f90:874,2939,4937,5121,5305,5445; filters.f90:1008)
  bool done = false;
  if (call.n == 1 && call.axis == 2) {
    const bool interp = (call.kind == IVP || call.kind == IPV);
    if (interp) {
      if (call.nm == 1) { X3D_CUDA(cudaMemcpyAsync(d_t, d_u, cnt_out * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream)); done = true; }
    } else if (call.kind == FIL) {
      if (call.ncl1 == 0 && call.ncln == 0) { X3D_CUDA(cudaMemcpyAsync(d_t, d_u, cnt_out * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream)); done = true; }
    } else {
      X3D_CUDA(cudaMemsetAsync(d_t, 0, cnt_out * sizeof(double), ctx.stream));
      done = true;
    }
  }
  // iibm = 2 / 3: the collocated derivatives and the filters first rebuild their input inside the bodies, in place
  // (derive.f90:23-24, filters.f90:235-236)
  bool u_modified = false;
  if (!done && (ctx.iibm == 2 || ctx.iibm == 3) && (call.kind == D1 || call.kind == D2 || call.kind == FIL)) {
    if (ctx.iibm == 2) lagpol_device(ctx, call.axis, const_cast<double *>(d_u), call.dims_in[0], call.dims_in[1], call.dims_in[2]);
    else cubspl_device(ctx, call.axis, const_cast<double *>(d_u), call.dims_in[0], call.dims_in[1], call.dims_in[2], call.lind);
    u_modified = true;
  }
  if (!done) {
    DevOp op;
    build_devop(ctx, call, op);
    if (op.untouched) {
      // the reference skips RHS and solve (t keeps its content) but still applies the trailing
      // stretching multiply (derive.f90:4572-4580)
      if (call.post && !call.rhs_only) {
        if (!t_dev) { X3D_CUDA(cudaMemcpyAsync(d_t, t, cnt_out * sizeof(double), cudaMemcpyHostToDevice, ctx.stream)); }
        DevBuf postbuf;
        postbuf.reserve(n_out * sizeof(double));
        X3D_CUDA(cudaMemcpyAsync(postbuf.p, call.post, n_out * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
        long long n_lane = 1, n_outer = 1;
        for (int d = 0; d < call.axis; ++d) n_lane *= call.dims_in[d];
        for (int d = call.axis + 1; d < 3; ++d) n_outer *= call.dims_in[d];
        k_scale_lines<<<ctx.sm_count * 8, 256, 0, ctx.stream>>>(d_t, static_cast<const double *>(postbuf.p), n_lane, n_out, n_outer);
        X3D_CUDA(cudaGetLastError());
        ctx.launches++;
        X3D_CUDA(cudaStreamSynchronize(ctx.stream));
        t_needs_upload = true;
      }
      // else: nothing to do for t (it stays untouched); a rebuilt input still goes back to the caller below
    } else {
      launch_line_op(ctx, op, call, d_u, d_t);
      t_needs_upload = true;
    }
  } else {
    t_needs_upload = true;
  }
  if (u_modified && !u_dev)  // the reference modifies the caller's array: hand the rebuilt input back
    X3D_CUDA(cudaMemcpyAsync(const_cast<double *>(u), d_u, cnt_in * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (!t_dev && t_needs_upload) {
    X3D_CUDA(cudaMemcpyAsync(t, d_t, cnt_out * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  } else if (!u_dev) {
    X3D_CUDA(cudaStreamSynchronize(ctx.stream));  // staging buffer is reused by the next call
  }
}

}  // namespace x3d
