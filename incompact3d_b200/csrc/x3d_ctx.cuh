// x3d_ctx.cuh -- the library context (opaque x3d_ctx of the C ABI).
#pragma once
#include "x3d_common.cuh"

namespace x3d {

// process-wide registry of the library's own device allocations (base -> size): lets the peer-to-peer
// transposes find the allocation a pointer belongs to (cudaIpcGetMemHandle needs the base)
void register_alloc(void *p, size_t n);
void unregister_alloc(void *p);
bool find_alloc(const void *q, void **base, size_t *size, unsigned long long *gen = nullptr);

// one item of a ring exchange between neighbouring z slabs (x3d_decomp.cu: ring_exchange)
struct RingCopy {
  const void *src;   // this rank's data
  void *dst;         // base pointer of a library-owned buffer allocated the same way on every rank
  size_t dst_off;    // byte offset inside the neighbour's buffer
  size_t bytes;
  int dir;           // +1: to the next rank of the ring, -1: to the previous one
};

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  void reserve(size_t n) {
    if (n <= bytes) return;
    release();
    // whole 2 MiB pages: the allocation is then never carved out of a driver pool page shared with other
    // allocations, so its CUDA IPC handle maps exactly this buffer in a peer process (peer-to-peer transposes)
    const size_t page = size_t(2) << 20;
    n = (n + page - 1) / page * page;
    X3D_CUDA(cudaMalloc(&p, n));
    bytes = n;
    register_alloc(p, n);
  }
  void release() {
    if (p) { unregister_alloc(p); cudaFree(p); }
    p = nullptr;
    bytes = 0;
  }
  ~DevBuf() { release(); }
};

struct ProfRec {
  const char *name;
  cudaEvent_t a, b;
};

struct PoissonState;
struct DecompState;
struct SolverState;

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  long long launches = 0;
  // kernel variants (tuning knobs; X3D_STRIDED_VARIANT / X3D_CONTIG_VARIANT override)
  int strided_variant = 6, contig_variant = 1;
  bool contig_compress = true;   // X3D_CONTIG_COMPRESS=0: long x lines keep the full coefficient table in shared memory
  // module state made explicit (x3d_set_deriv_coeffs / x3d_set_filter_coeffs / x3d_set_flags)
  x3d_deriv_coeffs dc[3]{};
  x3d_filter_coeffs fc[3]{};
  bool have_dc[3] = {false, false, false}, have_fc[3] = {false, false, false};
  int iibm = 0, istret = 0, iimplicit = 0;
  bool ncl[3] = {true, true, true};
  // stretched-mesh metrics of the host's stretching() (x3d_set_stretching); empty when istret == 0
  std::vector<double> st_yp, st_ypi, st_ppy, st_pp2y, st_pp4y, st_ppyi, st_pp2yi, st_pp4yi;
  // immersed-boundary geometry per direction (x3d_set_ibm_geometry); device copies
  struct IbmAxis {
    bool set = false;
    int nobjmax = 0, npif = 2, izap = 1, na = 0, nb = 0, ncoords = 0;
    double d = 0.0, len = 0.0;
    DevBuf nobj, xi, xf, nipif, nfpif, coords;
    bool analytic = false;        // iibm = 3 with ianal /= 0: analytic wall positions
    DevBuf ana_i, ana_f;
  };
  IbmAxis ibm[3];
  // optional per-launch CUDA-event timing (x3d_profile_begin / x3d_profile_end)
  bool profiling = false;
  std::vector<ProfRec> prof;
  // caches
  std::map<uint64_t, std::unique_ptr<TriTable>> tri_cache;
  // staging for host-pointer (drop-in) calls
  DevBuf stage_in, stage_out;
  // sub-systems
  std::unique_ptr<DecompState> decomp;
  std::unique_ptr<PoissonState> poisson;
  std::unique_ptr<SolverState> solver;
  Ctx();
  ~Ctx();
};

// RAII scope that brackets the launches of one kernel class with CUDA events when profiling is on
struct ProfScope {
  Ctx &ctx;
  bool on;
  ProfRec r{};
  ProfScope(Ctx &c, const char *name) : ctx(c), on(c.profiling) {
    if (!on) return;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, ctx.stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.b, ctx.stream);
    ctx.prof.push_back(r);
  }
};

// device pointer classification: true if p is device (or managed) memory
bool is_device_ptr(const void *p);

// LU arrays (host) -> cached device table for chunk length L
const TriTable &get_tri(Ctx &ctx, const double *f, const double *s, const double *w, int n, int L, bool periodic,
                        double alpha, const double *post);

// launchers (device pointers)
void launch_line_op(Ctx &ctx, const DevOp &op, const OpCall &call, const double *d_u, double *d_t);
// full operator call with host-or-device pointers
void run_op(Ctx &ctx, OpCall &call, const double *u, double *t);
// lagpolx/y/z on a device array (nx,ny,nz), in place
void lagpol_device(Ctx &ctx, int axis, double *d_u, int nx, int ny, int nz);
// cubsplx/y/z on a device array (nx,ny,nz), in place; lind = the value imposed on the walls
void cubspl_device(Ctx &ctx, int axis, double *d_u, int nx, int ny, int nz, double lind);

}  // namespace x3d
