// x3d_decomp.cu -- 2-D pencil decomposition bookkeeping and pencil transposes
// (2DECOMP&FFT v2.0.4, external: decomp_2d_init / decomp_info_init / transpose_*; call sites
// src/xcompact3d.f90:191-201, src/transeq.f90:163,236,318,437, src/poisson.f90:132-133).
//
// Distribution rule of 2DECOMP&FFT: an extent n over p ranks gives n/p points each, the LAST
// mod(n,p) ranks get one more.  The process grid is p_row x p_col; rank = row*p_col + col
// (MPI_CART row-major).  x-pencil (nx, ny/p_row, nz/p_col), y-pencil (nx/p_row, ny, nz/p_col),
// z-pencil (nx/p_row, ny/p_col, nz).  x<->y exchanges inside the p_row ranks that share a
// column index, y<->z inside the p_col ranks that share a row index.
//
// A transpose is  pack kernel -> all-to-all(v) -> unpack kernel.  The exchange is a grouped
// ncclSend/ncclRecv over NVLink/NVSwitch (NCCL resolved at run time from the libnccl already
// loaded in the process, so that the library has no link-time NCCL dependency).  One rank (or
// one rank per group) degenerates to a bit-exact device copy.
#include <dlfcn.h>
#include <nccl.h>
#include <array>
#include <cstdlib>
#include <cstring>
#include "x3d_state.cuh"

namespace x3d {

void distribute(int n, int p, std::vector<int> &st, std::vector<int> &sz) {
  st.assign(p, 0); sz.assign(p, 0);
  const int base = n / p, rem = n % p;
  int s = 0;
  for (int r = 0; r < p; ++r) {
    sz[r] = base + (r >= p - rem ? 1 : 0);
    st[r] = s;  // 0-based
    s += sz[r];
  }
}

// ---- NCCL, resolved lazily --------------------------------------------------------------------
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi &nccl() {
  static NcclApi api;
  if (api.h) return api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) { api.h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (api.h) break; }  // already in the process (torch)
  if (!api.h)
    for (const char *nm : names) { api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
  if (!api.h) throw Error(std::string("cannot load libnccl: ") + dlerror());
#define X3D_SYM(field, name)                                                          \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, name));               \
  if (!api.field) throw Error(std::string("libnccl lacks ") + name)
  X3D_SYM(GetUniqueId, "ncclGetUniqueId"); X3D_SYM(CommInitRank, "ncclCommInitRank"); X3D_SYM(CommSplit, "ncclCommSplit");
  X3D_SYM(CommDestroy, "ncclCommDestroy"); X3D_SYM(Send, "ncclSend"); X3D_SYM(Recv, "ncclRecv");
  X3D_SYM(AllReduce, "ncclAllReduce"); X3D_SYM(AllGather, "ncclAllGather"); X3D_SYM(GroupStart, "ncclGroupStart"); X3D_SYM(GroupEnd, "ncclGroupEnd");
  X3D_SYM(GetErrorString, "ncclGetErrorString");
#undef X3D_SYM
  return api;
}
#define X3D_NCCL(call)                                                                                 \
  do {                                                                                                 \
    ncclResult_t r_ = (call);                                                                          \
    if (r_ != ncclSuccess) throw ::x3d::Error(std::string(#call) + ": " + nccl().GetErrorString(r_));  \
  } while (0)

// ---- plans ------------------------------------------------------------------------------------------
struct SidePlan {       // how one pencil array maps onto the packed exchange buffer
  int dims[3];          // local pencil extents
  int axis;             // direction that is cut into per-peer blocks
  std::vector<int> blk_start, blk_size;     // per peer, along `axis` (local 0-based coordinates)
  std::vector<long long> disp, count;       // per peer, in elements
  int *d_meta = nullptr;                    // device: blk_of[dims[axis]] | blk_start[np] | blk_size[np]
  long long *d_disp = nullptr;
};
struct TransposePlan {
  int npeers = 1;
  SidePlan send, recv;
  bool built = false;
};

struct DecompImpl : DecompState {
  int nx = 0, ny = 0, nz = 0, p_row = 1, p_col = 1, rank = 0, nranks = 1;
  int row = 0, col = 0;
  bool have_nccl = false;
  ncclComm_t world = nullptr, comm_row = nullptr, comm_col = nullptr;  // comm_row: the p_row ranks of my column (x<->y)
  std::vector<x3d_decomp_info> infos;
  std::vector<std::array<int, 3>> gdims;
  std::vector<std::array<TransposePlan, 4>> plans;
  DevBuf sendbuf, recvbuf, meta;
  std::vector<void *> dev_allocs;
  // ---- peer-to-peer transposes (all ranks on one NVLink/NVSwitch node) ----------------------------
  // A transpose is ONE kernel that reads the local pencil and stores every element straight into the
  // destination pencil of the rank that owns it (peer memory mapped through CUDA IPC), bracketed by two
  // device-side flag barriers of the group; no pack / unpack pass and no staging buffer.
  struct Group {
    int np = 1, me = 0;
    ncclComm_t comm = nullptr;
    DevBuf flags;                         // my flag array [np] (peers store their epoch here)
    DevBuf d_peer_flags;                  // device array [np] of pointers to every member's flag array
    unsigned long long epoch = 0;
    long long timeout_cycles = 0;
    int *d_status = nullptr;
    // keyed by (local destination pointer, number of the allocation it lies in): an address reused by a later
    // allocation is a different buffer and is exchanged again (entries of released buffers are never hit)
    using Key = std::pair<const void *, unsigned long long>;
    std::map<Key, DevBuf> dst_cache;            // -> device array [np] of member pointers
    std::map<Key, bool> dst_bad;
    std::map<Key, std::vector<void *>> dst_host;   // the same member pointers on the host (block copies)
  };
  Group grp_row, grp_col;
  bool p2p = false;
  // how a peer-to-peer transpose moves whole-row blocks (y<->z): 0 = element kernel, 1 = copy engines
  // (cudaMemcpy2DAsync into the peers' pencils), 2 = k_p2p_blocks (vector copy kernel), -1 = by block size: copy
  // engines for blocks of 32 MiB and more (measured on 2 B200: 0.71 ms against 0.82 ms per transpose scope of the
  // 512^3 step, ~630 GB/s per direction), one kernel for all members below that, where the per-copy launch cost of
  // the engines would show.  X3D_P2P_MODE overrides.
  int p2p_mode = -1;
  // device-side barrier: give up after this many clock cycles and raise *status (mapped host memory) instead of
  // spinning for ever; the host turns the flag into an error at its next synchronisation point (decomp_check)
  long long barrier_timeout_cycles = 0;
  int *h_status = nullptr, *d_status = nullptr;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // bytes this rank sent to OTHER members (what crosses NVLink) and transposed fields, since x3d_decomp_init
  unsigned long long stat_remote_bytes = 0, stat_fields = 0;
  DevBuf xchg_send, xchg_recv;
  DevBuf self_src, self_dst, self_cnt;   // x3d_transpose_selftest
  std::map<std::array<unsigned char, 64>, void *> ipc_opened;
  ~DecompImpl() override {
    for (auto &kv : ipc_opened) cudaIpcCloseMemHandle(kv.second);
    if (h_status) cudaFreeHost(h_status);
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (void *p : dev_allocs) cudaFree(p);
    if (have_nccl) {
      if (comm_row) nccl().CommDestroy(comm_row);
      if (comm_col) nccl().CommDestroy(comm_col);
      if (world) nccl().CommDestroy(world);
    }
  }
};

static x3d_decomp_info make_info(int p_row, int p_col, int row, int col, int nx, int ny, int nz) {
  x3d_decomp_info I{};
  std::vector<int> st, sz;
  auto setp = [&](int *s, int *e, int *z, int d, int lo0, int n) { s[d] = lo0 + 1; z[d] = n; e[d] = lo0 + n; };
  setp(I.xst, I.xen, I.xsz, 0, 0, nx);
  distribute(ny, p_row, st, sz); setp(I.xst, I.xen, I.xsz, 1, st[row], sz[row]);
  distribute(nz, p_col, st, sz); setp(I.xst, I.xen, I.xsz, 2, st[col], sz[col]);
  distribute(nx, p_row, st, sz); setp(I.yst, I.yen, I.ysz, 0, st[row], sz[row]);
  setp(I.yst, I.yen, I.ysz, 1, 0, ny);
  distribute(nz, p_col, st, sz); setp(I.yst, I.yen, I.ysz, 2, st[col], sz[col]);
  distribute(nx, p_row, st, sz); setp(I.zst, I.zen, I.zsz, 0, st[row], sz[row]);
  distribute(ny, p_col, st, sz); setp(I.zst, I.zen, I.zsz, 1, st[col], sz[col]);
  setp(I.zst, I.zen, I.zsz, 2, 0, nz);
  return I;
}

void decomp_compute(int nx, int ny, int nz, int p_row, int p_col, int rank, x3d_decomp_info *out) {
  if (p_row < 1 || p_col < 1 || rank < 0 || rank >= p_row * p_col) throw Error("x3d_decomp_compute: bad process grid");
  *out = make_info(p_row, p_col, rank / p_col, rank % p_col, nx, ny, nz);
}

// which: 0 x->y, 1 y->z, 2 z->y, 3 y->x.  Host-only description of the exchange of rank (row,col).
static void build_plan_host(int p_row, int p_col, int row, int col, int nx, int ny, int nz, int which, TransposePlan &T) {
  std::vector<int> xs, xz, yxs, yxz, ycs, ycz, zs, zz;
  distribute(nx, p_row, xs, xz);     // x split (y- and z-pencils)
  distribute(ny, p_row, yxs, yxz);   // y split in x-pencils
  distribute(ny, p_col, ycs, ycz);   // y split in z-pencils
  distribute(nz, p_col, zs, zz);     // z split (x- and y-pencils)
  const bool xy = (which == 0 || which == 3);
  const int np = xy ? p_row : p_col;
  const int me = xy ? row : col;
  T.npeers = np;
  auto side = [&](SidePlan &S, int d0, int d1, int d2, int axis, const std::vector<int> &bs, const std::vector<int> &bz) {
    S.dims[0] = d0; S.dims[1] = d1; S.dims[2] = d2; S.axis = axis;
    S.blk_start = bs; S.blk_size = bz;
    S.disp.assign(np, 0); S.count.assign(np, 0);
    long long off = 0;
    for (int m = 0; m < np; ++m) {
      long long c = 1;
      for (int d = 0; d < 3; ++d) c *= (d == axis ? bz[m] : S.dims[d]);
      S.count[m] = c; S.disp[m] = off; off += c;
    }
  };
  const int X[3] = {nx, yxz[row], zz[col]}, Y[3] = {xz[row], ny, zz[col]}, Z[3] = {xz[row], ycz[col], nz};
  (void)me;
  switch (which) {
    case 0: side(T.send, X[0], X[1], X[2], 0, xs, xz); side(T.recv, Y[0], Y[1], Y[2], 1, yxs, yxz); break;   // x->y
    case 3: side(T.send, Y[0], Y[1], Y[2], 1, yxs, yxz); side(T.recv, X[0], X[1], X[2], 0, xs, xz); break;   // y->x
    case 1: side(T.send, Y[0], Y[1], Y[2], 1, ycs, ycz); side(T.recv, Z[0], Z[1], Z[2], 2, zs, zz); break;   // y->z
    case 2: side(T.send, Z[0], Z[1], Z[2], 2, zs, zz); side(T.recv, Y[0], Y[1], Y[2], 1, ycs, ycz); break;   // z->y
    default: throw Error("bad transpose selector");
  }
}

// CPU-only: the exchange plan of one rank (tests drive a gloo all-to-all with it)
void transpose_plan_host(int nx, int ny, int nz, int p_row, int p_col, int rank, int which, int *npeers, int *peer_ranks,
                         long long *scount, long long *sdispl, long long *rcount, long long *rdispl, int *send_dims,
                         int *recv_dims) {
  TransposePlan T;
  const int row = rank / p_col, col = rank % p_col;
  build_plan_host(p_row, p_col, row, col, nx, ny, nz, which, T);
  *npeers = T.npeers;
  const bool xy = (which == 0 || which == 3);
  for (int m = 0; m < T.npeers; ++m) {
    peer_ranks[m] = xy ? m * p_col + col : row * p_col + m;
    scount[m] = T.send.count[m]; sdispl[m] = T.send.disp[m];
    rcount[m] = T.recv.count[m]; rdispl[m] = T.recv.disp[m];
  }
  for (int d = 0; d < 3; ++d) { send_dims[d] = T.send.dims[d]; recv_dims[d] = T.recv.dims[d]; }
}

// ---- pack / unpack kernels ------------------------------------------------------------------------
// arr: local pencil (d0,d1,d2); the extent along `axis` is cut into blocks, block m goes to / comes from
// peer m and is stored contiguously (natural order of the sub-box) at packed + disp[m].
template <typename T, bool PACK>
__global__ void k_boxcopy(T *__restrict__ packed, T *__restrict__ arr, int d0, int d1, int d2, int axis,
                          const int *__restrict__ blk_of, const int *__restrict__ blk_start, const int *__restrict__ blk_size,
                          const long long *__restrict__ disp) {
  const long long tot = static_cast<long long>(d0) * d1 * d2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % d0);
    const int j = static_cast<int>((idx / d0) % d1);
    const int k = static_cast<int>(idx / (static_cast<long long>(d0) * d1));
    const int c = axis == 0 ? i : (axis == 1 ? j : k);
    const int m = blk_of[c];
    const int cl = c - blk_start[m], bs = blk_size[m];
    long long off;
    if (axis == 0) off = cl + static_cast<long long>(bs) * (j + static_cast<long long>(d1) * k);
    else if (axis == 1) off = i + static_cast<long long>(d0) * (cl + static_cast<long long>(bs) * k);
    else off = i + static_cast<long long>(d0) * (j + static_cast<long long>(d1) * cl);
    if (PACK) packed[disp[m] + off] = arr[idx];
    else arr[idx] = packed[disp[m] + off];
  }
}

static DecompImpl &DEC(Ctx &ctx) {
  auto *D = dynamic_cast<DecompImpl *>(ctx.decomp.get());
  if (!D) throw Error("decomposition: x3d_decomp_init has not been called");
  return *D;
}

static void upload_side(Ctx &ctx, DecompImpl &D, SidePlan &S) {
  const int np = static_cast<int>(S.blk_start.size());
  const int ext = S.dims[S.axis];
  std::vector<int> meta(ext + 2 * np);
  for (int m = 0; m < np; ++m)
    for (int q = 0; q < S.blk_size[m]; ++q) meta[S.blk_start[m] + q] = m;
  for (int m = 0; m < np; ++m) { meta[ext + m] = S.blk_start[m]; meta[ext + np + m] = S.blk_size[m]; }
  X3D_CUDA(cudaMalloc(&S.d_meta, meta.size() * sizeof(int)));
  X3D_CUDA(cudaMalloc(&S.d_disp, np * sizeof(long long)));
  D.dev_allocs.push_back(S.d_meta); D.dev_allocs.push_back(S.d_disp);
  X3D_CUDA(cudaMemcpyAsync(S.d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(S.d_disp, S.disp.data(), np * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}

static TransposePlan &get_plan(Ctx &ctx, DecompImpl &D, int id, int which) {
  if (id < 0 || id >= static_cast<int>(D.infos.size())) throw Error("bad decomposition id");
  TransposePlan &T = D.plans[id][which];
  if (!T.built) {
    build_plan_host(D.p_row, D.p_col, D.row, D.col, D.gdims[id][0], D.gdims[id][1], D.gdims[id][2], which, T);
    upload_side(ctx, D, T.send);
    upload_side(ctx, D, T.recv);
    T.built = true;
  }
  return T;
}

static void p2p_setup_group(Ctx &ctx, DecompImpl &D, DecompImpl::Group &G, ncclComm_t comm, int np, int me);

void decomp_init(Ctx &ctx, int nx, int ny, int nz, int p_row, int p_col, int rank, int nranks, const void *nccl_id) {
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (p_row < 1 || p_col < 1 || p_row * p_col != nranks) throw Error("x3d_decomp_init: p_row*p_col must equal nranks");
  if (rank < 0 || rank >= nranks) throw Error("x3d_decomp_init: bad rank");
  auto D = std::make_unique<DecompImpl>();
  D->nx = nx; D->ny = ny; D->nz = nz; D->p_row = p_row; D->p_col = p_col; D->rank = rank; D->nranks = nranks;
  D->row = rank / p_col; D->col = rank % p_col;
  D->infos.push_back(make_info(p_row, p_col, D->row, D->col, nx, ny, nz));
  D->gdims.push_back({nx, ny, nz});
  D->plans.emplace_back();
  if (nranks > 1 && nccl_id) {
    NcclApi &N = nccl();
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    std::memcpy(&id, nccl_id, sizeof(id));
    X3D_NCCL(N.CommInitRank(&D->world, nranks, id, rank));
    // x<->y: ranks sharing my column index; y<->z: ranks sharing my row index
    X3D_NCCL(N.CommSplit(D->world, D->col, D->row, &D->comm_row, nullptr));
    X3D_NCCL(N.CommSplit(D->world, D->row, D->col, &D->comm_col, nullptr));
    D->have_nccl = true;
    const char *e = getenv("X3D_P2P");
    D->p2p = !(e && atoi(e) == 0);
    if (const char *m = getenv("X3D_P2P_MODE")) D->p2p_mode = atoi(m);
    {
      double secs = 120.0;
      if (const char *t = getenv("X3D_BARRIER_TIMEOUT_S")) secs = atof(t);
      int khz = 1900000;
      cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx.device);
      D->barrier_timeout_cycles = static_cast<long long>(secs * 1e3 * static_cast<double>(khz));
      X3D_CUDA(cudaHostAlloc(&D->h_status, sizeof(int), cudaHostAllocMapped));
      *D->h_status = 0;
      X3D_CUDA(cudaHostGetDevicePointer(&D->d_status, D->h_status, 0));
    }
    // every rank must take the same path afterwards: agree on the outcome over the world communicator after each group
    auto agree = [&]() {
      int mine = D->p2p ? 1 : 0, *d_flag = nullptr;
      X3D_CUDA(cudaMalloc(&d_flag, sizeof(int)));
      X3D_CUDA(cudaMemcpyAsync(d_flag, &mine, sizeof(int), cudaMemcpyHostToDevice, ctx.stream));
      X3D_NCCL(N.AllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, D->world, ctx.stream));
      X3D_CUDA(cudaMemcpyAsync(&mine, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
      X3D_CUDA(cudaStreamSynchronize(ctx.stream));
      cudaFree(d_flag);
      D->p2p = mine != 0;
    };
    if (D->p2p) p2p_setup_group(ctx, *D, D->grp_row, D->comm_row, p_row, D->row);
    agree();
    if (D->p2p) p2p_setup_group(ctx, *D, D->grp_col, D->comm_col, p_col, D->col);
    agree();
  }
  ctx.decomp = std::move(D);
}

void nccl_unique_id(void *out128) {
  ncclUniqueId id;
  X3D_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
}

int decomp_info_init(Ctx &ctx, int nx, int ny, int nz) {
  DecompImpl &D = DEC(ctx);
  D.infos.push_back(make_info(D.p_row, D.p_col, D.row, D.col, nx, ny, nz));
  D.gdims.push_back({nx, ny, nz});
  D.plans.emplace_back();
  return static_cast<int>(D.infos.size()) - 1;
}

void decomp_info_get(Ctx &ctx, int id, x3d_decomp_info *out) {
  DecompImpl &D = DEC(ctx);
  if (id < 0 || id >= static_cast<int>(D.infos.size())) throw Error("bad decomposition id");
  *out = D.infos[id];
}

static int grid_for(const Ctx &ctx, long long n) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(ctx.sm_count) * 16;
  return static_cast<int>(b < cap ? b : cap);
}

template <bool PACK>
static void boxcopy(Ctx &ctx, const SidePlan &S, double *packed, double *arr, int elem) {
  const long long tot = static_cast<long long>(S.dims[0]) * S.dims[1] * S.dims[2];
  if (tot == 0) return;
  const int ext = S.dims[S.axis], np = static_cast<int>(S.blk_start.size());
  const int *blk_of = S.d_meta, *bst = S.d_meta + ext, *bsz = S.d_meta + ext + np;
  ProfScope ps(ctx, PACK ? "transpose_pack(k_boxcopy)" : "transpose_unpack(k_boxcopy)");
  if (elem == 1)
    k_boxcopy<double, PACK><<<grid_for(ctx, tot), 256, 0, ctx.stream>>>(packed, arr, S.dims[0], S.dims[1], S.dims[2], S.axis, blk_of, bst, bsz, S.d_disp);
  else
    k_boxcopy<double2, PACK><<<grid_for(ctx, tot), 256, 0, ctx.stream>>>(reinterpret_cast<double2 *>(packed), reinterpret_cast<double2 *>(arr),
                                                                     S.dims[0], S.dims[1], S.dims[2], S.axis, blk_of, bst, bsz, S.d_disp);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

void transpose_pack(Ctx &ctx, int which, const double *d_src, double *d_packed, int id, int elem) {
  DecompImpl &D = DEC(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  boxcopy<true>(ctx, get_plan(ctx, D, id, which).send, d_packed, const_cast<double *>(d_src), elem);
}
void transpose_unpack(Ctx &ctx, int which, const double *d_packed, double *d_dst, int id, int elem) {
  DecompImpl &D = DEC(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  boxcopy<false>(ctx, get_plan(ctx, D, id, which).recv, const_cast<double *>(d_packed), d_dst, elem);
}

// ---- peer-to-peer path ---------------------------------------------------------------------------------
namespace {

struct XchgRec {            // what every member publishes about one of its buffers
  unsigned char handle[64];
  unsigned long long offset;
  unsigned long long valid;
};

// one thread per group member: publish my epoch in the member's flag array, wait for the member's epoch in mine
__global__ void k_group_barrier(unsigned long long *const *__restrict__ peer_flags, unsigned long long *my_flags, int me, int np,
                                unsigned long long epoch, long long timeout_cycles, int *status) {
  const int m = threadIdx.x;
  if (m >= np) return;
  __threadfence_system();
  unsigned long long *theirs = peer_flags[m] + me;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  unsigned long long seen;
  const long long t0 = clock64();
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(my_flags + m) : "memory");
    if (seen >= epoch) break;
    __nanosleep(100);
    if (clock64() - t0 > timeout_cycles) {   // a member never arrived: report and go on instead of hanging or trapping
      *reinterpret_cast<volatile int *>(status) = 1 + m;
      __threadfence_system();
      break;
    }
  } while (true);
}

// element (i,j,k) of the local source pencil -> destination pencil of the member that owns it
template <typename T>
__global__ void k_p2p_transpose(const T *__restrict__ src, T *const *__restrict__ dst, int d0, int d1, int d2, int as, int ar,
                                const int *__restrict__ blk_of, const int *__restrict__ blk_start, const int *__restrict__ blk_size,
                                int my_off, int R0, int R1, int R2) {
  const long long tot = static_cast<long long>(d0) * d1 * d2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    int c[3];
    c[0] = static_cast<int>(idx % d0);
    c[1] = static_cast<int>((idx / d0) % d1);
    c[2] = static_cast<int>(idx / (static_cast<long long>(d0) * d1));
    const int cs = as == 0 ? c[0] : (as == 1 ? c[1] : c[2]);
    const int m = blk_of[cs];
    int D[3] = {R0, R1, R2};
    const int bs = blk_size[m], cl = cs - blk_start[m];
    if (as == 0) { c[0] = cl; D[0] = bs; } else if (as == 1) { c[1] = cl; D[1] = bs; } else { c[2] = cl; D[2] = bs; }
    if (ar == 0) c[0] += my_off; else if (ar == 1) c[1] += my_off; else c[2] += my_off;
    dst[m][c[0] + static_cast<long long>(D[0]) * (c[1] + static_cast<long long>(D[1]) * c[2])] = src[idx];
  }
}


// Whole-row blocks (the y<->z transposes of a slab or pencil decomposition: x rows are neither cut nor gathered).
// The block for member m is `height` contiguous runs of `width` bytes in the source pencil and in m's destination
// pencil, so the transpose is one pitched copy per member.
struct BlockCopy {
  const char *src;
  char *dst;
  unsigned long long width, spitch, dpitch;
  int height;
};
constexpr int P2P_MAX_MEMBERS = 16;
struct BlockCopies {
  BlockCopy b[P2P_MAX_MEMBERS];
};
constexpr int SEG_BYTES = 16384;   // one CTA iteration: 256 threads x 4 x 16 bytes
__global__ void __launch_bounds__(256) k_p2p_blocks(const __grid_constant__ BlockCopies a) {
  const BlockCopy &b = a.b[blockIdx.y];
  const unsigned long long segs_row = (b.width + SEG_BYTES - 1) / SEG_BYTES;
  const unsigned long long total = segs_row * static_cast<unsigned long long>(b.height);
  for (unsigned long long sg = blockIdx.x; sg < total; sg += gridDim.x) {
    const unsigned long long row = sg / segs_row, off = (sg - row * segs_row) * SEG_BYTES;
    const unsigned long long left = b.width - off;
    const int n16 = static_cast<int>((left < SEG_BYTES ? left : SEG_BYTES) >> 4);
    const uint4 *sp = reinterpret_cast<const uint4 *>(b.src + row * b.spitch + off);
    uint4 *dp = reinterpret_cast<uint4 *>(b.dst + row * b.dpitch + off);
    uint4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < n16) v[q] = __ldcs(sp + i);   // read once: streaming
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < n16) dp[i] = v[q];
    }
  }
}
}  // namespace

static void group_barrier(Ctx &ctx, DecompImpl::Group &G) {
  G.epoch++;
  k_group_barrier<<<1, 32, 0, ctx.stream>>>(static_cast<unsigned long long *const *>(G.d_peer_flags.p),
                                             static_cast<unsigned long long *>(G.flags.p), G.me, G.np, G.epoch, G.timeout_cycles, G.d_status);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

// all members publish (IPC handle, offset) of one local buffer; returns the members' mapped pointers, or false
// when any member cannot export its buffer (then every member falls back to the NCCL path)
static bool exchange_pointers(Ctx &ctx, DecompImpl &D, DecompImpl::Group &G, const void *local, std::vector<void *> &out) {
  XchgRec mine{};
  void *base = nullptr;
  size_t size = 0;
  if (local && find_alloc(local, &base, &size)) {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, base) == cudaSuccess) {
      static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
      std::memcpy(mine.handle, &h, 64);
      mine.offset = static_cast<unsigned long long>(static_cast<const char *>(local) - static_cast<const char *>(base));
      mine.valid = 1;
    } else {
      cudaGetLastError();
    }
  }
  D.xchg_send.reserve(sizeof(XchgRec));
  D.xchg_recv.reserve(sizeof(XchgRec) * G.np);
  X3D_CUDA(cudaMemcpyAsync(D.xchg_send.p, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx.stream));
  X3D_NCCL(nccl().AllGather(D.xchg_send.p, D.xchg_recv.p, sizeof(XchgRec), ncclChar, G.comm, ctx.stream));
  std::vector<XchgRec> all(G.np);
  X3D_CUDA(cudaMemcpyAsync(all.data(), D.xchg_recv.p, sizeof(XchgRec) * G.np, cudaMemcpyDeviceToHost, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  for (int m = 0; m < G.np; ++m)
    if (!all[m].valid) return false;
  out.assign(G.np, nullptr);
  bool ok = true;
  for (int m = 0; m < G.np; ++m) {
    if (m == G.me) { out[m] = const_cast<void *>(local); continue; }
    std::array<unsigned char, 64> key;
    std::memcpy(key.data(), all[m].handle, 64);
    auto it = D.ipc_opened.find(key);
    void *mapped = nullptr;
    if (it != D.ipc_opened.end()) {
      mapped = it->second;
    } else {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, all[m].handle, 64);
      if (cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; mapped = nullptr; }
      else D.ipc_opened[key] = mapped;
    }
    out[m] = mapped ? static_cast<char *>(mapped) + all[m].offset : nullptr;
  }
  // agreement: a member that failed to map makes everybody fall back
  unsigned long long flag = ok ? 1 : 0;
  XchgRec v{};
  v.valid = flag;
  X3D_CUDA(cudaMemcpyAsync(D.xchg_send.p, &v, sizeof(v), cudaMemcpyHostToDevice, ctx.stream));
  X3D_NCCL(nccl().AllGather(D.xchg_send.p, D.xchg_recv.p, sizeof(XchgRec), ncclChar, G.comm, ctx.stream));
  X3D_CUDA(cudaMemcpyAsync(all.data(), D.xchg_recv.p, sizeof(XchgRec) * G.np, cudaMemcpyDeviceToHost, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  for (int m = 0; m < G.np; ++m)
    if (!all[m].valid) return false;
  return true;
}

static void p2p_setup_group(Ctx &ctx, DecompImpl &D, DecompImpl::Group &G, ncclComm_t comm, int np, int me) {
  G.np = np; G.me = me; G.comm = comm;
  G.timeout_cycles = D.barrier_timeout_cycles; G.d_status = D.d_status;
  if (np <= 1) return;
  if (np > 32) { D.p2p = false; return; }   // the flag barrier has one thread per member: larger groups use NCCL
  G.flags.reserve(sizeof(unsigned long long) * 32);
  X3D_CUDA(cudaMemsetAsync(G.flags.p, 0, sizeof(unsigned long long) * 32, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  std::vector<void *> ptrs;
  if (!exchange_pointers(ctx, D, G, G.flags.p, ptrs)) { D.p2p = false; return; }
  G.d_peer_flags.reserve(sizeof(void *) * np);
  X3D_CUDA(cudaMemcpyAsync(G.d_peer_flags.p, ptrs.data(), sizeof(void *) * np, cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}

// member pointers of a destination buffer (device array), exchanged once per buffer; nullptr -> use NCCL
static void *const *p2p_dst(Ctx &ctx, DecompImpl &D, DecompImpl::Group &G, const void *dst, const std::vector<void *> **hosts = nullptr) {
  // Only buffers the library allocated itself take this path: they are allocated by the same code on every
  // member, so cache hits and misses (= collective pointer exchanges) happen on all members together.  A
  // caller-owned buffer (e.g. a framework tensor) has no such symmetry and always uses the NCCL exchange.
  void *base = nullptr;
  size_t size = 0;
  unsigned long long gen = 0;
  if (!find_alloc(dst, &base, &size, &gen)) return nullptr;
  const DecompImpl::Group::Key key{dst, gen};
  if (hosts) *hosts = nullptr;
  auto it = G.dst_cache.find(key);
  if (it != G.dst_cache.end()) { if (hosts) *hosts = &G.dst_host[key]; return static_cast<void *const *>(it->second.p); }
  if (G.dst_bad.count(key)) return nullptr;
  std::vector<void *> ptrs;
  if (!exchange_pointers(ctx, D, G, dst, ptrs)) { G.dst_bad[key] = true; return nullptr; }
  G.dst_host[key] = ptrs;
  if (hosts) *hosts = &G.dst_host[key];
  DevBuf &b = G.dst_cache[key];
  b.reserve(sizeof(void *) * G.np);
  X3D_CUDA(cudaMemcpyAsync(b.p, ptrs.data(), sizeof(void *) * G.np, cudaMemcpyHostToDevice, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  return static_cast<void *const *>(b.p);
}


// the per-member pitched copies of a whole-row transpose; false when the transpose is not of that shape
static bool block_copies(const TransposePlan &T, const DecompImpl::Group &G, const double *d_src, const std::vector<void *> &peers, int elem,
                         BlockCopies &bc, int &nb) {
  const SidePlan &S = T.send;
  const int as = S.axis, ar = T.recv.axis, np = T.npeers;
  if (as == 0 || ar == 0 || as == ar || np > P2P_MAX_MEMBERS) return false;
  const unsigned long long es = static_cast<unsigned long long>(elem) * sizeof(double);
  const unsigned long long d0 = S.dims[0], d1 = S.dims[1], d2 = S.dims[2];
  const unsigned long long R0 = T.recv.dims[0], R1 = T.recv.dims[1];
  const unsigned long long my_off = T.recv.blk_start[G.me];
  nb = 0;
  for (int q = 0; q < np; ++q) {
    const int m = (G.me + 1 + q) % np;   // start with the next member so that members do not all store to the same one; self last
    const unsigned long long bs = S.blk_size[m], b0 = S.blk_start[m];
    if (bs == 0 || d0 * d1 * d2 == 0) continue;
    BlockCopy &b = bc.b[nb];
    const char *src = reinterpret_cast<const char *>(d_src);
    char *dst = static_cast<char *>(peers[m]);
    if (as == 1) {  // cut along y, gathered along z: plane k of the block -> plane k + my_off of m's (d0, bs, nz) pencil
      b.width = d0 * bs * es; b.height = static_cast<int>(d2); b.spitch = d0 * d1 * es; b.dpitch = b.width;
      b.src = src + d0 * b0 * es; b.dst = dst + R0 * bs * my_off * es;
    } else {        // cut along z, gathered along y: plane b0 + l -> rows my_off.. of plane l of m's (d0, ny, bs) pencil
      b.width = d0 * d1 * es; b.height = static_cast<int>(bs); b.spitch = b.width; b.dpitch = R0 * R1 * es;
      b.src = src + d0 * d1 * b0 * es; b.dst = dst + R0 * my_off * es;
    }
    if ((b.width | b.spitch | b.dpitch | reinterpret_cast<uintptr_t>(b.src) | reinterpret_cast<uintptr_t>(b.dst)) & 15u) return false;
    ++nb;
  }
  return true;
}

// run the block copies of one field (between the two group barriers)
static void run_block_copies(Ctx &ctx, DecompImpl &D, const BlockCopies &bc, int nb) {
  if (nb == 0) return;
  int mode = D.p2p_mode;
  if (mode < 0) {
    unsigned long long bytes = 0;
    for (int q = 0; q < nb; ++q) bytes += bc.b[q].width * static_cast<unsigned long long>(bc.b[q].height);
    mode = bytes / nb >= (32ull << 20) ? 1 : 2;
  }
  if (mode == 1) {
    // copy engines: the members' blocks on the main stream, my own block (always the last entry when present) beside them
    if (!D.side) {
      X3D_CUDA(cudaStreamCreateWithFlags(&D.side, cudaStreamNonBlocking));
      X3D_CUDA(cudaEventCreateWithFlags(&D.ev_fork, cudaEventDisableTiming));
      X3D_CUDA(cudaEventCreateWithFlags(&D.ev_join, cudaEventDisableTiming));
    }
    X3D_CUDA(cudaEventRecord(D.ev_fork, ctx.stream));
    X3D_CUDA(cudaStreamWaitEvent(D.side, D.ev_fork, 0));
    for (int q = 0; q < nb; ++q) {
      const BlockCopy &b = bc.b[q];
      X3D_CUDA(cudaMemcpy2DAsync(b.dst, b.dpitch, b.src, b.spitch, b.width, b.height, cudaMemcpyDeviceToDevice, (q == nb - 1) ? D.side : ctx.stream));
    }
    X3D_CUDA(cudaEventRecord(D.ev_join, D.side));
    X3D_CUDA(cudaStreamWaitEvent(ctx.stream, D.ev_join, 0));
    return;
  }
  int gx = (8 * ctx.sm_count + nb - 1) / nb;
  if (gx < 1) gx = 1;
  k_p2p_blocks<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(nb)), 256, 0, ctx.stream>>>(bc);
  X3D_CUDA(cudaGetLastError());
  ctx.launches++;
}

static void count_traffic(DecompImpl &D, const TransposePlan &T, int which, int elem, int nf) {
  const int me = (which == 0 || which == 3) ? D.row : D.col;
  unsigned long long b = 0;
  for (int m = 0; m < T.npeers; ++m)
    if (m != me) b += static_cast<unsigned long long>(T.send.count[m]) * elem * sizeof(double);
  D.stat_remote_bytes += b * nf;
  D.stat_fields += nf;
}

void decomp_stats(Ctx &ctx, unsigned long long *remote_bytes, unsigned long long *fields) {
  DecompImpl &D = DEC(ctx);
  *remote_bytes = D.stat_remote_bytes;
  *fields = D.stat_fields;
}

// device pointers; src and dst pencils of decomposition `id`.  allow_p2p = false: the destination is not allocated the
// same way on every rank (the staging buffer of host-pointer calls grows with each rank's own sizes), so the collective
// pointer exchange of the peer-to-peer path could be entered by some ranks only: NCCL exchange instead.
static void transpose_device_impl(Ctx &ctx, int which, const double *d_src, double *d_dst, int id, int elem, bool allow_p2p) {
  DecompImpl &D = DEC(ctx);
  TransposePlan &T = get_plan(ctx, D, id, which);
  count_traffic(D, T, which, elem, 1);
  const long long ns = static_cast<long long>(T.send.dims[0]) * T.send.dims[1] * T.send.dims[2];
  const long long nr = static_cast<long long>(T.recv.dims[0]) * T.recv.dims[1] * T.recv.dims[2];
  if (T.npeers == 1) {  // the group has one rank: pencils coincide, bit-exact copy
    if (d_src != d_dst) X3D_CUDA(cudaMemcpyAsync(d_dst, d_src, ns * elem * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    return;
  }
  if (!D.have_nccl) throw Error("transpose: this context was initialised without a NCCL id (pack/unpack only)");
  if (D.p2p) {
    DecompImpl::Group &G = (which == 0 || which == 3) ? D.grp_row : D.grp_col;
    const std::vector<void *> *hosts = nullptr;
    void *const *peers = allow_p2p ? p2p_dst(ctx, D, G, d_dst, &hosts) : nullptr;
    if (peers) {
      const SidePlan &S = T.send;
      const int ext = S.dims[S.axis], np = T.npeers;
      const int *blk_of = S.d_meta, *bst = S.d_meta + ext, *bsz = S.d_meta + ext + np;
      const int my_off = T.recv.blk_start[G.me];
      group_barrier(ctx, G);  // every member has finished with its destination pencil
      BlockCopies bc;
      int nb = 0;
      if (D.p2p_mode != 0 && block_copies(T, G, d_src, *hosts, elem, bc, nb)) {
        ProfScope ps(ctx, "transpose_p2p(block copies)");
        run_block_copies(ctx, D, bc, nb);
      } else if (ns > 0) {
        ProfScope ps(ctx, "transpose_p2p(k_p2p_transpose)");
        // rows stay whole when neither the cut axis nor the gathered axis is x: move them as 16-byte pairs
        const bool pairs = elem == 1 && S.axis != 0 && T.recv.axis != 0 && (S.dims[0] % 2 == 0) &&
                           (reinterpret_cast<uintptr_t>(d_src) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_dst) % 16 == 0);
        if (pairs)
          k_p2p_transpose<double2><<<grid_for(ctx, ns / 2), 256, 0, ctx.stream>>>(reinterpret_cast<const double2 *>(d_src),
                                                                               reinterpret_cast<double2 *const *>(peers), S.dims[0] / 2, S.dims[1],
                                                                               S.dims[2], S.axis, T.recv.axis, blk_of, bst, bsz, my_off,
                                                                               T.recv.dims[0] / 2, T.recv.dims[1], T.recv.dims[2]);
        else if (elem == 1)
          k_p2p_transpose<double><<<grid_for(ctx, ns), 256, 0, ctx.stream>>>(d_src, reinterpret_cast<double *const *>(peers), S.dims[0], S.dims[1],
                                                                          S.dims[2], S.axis, T.recv.axis, blk_of, bst, bsz, my_off,
                                                                          T.recv.dims[0], T.recv.dims[1], T.recv.dims[2]);
        else
          k_p2p_transpose<double2><<<grid_for(ctx, ns), 256, 0, ctx.stream>>>(reinterpret_cast<const double2 *>(d_src),
                                                                           reinterpret_cast<double2 *const *>(peers), S.dims[0], S.dims[1], S.dims[2],
                                                                           S.axis, T.recv.axis, blk_of, bst, bsz, my_off, T.recv.dims[0],
                                                                           T.recv.dims[1], T.recv.dims[2]);
        X3D_CUDA(cudaGetLastError());
        ctx.launches++;
      }
      group_barrier(ctx, G);  // every member's stores into my pencil have landed
      return;
    }
  }
  D.sendbuf.reserve(ns * elem * sizeof(double));
  D.recvbuf.reserve(nr * elem * sizeof(double));
  double *sb = static_cast<double *>(D.sendbuf.p), *rb = static_cast<double *>(D.recvbuf.p);
  boxcopy<true>(ctx, T.send, sb, const_cast<double *>(d_src), elem);
  {
    ProfScope ps(ctx, "transpose_alltoall(NCCL)");
    NcclApi &N = nccl();
    ncclComm_t comm = (which == 0 || which == 3) ? D.comm_row : D.comm_col;
    X3D_NCCL(N.GroupStart());
    for (int m = 0; m < T.npeers; ++m) {
      if (T.send.count[m]) X3D_NCCL(N.Send(sb + T.send.disp[m] * elem, T.send.count[m] * elem, ncclDouble, m, comm, ctx.stream));
      if (T.recv.count[m]) X3D_NCCL(N.Recv(rb + T.recv.disp[m] * elem, T.recv.count[m] * elem, ncclDouble, m, comm, ctx.stream));
    }
    X3D_NCCL(N.GroupEnd());
  }
  boxcopy<false>(ctx, T.recv, rb, d_dst, elem);
}

void transpose_device(Ctx &ctx, int which, const double *d_src, double *d_dst, int id, int elem) {
  transpose_device_impl(ctx, which, d_src, d_dst, id, elem, true);
}

// the barrier kernels report a member that never arrived through mapped host memory: turn it into an error
void decomp_check(Ctx &ctx) {
  auto *D = dynamic_cast<DecompImpl *>(ctx.decomp.get());
  if (!D || !D->h_status) return;
  const int st = *reinterpret_cast<volatile int *>(D->h_status);
  if (st != 0)
    throw Error("pencil transpose: group member " + std::to_string(st - 1) + " did not reach a device-side barrier within "
                "X3D_BARRIER_TIMEOUT_S (default 120 s); the data of this context is no longer valid");
}

// several fields through the same transpose: one pair of group barriers around all the peer-store kernels
void transpose_device_multi(Ctx &ctx, int which, int nf, const double *const *d_src, double *const *d_dst, int id, int elem) {
  DecompImpl &D = DEC(ctx);
  TransposePlan &T = get_plan(ctx, D, id, which);
  bool p2p = D.p2p && D.have_nccl && T.npeers > 1 && nf > 1;
  DecompImpl::Group &G = (which == 0 || which == 3) ? D.grp_row : D.grp_col;
  std::vector<void *const *> peers(nf, nullptr);
  std::vector<const std::vector<void *> *> hosts(nf, nullptr);
  if (p2p)
    for (int f = 0; f < nf; ++f) {  // every member resolves every field (collective on a first use), then all agree
      peers[f] = p2p_dst(ctx, D, G, d_dst[f], &hosts[f]);
      if (!peers[f]) p2p = false;
    }
  if (!p2p) {
    for (int f = 0; f < nf; ++f) transpose_device(ctx, which, d_src[f], d_dst[f], id, elem);
    return;
  }
  count_traffic(D, T, which, elem, nf);
  const SidePlan &S = T.send;
  const long long ns = static_cast<long long>(S.dims[0]) * S.dims[1] * S.dims[2];
  const int ext = S.dims[S.axis], np = T.npeers;
  const int *blk_of = S.d_meta, *bst = S.d_meta + ext, *bsz = S.d_meta + ext + np;
  const int my_off = T.recv.blk_start[G.me];
  group_barrier(ctx, G);
  std::vector<BlockCopies> bcs(nf);
  std::vector<int> nbs(nf, 0);
  bool blocks = D.p2p_mode != 0;
  for (int f = 0; f < nf && blocks; ++f) blocks = block_copies(T, G, d_src[f], *hosts[f], elem, bcs[f], nbs[f]);
  if (blocks) {
    ProfScope ps(ctx, "transpose_p2p(block copies)");
    for (int f = 0; f < nf; ++f) run_block_copies(ctx, D, bcs[f], nbs[f]);
  } else if (ns > 0) {
    ProfScope ps(ctx, "transpose_p2p(k_p2p_transpose)");
    for (int f = 0; f < nf; ++f) {
      const bool pairs = elem == 1 && S.axis != 0 && T.recv.axis != 0 && (S.dims[0] % 2 == 0) &&
                         (reinterpret_cast<uintptr_t>(d_src[f]) % 16 == 0) && (reinterpret_cast<uintptr_t>(d_dst[f]) % 16 == 0);
      if (pairs || elem == 2) {
        const int half = pairs ? 2 : 1;
        k_p2p_transpose<double2><<<grid_for(ctx, ns / half), 256, 0, ctx.stream>>>(
            reinterpret_cast<const double2 *>(d_src[f]), reinterpret_cast<double2 *const *>(peers[f]), S.dims[0] / half, S.dims[1], S.dims[2],
            S.axis, T.recv.axis, blk_of, bst, bsz, my_off, T.recv.dims[0] / half, T.recv.dims[1], T.recv.dims[2]);
      } else {
        k_p2p_transpose<double><<<grid_for(ctx, ns), 256, 0, ctx.stream>>>(d_src[f], reinterpret_cast<double *const *>(peers[f]), S.dims[0],
                                                                          S.dims[1], S.dims[2], S.axis, T.recv.axis, blk_of, bst, bsz, my_off,
                                                                          T.recv.dims[0], T.recv.dims[1], T.recv.dims[2]);
      }
      X3D_CUDA(cudaGetLastError());
      ctx.launches++;
    }
  }
  group_barrier(ctx, G);
}

// ---- ring exchange of planes between neighbouring z slabs (column group) --------------------------------------
// Every item copies `bytes` from this rank's `src` into the buffer `dst` (a library-owned buffer that every rank allocates
// the same way) of the member `dir` places further along the ring (+1 next, -1 previous; periodic), at byte offset
// `dst_off`, with peer stores over NVLink between the same two flag barriers as the transposes.  All items of a call share
// one barrier pair.  False when the group has no peer-to-peer path (the caller then uses the transposes).
bool ring_exchange(Ctx &ctx, int n, const RingCopy *items) {
  DecompImpl &D = DEC(ctx);
  DecompImpl::Group &G = D.grp_col;
  if (!D.p2p || !D.have_nccl || G.np < 2 || n < 1 || n > P2P_MAX_MEMBERS) return false;
  X3D_CUDA(cudaSetDevice(ctx.device));
  std::vector<const std::vector<void *> *> hosts(n, nullptr);
  for (int q = 0; q < n; ++q)
    if (!p2p_dst(ctx, D, G, items[q].dst, &hosts[q])) return false;   // collective on a first use; all members agree
  BlockCopies bc;
  unsigned long long total = 0;
  for (int q = 0; q < n; ++q) {
    const int m = ((G.me + items[q].dir) % G.np + G.np) % G.np;
    BlockCopy &b = bc.b[q];
    b.src = static_cast<const char *>(items[q].src);
    b.dst = static_cast<char *>((*hosts[q])[m]) + items[q].dst_off;
    b.width = b.spitch = b.dpitch = items[q].bytes;
    b.height = 1;
    if ((b.width | reinterpret_cast<uintptr_t>(b.src) | reinterpret_cast<uintptr_t>(b.dst)) & 15u) throw Error("ring exchange: unaligned item");
    total += items[q].bytes;
  }
  D.stat_remote_bytes += total;
  group_barrier(ctx, G);   // the neighbours have finished reading what these stores overwrite
  {
    ProfScope ps(ctx, "slab_ring_exchange(k_p2p_blocks)");
    int gx = (8 * ctx.sm_count + n - 1) / n;
    k_p2p_blocks<<<dim3(static_cast<unsigned>(gx), static_cast<unsigned>(n)), 256, 0, ctx.stream>>>(bc);
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  }
  group_barrier(ctx, G);   // the neighbours' stores into my buffers have landed
  return true;
}
bool ring_available(Ctx &ctx) {
  auto *D = dynamic_cast<DecompImpl *>(ctx.decomp.get());
  return D && D->p2p && D->have_nccl && D->grp_col.np >= 2 && D->p_row == 1;
}

// ---- self-test of the production data plane --------------------------------------------------------------
// Fills a library-owned source pencil with the global linear index of every element (exact in a double), runs the
// transpose through the path the solver takes (peer stores / block copies / copy engines between library-owned
// pencils, or NCCL when peer access is unavailable) and counts, on the device, the destination elements that do not
// hold their own global index.  A transpose is pure data movement: the count must be zero (bit-exact).
namespace {
template <int ELEM>
__global__ void k_pattern(double *__restrict__ a, int d0, int d1, int d2, int o0, int o1, int o2, long long G0, long long G1, int check,
                          unsigned long long *bad) {
  const long long tot = static_cast<long long>(d0) * d1 * d2;
  unsigned long long nbad = 0;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < tot;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(idx % d0), j = static_cast<int>((idx / d0) % d1), k = static_cast<int>(idx / (static_cast<long long>(d0) * d1));
    const double v = static_cast<double>((i + o0) + G0 * ((j + o1) + G1 * (k + o2)));
    for (int e = 0; e < ELEM; ++e) {
      const double want = e == 0 ? v : -(v + 0.5);
      if (check) nbad += (a[idx * ELEM + e] != want);
      else a[idx * ELEM + e] = want;
    }
  }
  if (check && nbad) atomicAdd(bad, nbad);
}
}  // namespace

// which: 0 x->y, 1 y->z, 2 z->y, 3 y->x; elem 1 (real) / 2 (complex); mode: -1 default, else forces X3D_P2P_MODE for this call
long long transpose_selftest(Ctx &ctx, int which, int id, int elem, int mode) {
  DecompImpl &D = DEC(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (id < 0 || id >= static_cast<int>(D.infos.size())) throw Error("bad decomposition id");
  const x3d_decomp_info &I = D.infos[id];
  const int *sst, *ssz, *dst_, *dsz;
  switch (which) {
    case 0: sst = I.xst; ssz = I.xsz; dst_ = I.yst; dsz = I.ysz; break;
    case 1: sst = I.yst; ssz = I.ysz; dst_ = I.zst; dsz = I.zsz; break;
    case 2: sst = I.zst; ssz = I.zsz; dst_ = I.yst; dsz = I.ysz; break;
    case 3: sst = I.yst; ssz = I.ysz; dst_ = I.xst; dsz = I.xsz; break;
    default: throw Error("bad transpose selector");
  }
  const long long ns = static_cast<long long>(ssz[0]) * ssz[1] * ssz[2], nd = static_cast<long long>(dsz[0]) * dsz[1] * dsz[2];
  DevBuf &src = D.self_src, &dst = D.self_dst, &cnt = D.self_cnt;
  // Symmetric allocations: every rank reserves the size of the LARGEST pencil any rank holds in this decomposition (the
  // last mod(n, p) members of a group hold one more), not its own.  With per-rank sizes an uneven split (257 spectral
  // planes over 2 ranks) made one rank reallocate -- a new buffer generation, i.e. a collective pointer exchange -- while
  // the other hit its cache and went on to the flag barrier: a dead-lock.
  {
    auto cdiv = [](long long a, long long b) { return (a + b - 1) / b; };
    const long long g0 = D.gdims[id][0], g1 = D.gdims[id][1], g2 = D.gdims[id][2], pr = D.p_row, pc = D.p_col;
    const long long big = std::max({g0 * cdiv(g1, pr) * cdiv(g2, pc), cdiv(g0, pr) * g1 * cdiv(g2, pc), cdiv(g0, pr) * cdiv(g1, pc) * g2, 1LL});
    src.reserve(big * elem * sizeof(double));
    dst.reserve(big * elem * sizeof(double));
  }
  cnt.reserve(sizeof(unsigned long long));
  X3D_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), ctx.stream));
  X3D_CUDA(cudaMemsetAsync(dst.p, 0xff, std::max<long long>(nd, 1) * elem * sizeof(double), ctx.stream));
  const long long G0 = D.gdims[id][0], G1 = D.gdims[id][1];
  auto launch = [&](double *a, const int *sz, const int *st, long long n, int check) {
    if (n == 0) return;
    const int g = grid_for(ctx, n);
    if (elem == 1) k_pattern<1><<<g, 256, 0, ctx.stream>>>(a, sz[0], sz[1], sz[2], st[0] - 1, st[1] - 1, st[2] - 1, G0, G1, check, static_cast<unsigned long long *>(cnt.p));
    else k_pattern<2><<<g, 256, 0, ctx.stream>>>(a, sz[0], sz[1], sz[2], st[0] - 1, st[1] - 1, st[2] - 1, G0, G1, check, static_cast<unsigned long long *>(cnt.p));
    X3D_CUDA(cudaGetLastError());
    ctx.launches++;
  };
  launch(static_cast<double *>(src.p), ssz, sst, ns, 0);
  const int keep = D.p2p_mode;
  if (mode >= 0) D.p2p_mode = mode;
  try {
    transpose_device(ctx, which, static_cast<double *>(src.p), static_cast<double *>(dst.p), id, elem);
  } catch (...) { D.p2p_mode = keep; throw; }
  D.p2p_mode = keep;
  launch(static_cast<double *>(dst.p), dsz, dst_, nd, 1);
  unsigned long long bad = 0;
  X3D_CUDA(cudaMemcpyAsync(&bad, cnt.p, sizeof(bad), cudaMemcpyDeviceToHost, ctx.stream));
  X3D_CUDA(cudaStreamSynchronize(ctx.stream));
  return static_cast<long long>(bad);
}

// host-or-device entry of the C ABI
void transpose(Ctx &ctx, int which, const double *src, double *dst, int id, int elem) {
  DecompImpl &D = DEC(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  TransposePlan &T = get_plan(ctx, D, id, which);
  const size_t bs = static_cast<size_t>(T.send.dims[0]) * T.send.dims[1] * T.send.dims[2] * elem * sizeof(double);
  const size_t br = static_cast<size_t>(T.recv.dims[0]) * T.recv.dims[1] * T.recv.dims[2] * elem * sizeof(double);
  const bool sdev = is_device_ptr(src), ddev = is_device_ptr(dst);
  const double *ds = src;
  double *dd = dst;
  if (!sdev) { ctx.stage_in.reserve(bs); X3D_CUDA(cudaMemcpyAsync(ctx.stage_in.p, src, bs, cudaMemcpyHostToDevice, ctx.stream)); ds = static_cast<double *>(ctx.stage_in.p); }
  if (!ddev) { ctx.stage_out.reserve(br); dd = static_cast<double *>(ctx.stage_out.p); }
  transpose_device_impl(ctx, which, ds, dd, id, elem, ddev);
  if (!ddev) X3D_CUDA(cudaMemcpyAsync(dst, dd, br, cudaMemcpyDeviceToHost, ctx.stream));
  if (!sdev || !ddev) X3D_CUDA(cudaStreamSynchronize(ctx.stream));
}

// sum / max all-reduce of a few doubles on the device (diagnostics; navier.f90:360-361)
void allreduce(Ctx &ctx, double *d_buf, int n, bool is_max) {
  auto *D = dynamic_cast<DecompImpl *>(ctx.decomp.get());
  if (!D || D->nranks == 1) return;
  if (!D->have_nccl) throw Error("allreduce: no NCCL communicator");
  X3D_NCCL(nccl().AllReduce(d_buf, d_buf, n, ncclDouble, is_max ? ncclMax : ncclSum, D->world, ctx.stream));
}

void decomp_shape(Ctx &ctx, int *p_row, int *p_col, int *rank, int *nranks) {
  DecompImpl &D = DEC(ctx);
  *p_row = D.p_row; *p_col = D.p_col; *rank = D.rank; *nranks = D.nranks;
}

}  // namespace x3d
