// x3d_decomp.cu -- 2-D pencil decomposition bookkeeping and pencil transposes
// (2DECOMP&FFT v2.0.4, external: decomp_2d_init / decomp_info_init / transpose_*; call sites
// src/xcompact3d.f90:191-201, src/transeq.f90:163,236,318,437, src/poisson.f90:132-133).
//
// Distribution rule of 2DECOMP&FFT: an extent n over p ranks gives n/p points each, the LAST
// mod(n,p) ranks get one more.  The process grid is p_row x p_col; rank = row*p_col + col
// (MPI_CART row-major).  x-pencil (nx, ny/p_row, nz/p_col), y-pencil (nx/p_row, ny, nz/p_col),
// z-pencil (nx/p_row, ny/p_col, nz).  x<->y exchanges inside a column group of p_row ranks,
// y<->z inside a row group of p_col ranks.
#include "x3d_state.cuh"

namespace x3d {

void distribute(int n, int p, std::vector<int> &st, std::vector<int> &sz) {
  st.assign(p, 0); sz.assign(p, 0);
  const int base = n / p, rem = n % p;
  int s = 1;
  for (int r = 0; r < p; ++r) {
    sz[r] = base + (r >= p - rem ? 1 : 0);
    st[r] = s;
    s += sz[r];
  }
}

struct DecompImpl : DecompState {
  int nx = 0, ny = 0, nz = 0, p_row = 1, p_col = 1, rank = 0, nranks = 1;
  int row = 0, col = 0;
  std::vector<x3d_decomp_info> infos;
  std::vector<int> dims;  // 3 per info
};

static x3d_decomp_info make_info(const DecompImpl &D, int nx, int ny, int nz) {
  x3d_decomp_info I{};
  std::vector<int> st, sz;
  auto setp = [&](int *s, int *e, int *z, int d, int lo, int n) { s[d] = lo; z[d] = n; e[d] = lo + n - 1; };
  // x-pencil: x complete, y split by p_row (row index), z split by p_col (col index)
  setp(I.xst, I.xen, I.xsz, 0, 1, nx);
  distribute(ny, D.p_row, st, sz); setp(I.xst, I.xen, I.xsz, 1, st[D.row], sz[D.row]);
  distribute(nz, D.p_col, st, sz); setp(I.xst, I.xen, I.xsz, 2, st[D.col], sz[D.col]);
  // y-pencil: x split by p_row, y complete, z split by p_col
  distribute(nx, D.p_row, st, sz); setp(I.yst, I.yen, I.ysz, 0, st[D.row], sz[D.row]);
  setp(I.yst, I.yen, I.ysz, 1, 1, ny);
  distribute(nz, D.p_col, st, sz); setp(I.yst, I.yen, I.ysz, 2, st[D.col], sz[D.col]);
  // z-pencil: x split by p_row, y split by p_col, z complete
  distribute(nx, D.p_row, st, sz); setp(I.zst, I.zen, I.zsz, 0, st[D.row], sz[D.row]);
  distribute(ny, D.p_col, st, sz); setp(I.zst, I.zen, I.zsz, 1, st[D.col], sz[D.col]);
  setp(I.zst, I.zen, I.zsz, 2, 1, nz);
  return I;
}

static DecompImpl &DEC(Ctx &ctx) {
  auto *D = dynamic_cast<DecompImpl *>(ctx.decomp.get());
  if (!D) throw Error("decomposition: x3d_decomp_init has not been called");
  return *D;
}

void decomp_init(Ctx &ctx, int nx, int ny, int nz, int p_row, int p_col, int rank, int nranks, const void *nccl_id) {
  if (p_row < 1 || p_col < 1 || p_row * p_col != nranks) throw Error("x3d_decomp_init: p_row*p_col must equal nranks");
  if (rank < 0 || rank >= nranks) throw Error("x3d_decomp_init: bad rank");
  if (nranks > 1) throw Error("x3d_decomp_init: multi-rank transposes (NCCL) not wired yet");
  (void)nccl_id;
  auto D = std::make_unique<DecompImpl>();
  D->nx = nx; D->ny = ny; D->nz = nz; D->p_row = p_row; D->p_col = p_col; D->rank = rank; D->nranks = nranks;
  D->row = rank / p_col; D->col = rank % p_col;
  D->infos.push_back(make_info(*D, nx, ny, nz));
  D->dims = {nx, ny, nz};
  ctx.decomp = std::move(D);
}

int decomp_info_init(Ctx &ctx, int nx, int ny, int nz) {
  DecompImpl &D = DEC(ctx);
  D.infos.push_back(make_info(D, nx, ny, nz));
  D.dims.insert(D.dims.end(), {nx, ny, nz});
  return static_cast<int>(D.infos.size()) - 1;
}

void decomp_info_get(Ctx &ctx, int id, x3d_decomp_info *out) {
  DecompImpl &D = DEC(ctx);
  if (id < 0 || id >= static_cast<int>(D.infos.size())) throw Error("bad decomposition id");
  *out = D.infos[id];
}

// which: 0 x->y, 1 y->z, 2 z->y, 3 y->x ; elem = doubles per element (1 real, 2 complex)
void transpose(Ctx &ctx, int which, const double *src, double *dst, int id, int elem) {
  DecompImpl &D = DEC(ctx);
  X3D_CUDA(cudaSetDevice(ctx.device));
  if (id < 0 || id >= static_cast<int>(D.infos.size())) throw Error("bad decomposition id");
  const x3d_decomp_info &I = D.infos[id];
  const int *ssz = (which == 0) ? I.xsz : (which == 1 || which == 3) ? I.ysz : I.zsz;
  const size_t bytes = static_cast<size_t>(ssz[0]) * ssz[1] * ssz[2] * elem * sizeof(double);
  if (D.nranks == 1) {
    // one rank: all pencils coincide; the transpose is a plain (bit-exact) copy
    if (src != dst) X3D_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx.stream));
    if (!is_device_ptr(dst) || !is_device_ptr(src)) X3D_CUDA(cudaStreamSynchronize(ctx.stream));
    return;
  }
  throw Error("multi-rank transpose not wired yet");
}

}  // namespace x3d
