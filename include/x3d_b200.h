/*
 * x3d_b200.h — C ABI of the B200-native hot path for Xcompact3d.
 *
 * Every entry point below replaces one procedure of the reference
 * (xcompact3d/Incompact3d v5.0, paths relative to the reference tree); the
 * file:line of the interface it replaces is cited beside it.  The ABI is what
 * an ISO_C_BINDING `interface ... bind(C)` block binds to (see
 * incompact3d_b200/fortran/x3d_gpu.f90 and INTEGRATION.md):
 *
 *   - plain pointers and sizes only, no C++/torch types;
 *   - operator entry points keep the reference's argument list and order and
 *     take scalars BY REFERENCE (Fortran convention), preceded by the context
 *     handle (passed by value);
 *   - arrays are column-major (i fastest), real(8) / complex(8) / integer(4);
 *   - every pointer argument may be a HOST pointer or a DEVICE pointer; the
 *     library classifies it (cudaPointerGetAttributes).  Host fields are staged
 *     through device scratch (drop-in mode); device fields are used in place;
 *   - every function returns 0 on success, non-zero on error and never aborts;
 *     x3d_last_error() gives the message (the Fortran shim calls
 *     decomp_2d_abort, the reference's convention, src/schemes.f90:472-473).
 *
 * There is NO CPU fallback: without a CUDA device x3d_create fails.
 */
#ifndef X3D_B200_H
#define X3D_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct x3d_ctx x3d_ctx;

/* ---- lifecycle ------------------------------------------------------- */
int x3d_create(x3d_ctx **ctx, int device);
int x3d_destroy(x3d_ctx *ctx);
const char *x3d_last_error(void);
int x3d_version(void);
/* number of kernels of this library launched since context creation */
long long x3d_launch_count(const x3d_ctx *ctx);
/* blocks until all work queued on the context's stream has finished */
int x3d_sync(x3d_ctx *ctx);
/* the cudaStream_t the context launches on (as an integer handle) */
unsigned long long x3d_stream(x3d_ctx *ctx);

/* optional per-launch timing with CUDA events on the context stream, grouped by kernel class;
 * x3d_profile_end writes a JSON array of {name,count,total_ms,avg_ms} into buf                 */
int x3d_profile_begin(x3d_ctx *ctx);
int x3d_profile_end(x3d_ctx *ctx, char *buf, int cap);

/* ---- hidden module state made explicit -------------------------------
 * The reference operators read stencil scalars from modules derivX/Y/Z
 * (src/module_param.f90:559-617), filter scalars from parfiX/Y/Z
 * (src/module_param.f90:621-656) and flags from module param
 * (iibm src/derive.f90:23, istret :409, iimplicit :2166, nclx/y/z :3816).   */
typedef struct x3d_deriv_coeffs {
  /* first derivative, src/schemes.f90:443-520 */
  double alfa1, af1, bf1, cf1, df1, alfa2, af2;
  double alfan, afn, bfn, cfn, dfn, alfam, afm;
  double alfai, afi, bfi;
  /* second derivative, src/schemes.f90:602-740 */
  double alsa1, as1, bs1, cs1, ds1, alsa2, as2;
  double alsan, asn, bsn, csn, dsn, alsam, asm_;
  double alsa3, as3, bs3, alsat, ast, bst;
  double alsa4, as4, bs4, cs4, alsatt, astt, bstt, cstt;
  double alsai, asi, bsi, csi, dsi;
  /* staggered derivative / interpolation, src/schemes.f90:860-931 */
  double alcai6, aci6, bci6;
  double ailcai6, aici6, bici6, cici6, dici6;
} x3d_deriv_coeffs;

typedef struct x3d_filter_coeffs { /* src/filters.f90:62-138 */
  double fial1, fia1, fib1, fic1, fid1;
  double fial2, fia2, fib2, fic2, fid2;
  double fial3, fia3, fib3, fic3, fid3, fie3, fif3;
  double fialn, fian, fibn, ficn, fidn;
  double fialm, fiam, fibm, ficm, fidm;
  double fialp, fiap, fibp, ficp, fidp, fiep, fifp;
  double fiali, fiai, fibi, fici, fidi;
} x3d_filter_coeffs;

/* axis: 0 = x, 1 = y, 2 = z */
int x3d_set_deriv_coeffs(x3d_ctx *ctx, int axis, const x3d_deriv_coeffs *c);
int x3d_set_filter_coeffs(x3d_ctx *ctx, int axis, const x3d_filter_coeffs *c);
/* module param flags; nclx/ncly/nclz are the LOGICALs (1 = periodic)       */
int x3d_set_flags(x3d_ctx *ctx, int iibm, int istret, int iimplicit,
                  int nclx, int ncly, int nclz);
/* stretched-mesh metrics computed by the host's stretching() (src/stretching.f90:96-318), ny entries
 * each (the ...i arrays are the pressure-mesh ones); needed by x3d_poisson_init when istret != 0
 * (matrice_refinement, src/poisson.f90:1814-2249) and by the device solver                          */
int x3d_set_stretching(x3d_ctx *ctx, int ny, const double *yp, const double *ypi,
                       const double *ppy, const double *pp2y, const double *pp4y,
                       const double *ppyi, const double *pp2yi, const double *pp4yi);

/* ---- immersed-boundary pre-pass (iibm = 2, 3) ---------------------------
 * When iibm = 2 every derx/dery/derz, derxx/deryy/derzz and filx/fily/filz first rebuilds its INPUT inside the solid
 * bodies by Lagrange interpolation (lagpolx/y/z + polint, src/ibm.f90:83-389; call sites src/derive.f90:23,84,157,...,
 * src/filters.f90:235,...); when iibm = 3 by clamped cubic splines with the operator's `lind` as wall value
 * (cubsplx/y/z + cubic_spline, src/ibm.f90:399-968; call sites src/derive.f90:24,..., src/filters.f90:236,...).
 * The geometry is module complex_geometry (src/module_param.f90:546-556), filled by genepsi3d on the host:
 *   nobj(na,nb), xi/xf(nobjmax,na,nb), nipif/nfpif(0:nobjmax,na,nb) with (na,nb) = (ny,nz) for x lines,
 *   (nx,nz) for y lines, (nx,ny) for z lines -- local pencil extents; npif, izap from module param;
 *   d = mesh step, len = domain length; coords = yp(ncoords = ny) for axis 1 (NULL, 0 on the uniform axes).    */
int x3d_set_ibm_geometry(x3d_ctx *ctx, int axis, int nobjmax, int npif, int izap, int na, int nb, const int *nobj,
                         const double *xi, const double *xf, const int *nipif, const int *nfpif, const double *coords,
                         int ncoords, double d, double len);
/* lagpolx(u) / lagpoly(u) / lagpolz(u), src/ibm.f90:83,168,260: u (host or device) is modified in place */
int x3d_lagpolx(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz);
int x3d_lagpoly(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz);
int x3d_lagpolz(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz);
/* iibm = 3 with ianal /= 0: the wall positions analitic_x / analitic_y return for xi / xf (src/ibm.f90:412-417,
 * 1034-1070), computed by the host for its case; same shape as xi / xf.  NULL, NULL = ianal 0 (the default).        */
int x3d_set_ibm_analytic(x3d_ctx *ctx, int axis, const double *ana_i, const double *ana_f);
/* cubsplx(u,lind) / cubsply(u,lind) / cubsplz(u,lind), src/ibm.f90:399,559,724: u (host or device) is modified in
 * place.  Where the reference leaves a value undefined (a node that lies in no spline interval gets what the previous
 * spline call returned, possibly on another line) the library stores the previous value of the same line, starting
 * from lind; a body with xi == xf (the reference aborts) is left untouched.                                          */
int x3d_cubsplx(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz, const double *lind);
int x3d_cubsply(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz, const double *lind);
int x3d_cubsplz(x3d_ctx *ctx, double *u, const int *nx, const int *ny, const int *nz, const double *lind);

/* ---- compact operators ------------------------------------------------
 * abstract interfaces DERIVATIVE_X/Y/Z, src/module_param.f90:136-167 and
 * FILTER_X/Y/Z :203-226.  t is output; u input; r,s caller scratch (ignored:
 * the Sherman-Morrison vector is precomputed per coefficient set); ff,fs,fw
 * the LU arrays from prepare(), src/schemes.f90:413-439 (host pointers).     */
#define X3D_DECL_DERX(name)                                                    \
  int x3d_##name(x3d_ctx *ctx, double *tx, const double *ux, double *rx,       \
                 double *sx, const double *ffx, const double *fsx,             \
                 const double *fwx, const int *nx, const int *ny,              \
                 const int *nz, const int *npaire, const double *lind)
#define X3D_DECL_DERY(name)                                                    \
  int x3d_##name(x3d_ctx *ctx, double *ty, const double *uy, double *ry,       \
                 double *sy, const double *ffy, const double *fsy,             \
                 const double *fwy, const double *ppy, const int *nx,          \
                 const int *ny, const int *nz, const int *npaire,              \
                 const double *lind)
/* src/derive.f90:7,68,141,211,281 */
X3D_DECL_DERX(derx_00); X3D_DECL_DERX(derx_11); X3D_DECL_DERX(derx_12);
X3D_DECL_DERX(derx_21); X3D_DECL_DERX(derx_22);
/* src/derive.f90:325,423,545,664,783 */
X3D_DECL_DERY(dery_00); X3D_DECL_DERY(dery_11); X3D_DECL_DERY(dery_12);
X3D_DECL_DERY(dery_21); X3D_DECL_DERY(dery_22);
/* src/derive.f90:858,951,1064,1176,1288 */
X3D_DECL_DERX(derz_00); X3D_DECL_DERX(derz_11); X3D_DECL_DERX(derz_12);
X3D_DECL_DERX(derz_21); X3D_DECL_DERX(derz_22);
/* src/derive.f90:1354,1481,1666,1822,1978 */
X3D_DECL_DERX(derxx_00); X3D_DECL_DERX(derxx_11); X3D_DECL_DERX(derxx_12);
X3D_DECL_DERX(derxx_21); X3D_DECL_DERX(derxx_22);
/* src/derive.f90:2052,2207,2433,2631,2829 (no ppy argument) */
X3D_DECL_DERX(deryy_00); X3D_DECL_DERX(deryy_11); X3D_DECL_DERX(deryy_12);
X3D_DECL_DERX(deryy_21); X3D_DECL_DERX(deryy_22);
/* src/derive.f90:2923,3082,3307,3503,3699 */
X3D_DECL_DERX(derzz_00); X3D_DECL_DERX(derzz_11); X3D_DECL_DERX(derzz_12);
X3D_DECL_DERX(derzz_21); X3D_DECL_DERX(derzz_22);
/* src/filters.f90:221,292,384,470,558 / 605.. / 990.. */
X3D_DECL_DERX(filx_00); X3D_DECL_DERX(filx_11); X3D_DECL_DERX(filx_12);
X3D_DECL_DERX(filx_21); X3D_DECL_DERX(filx_22);
X3D_DECL_DERX(fily_00); X3D_DECL_DERX(fily_11); X3D_DECL_DERX(fily_12);
X3D_DECL_DERX(fily_21); X3D_DECL_DERX(fily_22);
X3D_DECL_DERX(filz_00); X3D_DECL_DERX(filz_11); X3D_DECL_DERX(filz_12);
X3D_DECL_DERX(filz_21); X3D_DECL_DERX(filz_22);

/* staggered operators (velocity mesh <-> pressure mesh).  Argument lists as in
 * the reference; note the order of (n, nm) differs between vp and pv.        */
/* src/derive.f90:3796 derxvp, :3911 interxvp */
int x3d_derxvp(x3d_ctx *ctx, double *tx, const double *ux, double *rx,
               double *sx, const double *cfx6, const double *csx6,
               const double *cwx6, const int *nx, const int *nxm,
               const int *ny, const int *nz, const int *npaire);
int x3d_interxvp(x3d_ctx *ctx, double *tx, const double *ux, double *rx,
                 double *sx, const double *cifx6, const double *cisx6,
                 const double *ciwx6, const int *nx, const int *nxm,
                 const int *ny, const int *nz, const int *npaire);
/* src/derive.f90:4041 derxpv, :4126 interxpv */
int x3d_derxpv(x3d_ctx *ctx, double *tx, const double *ux, double *rx,
               double *sx, const double *cfi6, const double *csi6,
               const double *cwi6, const double *cfx6, const double *csx6,
               const double *cwx6, const int *nxm, const int *nx,
               const int *ny, const int *nz, const int *npaire);
int x3d_interxpv(x3d_ctx *ctx, double *tx, const double *ux, double *rx,
                 double *sx, const double *cifi6, const double *cisi6,
                 const double *ciwi6, const double *cifx6, const double *cisx6,
                 const double *ciwx6, const int *nxm, const int *nx,
                 const int *ny, const int *nz, const int *npaire);
/* src/derive.f90:4265 interyvp, :4442 deryvp, :4587 interypv, :4775 derypv */
int x3d_interyvp(x3d_ctx *ctx, double *ty, const double *uy, double *ry,
                 double *sy, const double *cify6, const double *cisy6,
                 const double *ciwy6, const int *nx, const int *ny,
                 const int *nym, const int *nz, const int *npaire);
int x3d_deryvp(x3d_ctx *ctx, double *ty, const double *uy, double *ry,
               double *sy, const double *cfy6, const double *csy6,
               const double *cwy6, const double *ppyi, const int *nx,
               const int *ny, const int *nym, const int *nz,
               const int *npaire);
int x3d_interypv(x3d_ctx *ctx, double *ty, const double *uy, double *ry,
                 double *sy, const double *cifi6y, const double *cisi6y,
                 const double *ciwi6y, const double *cify6, const double *cisy6,
                 const double *ciwy6, const int *nx, const int *nym,
                 const int *ny, const int *nz, const int *npaire);
int x3d_derypv(x3d_ctx *ctx, double *ty, const double *uy, double *ry,
               double *sy, const double *cfi6y, const double *csi6y,
               const double *cwi6y, const double *cfy6, const double *csy6,
               const double *cwy6, const double *ppy, const int *nx,
               const int *nym, const int *ny, const int *nz,
               const int *npaire);
/* src/derive.f90:4920 derzvp, :5105 interzvp, :5287 derzpv, :5426 interzpv */
int x3d_derzvp(x3d_ctx *ctx, double *tz, const double *uz, double *rz,
               double *sz, const double *cfz6, const double *csz6,
               const double *cwz6, const int *nx, const int *ny, const int *nz,
               const int *nzm, const int *npaire);
int x3d_interzvp(x3d_ctx *ctx, double *tz, const double *uz, double *rz,
                 double *sz, const double *cifz6, const double *cisz6,
                 const double *ciwz6, const int *nx, const int *ny,
                 const int *nz, const int *nzm, const int *npaire);
int x3d_derzpv(x3d_ctx *ctx, double *tz, const double *uz, double *rz,
               double *sz, const double *cfiz6, const double *csiz6,
               const double *cwiz6, const double *cfz6, const double *csz6,
               const double *cwz6, const int *nx, const int *ny,
               const int *nzm, const int *nz, const int *npaire);
int x3d_interzpv(x3d_ctx *ctx, double *tz, const double *uz, double *rz,
                 double *sz, const double *cifiz6, const double *cisiz6,
                 const double *ciwiz6, const double *cifz6, const double *cisz6,
                 const double *ciwz6, const int *nx, const int *ny,
                 const int *nzm, const int *nz, const int *npaire);

/* ---- 2DECOMP&FFT pencil decomposition (external library v2.0.4) ---------
 * decomp_2d_init (call site src/xcompact3d.f90:191) and decomp_info_init
 * (:197-201, src/poisson.f90:132-133).  A decomposition handle plays the role
 * of TYPE(DECOMP_INFO); id 0 is the main (nx,ny,nz) decomposition.           */
typedef struct x3d_decomp_info {
  int xst[3], xen[3], xsz[3]; /* 1-based inclusive, as in DECOMP_INFO */
  int yst[3], yen[3], ysz[3];
  int zst[3], zen[3], zsz[3];
} x3d_decomp_info;
/* nranks = p_row*p_col, rank in [0,nranks). Single-process contexts use
 * nranks = 1.  nccl_unique_id: 128 opaque bytes shared by all ranks (may be
 * NULL when nranks == 1).                                                    */
int x3d_decomp_init(x3d_ctx *ctx, int nx, int ny, int nz, int p_row, int p_col,
                    int rank, int nranks, const void *nccl_unique_id);
int x3d_nccl_unique_id(void *out128);
int x3d_decomp_info_init(x3d_ctx *ctx, int nx, int ny, int nz, int *decomp_id);
int x3d_decomp_info_get(x3d_ctx *ctx, int decomp_id, x3d_decomp_info *out);
/* traffic counters since x3d_decomp_init: bytes this rank sent to OTHER ranks through transposes (what crosses
 * NVLink), and the number of transposed fields                                                              */
int x3d_decomp_stats(x3d_ctx *ctx, unsigned long long *remote_bytes, unsigned long long *fields);
/* Self-test of the data plane the device solver uses (collective: every rank calls it): a library-owned source pencil
 * is filled with the global index of each element, transposed (which: 0 x->y, 1 y->z, 2 z->y, 3 y->x) through the
 * peer-to-peer path (mode -1: as the solver would; 0 element kernel, 1 copy engines, 2 vector-copy kernel) and the
 * destination is compared, on the device, with the indices it must hold.  *mismatches = 0 <=> bit-exact.        */
int x3d_transpose_selftest(x3d_ctx *ctx, int which, int decomp_id, int complex_, int mode, long long *mismatches);
/* transpose_x_to_y etc. (call sites src/transeq.f90:163,236,318,437); the
 * complex variants are used on the spectral decomposition sp
 * (src/poisson.f90:759).  Bit-exact data movement.                           */
/* CPU-only helpers (no context, no GPU): the pencil extents of any rank, and the all-to-all(v)
 * plan of one transpose (which: 0 x->y, 1 y->z, 2 z->y, 3 y->x): peers (global ranks), element counts
 * and displacements in the packed buffers, and the local send / receive pencil shapes.              */
int x3d_decomp_compute(int nx, int ny, int nz, int p_row, int p_col, int rank, x3d_decomp_info *out);
int x3d_transpose_plan(int nx, int ny, int nz, int p_row, int p_col, int rank, int which, int *npeers,
                       int *peer_ranks, long long *scount, long long *sdispl, long long *rcount,
                       long long *rdispl, int *send_dims, int *recv_dims);
/* the two halves of a transpose, device pointers: pencil -> packed send buffer, packed receive buffer ->
 * pencil (what x3d_transpose_* runs around the NCCL exchange)                                       */
int x3d_transpose_pack(x3d_ctx *ctx, int which, const double *src, double *packed, int decomp_id, int complex_);
int x3d_transpose_unpack(x3d_ctx *ctx, int which, const double *packed, double *dst, int decomp_id, int complex_);
int x3d_transpose_x_to_y(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_y_to_z(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_z_to_y(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_y_to_x(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_x_to_y_complex(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_y_to_z_complex(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_z_to_y_complex(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);
int x3d_transpose_y_to_x_complex(x3d_ctx *ctx, const double *src, double *dst, int decomp_id);

/* ---- spectral Poisson solver -------------------------------------------
 * decomp_2d_poisson_init src/poisson.f90:73 and the `poisson` procedure
 * pointer :63 (poisson_000 :298, _100 :413, _010 :665, _11x :1019).          */
typedef struct x3d_poisson_params {
  int nx, ny, nz;          /* velocity-mesh node counts (nx_global ...)     */
  int bcx, bcy, bcz;       /* 0 periodic, 1 otherwise (src/poisson.f90:79-93)*/
  double xlx, yly, zlz;    /* domain lengths                                */
  int istret;              /* stretched-mesh option                         */
  double alpha, beta;      /* mod_stret parameters (src/stretching.f90)     */
} x3d_poisson_params;
int x3d_poisson_init(x3d_ctx *ctx, const x3d_poisson_params *p);
/* rhs: z-pencil of the pressure mesh (ph%zsz), in place */
int x3d_poisson(x3d_ctx *ctx, double *rhs);

/* ---- device-resident solver -------------
 * momentum_rhs_eq src/transeq.f90:73, intt src/time_integrators.f90:18,
 * pre_correc/divergence/gradp/cor_vel src/navier.f90:502,257,386,206,
 * init_tgv / postprocess_tgv src/Case-TGV.f90:25,189.                        */
typedef struct x3d_solver_params {
  int nx, ny, nz;                 /* nodes */
  int nclx1, nclxn, ncly1, nclyn, nclz1, nclzn;
  double xlx, yly, zlz;
  double re, dt;
  int ifirstder, isecondder, ipinter, itimescheme; /* 1 Euler, 2 AB2, 3 AB3, 5 RK3 (src/variables.f90:1340-1399) */
  int istret; double beta;
  double nu0nu, cnu;
  int p_row, p_col;
  int itype;                      /* 0: box without forcing (TGV); 3: channel (itype_channel, src/module_param.f90):
                                     constant flow rate channel_cfr, src/Case-Channel.f90:150-170,220-261;
                                     5: cylinder wake (itype_cyl): inflow / convective outflow, Case-Cylinder-wake.f90 */
} x3d_solver_params;
/* Case parameters beyond the mesh and the schemes (call after x3d_solver_init, before stepping).
 *  channel (itype 3): momentum_forcing_channel, src/Case-Channel.f90:396-420 -- cpg /= 0: constant pressure gradient
 *    (re is then Re_tau: xnu = 1/re_cent, fcpg = 2/yly (re/re_cent)^2, src/parameters.f90:303-311, and channel_cfr is
 *    off, src/Case-Channel.f90:157); spin-up rotation wrotation while itime < spinup_time and iin <= 2.
 *  cylinder (itype 5): u1, u2, inflow_noise of inflow / outflow, src/Case-Cylinder-wake.f90:100-203.
 *  immersed boundary: iibm 0 | 2 (lagpol* in front of every derivative of momentum_rhs_eq, src/derive.f90:23-24) |
 *    3 (cubspl*); ubcx, ubcy, ubcz = the body's velocity, used for lind and by the ep1 mask in divergence
 *    (src/navier.f90:285-293).  The geometry goes in through x3d_set_ibm_geometry, the mask through
 *    x3d_solver_set_ibm_mask.                                                                                   */
typedef struct x3d_case_params {
  int cpg; double wrotation; int spinup_time; int iin;
  double u1, u2, inflow_noise;
  int iibm; double ubcx, ubcy, ubcz;
} x3d_case_params;
int x3d_solver_set_case(x3d_ctx *ctx, const x3d_case_params *c);
/* ep1 of this rank's x-pencil (nx, ny, nz_local), host or device pointer (genepsi3d's output; 1 inside the body) */
int x3d_solver_set_ibm_mask(x3d_ctx *ctx, const double *ep1);
/* the random planes bxo, byo, bzo (ny, nz_local) that inflow() scales by inflow_noise (NULL = zero) */
int x3d_solver_set_inflow_noise(x3d_ctx *ctx, const double *bxo, const double *byo, const double *bzo);
/* wall velocities of the x faces for pre_correc (src/navier.f90:564-595): bxx1 bxy1 bxz1 bxxn bxyn bxzn, (ny, nz_local)
 * planes; NULL keeps a plane.  The cylinder case overwrites them every sub-step (inflow / outflow).             */
int x3d_solver_set_wall_velocity_x(x3d_ctx *ctx, const double *const planes6[6]);
int x3d_solver_get_wall_velocity_x(x3d_ctx *ctx, double *const planes6[6]);
/* apply_spatial_filter (src/tools.f90:600-675) on the solver's velocity: filx / fily / filz with the npaire pairing of
 * the reference (the component along the filtered direction is odd), transposes included; ifilter 1 (all), 2 (x and
 * z), 3 (y); af = the parameter of set_filter_coefficients (src/filters.f90:62-219)                             */
int x3d_solver_apply_spatial_filter(x3d_ctx *ctx, int ifilter, double af);
/* init_cyl with iin = 0 (src/Case-Cylinder-wake.f90:205-279): uniform stream ux = u1 */
int x3d_solver_init_cyl(x3d_ctx *ctx);
int x3d_solver_init(x3d_ctx *ctx, const x3d_solver_params *p);
int x3d_solver_init_tgv(x3d_ctx *ctx);
/* init_channel with iin = 0 (src/Case-Channel.f90:71-94): ux = 1 - y^2, uz = sin(x) + cos(z) */
int x3d_solver_init_channel(x3d_ctx *ctx);
/* set / get the x-pencil velocity fields (host or device pointers) */
int x3d_solver_set_velocity(x3d_ctx *ctx, const double *ux, const double *uy, const double *uz);
int x3d_solver_get_velocity(x3d_ctx *ctx, double *ux, double *uy, double *uz);
/* local x-pencil extents (nx, ny, nz_local) of this rank and its 0-based z offset */
int x3d_solver_local_shape(x3d_ctx *ctx, int *dims3, int *zstart0);
/* advance nsteps full time steps (iadvance_time sub-steps each) */
int x3d_solver_step(x3d_ctx *ctx, int nsteps);
/* One asynchronous job: copy the host x-pencil velocity (ux_in, uy_in, uz_in) to the device, advance it nsteps time
 * steps, copy the result to (ux_out, uy_out, uz_out).  Returns once the work is queued; x3d_solver_host_sync waits for
 * every queued job.  Consecutive jobs are independent of each other (ensemble members, parameter sweeps; the reference
 * would run them as separate MPI jobs, src/xcompact3d.f90:29-102 each), so the H2D copy of the next job and the D2H
 * copy of the previous one run on their own streams beside the kernels of the current one.  Host arrays should be
 * page-locked.  For self-starting schemes (itimescheme 1 and 5): Adams-Bashforth history is not part of a job.   */
int x3d_solver_advance_host(x3d_ctx *ctx, const double *ux_in, const double *uy_in, const double *uz_in,
                            double *ux_out, double *uy_out, double *uz_out, int nsteps);
int x3d_solver_host_sync(x3d_ctx *ctx);
/* out5 = (eek, eps, eps2, enst, divmax) as postprocess_tgv writes them      */
int x3d_solver_diagnostics_tgv(x3d_ctx *ctx, double *out5);
/* DIV U max / mean of the current velocity (divergence nlock=2)             */
int x3d_solver_divergence(x3d_ctx *ctx, double *divmax, double *divmean);

/* ---- schemes() for hosts that are not the reference's Fortran (src/schemes.f90:443-1066) ------
 * CPU-only helper: stencil scalars of one direction and one prepared LU triple.
 * which: 0 ff,fs,fw | 1 ffp.. | 2 sf,ss,sw | 3 sfp.. | 4 cfx6.. | 5 cfxp6.. | 6 cifx6.. | 7 cifxp6.. |
 *        8 cfi6.. | 9 cfip6.. | 10 cifi6.. | 11 cifip6..   (arrays 4-7 have nm entries, the others n) */
int x3d_schemes_axis(int n, int ncl1, int ncln, double len, int ifirstder, int isecondder, int ipinter,
                     double nu0nu, double cnu, x3d_deriv_coeffs *coeffs, int which, double *f, double *s,
                     double *w);

/* set_filter_coefficients (src/filters.f90:62-219) likewise: the parfiX/Y/Z scalars for filter parameter af and one
 * prepared LU triple (which: 0 fiffx,fifsx,fifwx | 1 fiffxp,fifsxp,fifwxp), n entries each                       */
int x3d_filter_axis(int n, int ncl1, int ncln, double af, x3d_filter_coeffs *coeffs, int which, double *f, double *s,
                    double *w);

/* stretching() (src/stretching.f90:96-318) likewise: out8 = yp, ypi, ppy, pp2y, pp4y, ppyi, pp2yi, pp4yi (ny entries
 * each, what x3d_set_stretching takes), alpha = the mod_stret parameter x3d_poisson_init needs; istret 1, 2, 3;
 * nym = ny - 1 with walls in y, ny when periodic                                                                  */
int x3d_stretching(int istret, double beta, double yly, int ny, int nym, double *out8, double *alpha);

#ifdef __cplusplus
}
#endif
#endif /* X3D_B200_H */
