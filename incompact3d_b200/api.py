"""Host-side mirror of the reference's operator interface (Python host).

`X3D` wraps one library context.  Its methods carry the reference's procedure
names and argument lists (src/module_param.f90:136-226, src/derive.f90,
src/filters.f90): e.g.

    x3d.derx_00(tx, ux, rx, sx, ffx, fsx, fwx, nx, ny, nz, npaire, lind)

Arrays may be numpy arrays (host: staged through the GPU, result copied back --
drop-in mode), torch CUDA tensors or raw device addresses (used in place, work is
queued on the context stream).  Every call runs the CUDA path; errors raise
`X3DError` (the reference aborts instead, src/schemes.f90:472-473).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DerivCoeffs, FilterCoeffs


class X3DError(RuntimeError):
    pass


def _addr(a, keep):
    """address of a numpy array / torch tensor / int / None"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    if isinstance(a, np.ndarray):
        if a.dtype != np.float64 and a.dtype != np.complex128:
            raise TypeError("x3d arrays must be float64/complex128")
        if not (a.flags.f_contiguous or a.flags.c_contiguous and a.ndim <= 1):
            raise ValueError("x3d arrays must be Fortran-contiguous (i fastest)")
        keep.append(a)
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):  # torch tensor
        keep.append(a)
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"unsupported array type {type(a)}")


def _coef(a, keep):
    a = np.ascontiguousarray(a, dtype=np.float64)
    keep.append(a)
    return C.c_void_p(a.ctypes.data)


_COLLOC_X = [f"{s}_{bc}" for s in ("derx", "derz", "derxx", "deryy", "derzz", "filx", "fily", "filz")
             for bc in ("00", "11", "12", "21", "22")]
_COLLOC_Y = [f"dery_{bc}" for bc in ("00", "11", "12", "21", "22")]


class X3D:
    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        if self._L.x3d_create(C.byref(h), int(device)):
            raise X3DError(self._L.x3d_last_error().decode())
        self._h = h
        self.device = device

    # -- lifecycle ------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.x3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise X3DError(self._L.x3d_last_error().decode())

    def sync(self):
        self._check(self._L.x3d_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.x3d_stream(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.x3d_launch_count(self._h))

    # -- module state -------------------------------------------------------------
    def set_deriv_coeffs(self, axis: int, c: DerivCoeffs):
        if not isinstance(c, DerivCoeffs):
            raise TypeError("set_deriv_coeffs expects incompact3d_b200.DerivCoeffs")
        self._check(self._L.x3d_set_deriv_coeffs(self._h, axis, C.byref(c)))

    def set_filter_coeffs(self, axis: int, c: FilterCoeffs):
        self._check(self._L.x3d_set_filter_coeffs(self._h, axis, C.byref(c)))

    def set_flags(self, iibm=0, istret=0, iimplicit=0, nclx=True, ncly=True, nclz=True):
        self._check(self._L.x3d_set_flags(self._h, int(iibm), int(istret), int(iimplicit),
                                          int(bool(nclx)), int(bool(ncly)), int(bool(nclz))))

    # -- operators: generic call helpers ---------------------------------------------
    def _call(self, name, arrays, coefs, ints, lind=None):
        keep = []
        args = [self._h]
        args += [_addr(a, keep) for a in arrays]
        args += [_coef(a, keep) for a in coefs]
        ints_c = [C.c_int(int(v)) for v in ints]
        args += [C.byref(v) for v in ints_c]
        if lind is not None:
            ld = C.c_double(float(lind))
            args.append(C.byref(ld))
        fn = getattr(self._L, "x3d_" + name)
        fn.restype = C.c_int
        self._check(fn(*args))


def _make_colloc(name, with_pp):
    if with_pp:
        def f(self, ty, uy, ry, sy, ffy, fsy, fwy, ppy, nx, ny, nz, npaire, lind=0.0):
            self._call(name, (ty, uy, ry, sy), (ffy, fsy, fwy, ppy), (nx, ny, nz, npaire), lind)
    else:
        def f(self, tx, ux, rx, sx, ffx, fsx, fwx, nx, ny, nz, npaire, lind=0.0):
            self._call(name, (tx, ux, rx, sx), (ffx, fsx, fwx), (nx, ny, nz, npaire), lind)
    f.__name__ = name
    f.__doc__ = f"reference procedure `{name}` (src/derive.f90 / src/filters.f90), same argument list"
    return f


for _n in _COLLOC_X:
    setattr(X3D, _n, _make_colloc(_n, False))
for _n in _COLLOC_Y:
    setattr(X3D, _n, _make_colloc(_n, True))


def _make_stag(name, ncoef, with_pp):
    def f(self, t, u, r, s, *rest):
        coefs = rest[:ncoef + (1 if with_pp else 0)]
        ints = rest[ncoef + (1 if with_pp else 0):]
        if len(ints) != 5:
            raise TypeError(f"{name}: expected 5 integer arguments, got {len(ints)}")
        self._call(name, (t, u, r, s), coefs, ints)
    f.__name__ = name
    f.__doc__ = f"reference procedure `{name}` (src/derive.f90:3796-5615), same argument list"
    return f


# (name, number of LU arrays, has ppy/ppyi argument)
for _n, _k, _pp in (("derxvp", 3, False), ("interxvp", 3, False), ("derxpv", 6, False), ("interxpv", 6, False),
                    ("interyvp", 3, False), ("deryvp", 3, True), ("interypv", 6, False), ("derypv", 6, True),
                    ("derzvp", 3, False), ("interzvp", 3, False), ("derzpv", 6, False), ("interzpv", 6, False)):
    setattr(X3D, _n, _make_stag(_n, _k, _pp))


# ---------------------------------------------------------------------------------------
# schemes() for Python hosts, Poisson solver, decomposition, device-resident solver
# ---------------------------------------------------------------------------------------
_LU_NAMES = {"d1": 0, "d1p": 1, "d2": 2, "d2p": 3, "vp": 4, "vpp": 5, "ivp": 6, "ivpp": 7,
             "pv": 8, "pvp": 9, "ipv": 10, "ipvp": 11}


class AxisSchemes:
    """Coefficients of one direction as `schemes()` builds them (src/schemes.f90): `c` holds the
    derivX/Y/Z scalars, `lu(name)` the prepared (f, s, w) arrays:
    d1/d1p = ffx../ffxp.., d2/d2p = sfx../sfxp.., vp/vpp = cfx6../cfxp6.., ivp/ivpp = cifx6../cifxp6..,
    pv/pvp = cfi6../cfip6.., ipv/ipvp = cifi6../cifip6.. .  Host-side, CPU only."""

    def __init__(self, n, ncl1, ncln, length, ifirstder=4, isecondder=4, ipinter=3, nu0nu=4.0, cnu=0.44):
        self._L = _lib.load()
        self.n, self.ncl1, self.ncln, self.length = int(n), int(ncl1), int(ncln), float(length)
        self.periodic = ncl1 == 0 and ncln == 0
        self.nm = self.n if self.periodic else self.n - 1
        self.d = self.length / self.nm
        self._opts = (int(ifirstder), int(isecondder), int(ipinter), float(nu0nu), float(cnu))
        self.c = DerivCoeffs()
        self._cache = {}
        self.lu("d1")

    def lu(self, name):
        if name not in self._cache:
            which = _LU_NAMES[name]
            m = self.nm if 4 <= which <= 7 else self.n
            f, s, w = (np.zeros(m) for _ in range(3))
            fn = self._L.x3d_schemes_axis
            fn.restype = C.c_int
            rc = fn(C.c_int(self.n), C.c_int(self.ncl1), C.c_int(self.ncln), C.c_double(self.length),
                    C.c_int(self._opts[0]), C.c_int(self._opts[1]), C.c_int(self._opts[2]),
                    C.c_double(self._opts[3]), C.c_double(self._opts[4]), C.byref(self.c), C.c_int(which),
                    C.c_void_p(f.ctypes.data), C.c_void_p(s.ctypes.data), C.c_void_p(w.ctypes.data))
            if rc:
                raise X3DError(self._L.x3d_last_error().decode())
            self._cache[name] = (f, s, w)
        return self._cache[name]

    def filter(self, af, p=False):
        """`set_filter_coefficients` (src/filters.f90:62-219) for this direction: returns (FilterCoeffs, (f, s, w)) --
        the parfiX/Y/Z scalars and fiffx,fifsx,fifwx (p=False) or fiffxp,fifsxp,fifwxp (p=True).  Host-side, CPU only."""
        c = FilterCoeffs()
        f, s, w = (np.zeros(self.n) for _ in range(3))
        fn = self._L.x3d_filter_axis
        fn.restype = C.c_int
        rc = fn(C.c_int(self.n), C.c_int(self.ncl1), C.c_int(self.ncln), C.c_double(float(af)), C.byref(c), C.c_int(1 if p else 0),
                C.c_void_p(f.ctypes.data), C.c_void_p(s.ctypes.data), C.c_void_p(w.ctypes.data))
        if rc:
            raise X3DError(self._L.x3d_last_error().decode())
        return c, (f, s, w)


def stretching(istret, beta, yly, ny, nym):
    """`stretching()` (src/stretching.f90:96-318) for hosts that are not the reference's Fortran: returns
    ({yp, ypi, ppy, pp2y, pp4y, ppyi, pp2yi, pp4yi}, alpha).  Host-side, CPU only."""
    L = _lib.load()
    out = np.zeros(8 * int(ny))
    alpha = C.c_double()
    fn = L.x3d_stretching
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    if fn(int(istret), float(beta), float(yly), int(ny), int(nym), out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(alpha)):
        raise X3DError(L.x3d_last_error().decode())
    names = ("yp", "ypi", "ppy", "pp2y", "pp4y", "ppyi", "pp2yi", "pp4yi")
    return {nm: out[q * ny:(q + 1) * ny].copy() for q, nm in enumerate(names)}, alpha.value


def _poisson_init(self, nx, ny, nz, bcx, bcy, bcz, xlx, yly, zlz, istret=0, alpha=0.0, beta=0.0):
    """decomp_2d_poisson_init (src/poisson.f90:73); nx,ny,nz are the velocity-mesh node counts"""
    p = _lib.PoissonParams(int(nx), int(ny), int(nz), int(bcx), int(bcy), int(bcz), float(xlx), float(yly), float(zlz),
                           int(istret), float(alpha), float(beta))
    fn = self._L.x3d_poisson_init
    fn.argtypes = [C.c_void_p, C.POINTER(_lib.PoissonParams)]
    self._check(fn(self._h, C.byref(p)))


def _poisson(self, rhs):
    """`poisson(rhs)` (src/poisson.f90:63): rhs is the z-pencil of the pressure mesh, solved in place"""
    keep = []
    fn = self._L.x3d_poisson
    fn.argtypes = [C.c_void_p, C.c_void_p]
    self._check(fn(self._h, _addr(rhs, keep)))


def _decomp_init(self, nx, ny, nz, p_row=1, p_col=1, rank=0, nranks=1, nccl_id=None):
    fn = self._L.x3d_decomp_init
    fn.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
    buf = None if nccl_id is None else C.create_string_buffer(bytes(nccl_id), 128)
    self._check(fn(self._h, nx, ny, nz, p_row, p_col, rank, nranks, buf))


def _decomp_info_init(self, nx, ny, nz):
    out = C.c_int()
    fn = self._L.x3d_decomp_info_init
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    self._check(fn(self._h, nx, ny, nz, C.byref(out)))
    return out.value


def _decomp_info(self, decomp_id=0):
    info = _lib.DecompInfo()
    fn = self._L.x3d_decomp_info_get
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(_lib.DecompInfo)]
    self._check(fn(self._h, decomp_id, C.byref(info)))
    return {k: list(getattr(info, k)) for k, _ in _lib.DecompInfo._fields_}


def _make_transpose(name):
    def f(self, src, dst, decomp_id=0):
        keep = []
        cplx = isinstance(src, np.ndarray) and src.dtype == np.complex128
        if hasattr(src, "is_complex") and src.is_complex():
            cplx = True
        fn = getattr(self._L, "x3d_" + name + ("_complex" if cplx else ""))
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self._check(fn(self._h, _addr(src, keep), _addr(dst, keep), int(decomp_id)))
    f.__name__ = name
    f.__doc__ = f"2DECOMP&FFT `{name}(src, dst[, decomp])`; real or complex arrays"
    return f


def _solver_init(self, nx, ny, nz, ncl=(0, 0, 0, 0, 0, 0), xlx=2 * np.pi, yly=2 * np.pi, zlz=2 * np.pi, re=1600.0,
                 dt=0.005, ifirstder=4, isecondder=4, ipinter=3, itimescheme=5, istret=0, beta=0.0, nu0nu=4.0,
                 cnu=0.44, p_row=1, p_col=1, itype=0):
    p = _lib.SolverParams(int(nx), int(ny), int(nz), *[int(v) for v in ncl], float(xlx), float(yly), float(zlz),
                          float(re), float(dt), int(ifirstder), int(isecondder), int(ipinter), int(itimescheme),
                          int(istret), float(beta), float(nu0nu), float(cnu), int(p_row), int(p_col), int(itype))
    fn = self._L.x3d_solver_init
    fn.argtypes = [C.c_void_p, C.POINTER(_lib.SolverParams)]
    self._check(fn(self._h, C.byref(p)))
    d3, z0 = (C.c_int * 3)(), C.c_int()
    f2 = self._L.x3d_solver_local_shape
    f2.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    self._check(f2(self._h, d3, C.byref(z0)))
    self._solver_shape = tuple(d3)
    self.solver_zstart = z0.value


def _solver_init_tgv(self):
    fn = self._L.x3d_solver_init_tgv
    fn.argtypes = [C.c_void_p]
    self._check(fn(self._h))


def _set_ibm_geometry(self, axis, nobjmax, npif, izap, nobj, xi, xf, nipif, nfpif, d, length, coords=None):
    """module complex_geometry of one direction (src/module_param.f90:546-556): nobj(na,nb), xi/xf(nobjmax,na,nb),
    nipif/nfpif(0:nobjmax,na,nb), Fortran order; coords = yp for axis 1"""
    keep = []

    def iarr(a):
        a = np.asfortranarray(a, dtype=np.int32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int))

    def darr(a):
        a = np.asfortranarray(a, dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double))
    nobj = np.asarray(nobj)
    na, nb = nobj.shape
    fn = self._L.x3d_set_ibm_geometry
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    fn.argtypes = [C.c_void_p] + [C.c_int] * 6 + [ip, dp, dp, ip, ip, dp, C.c_int, C.c_double, C.c_double]
    cptr = darr(coords) if coords is not None else None
    self._check(fn(self._h, int(axis), int(nobjmax), int(npif), int(izap), int(na), int(nb), iarr(nobj), darr(xi), darr(xf),
                   iarr(nipif), iarr(nfpif), cptr, 0 if coords is None else len(coords), float(d), float(length)))


def _make_lagpol(ax):
    def f(self, u, nx=None, ny=None, nz=None):
        keep = []
        shape = list(u.shape) if isinstance(u, np.ndarray) else list(reversed(u.shape))
        n = [C.c_int(int(v)) for v in (shape if nx is None else (nx, ny, nz))]
        fn = getattr(self._L, "x3d_lagpol" + ax)
        fn.argtypes = [C.c_void_p, C.c_void_p] + [C.POINTER(C.c_int)] * 3
        self._check(fn(self._h, _addr(u, keep), *[C.byref(v) for v in n]))
    f.__name__ = "lagpol" + ax
    f.__doc__ = f"reference procedure `lagpol{ax}(u)` (src/ibm.f90): u is rebuilt inside the bodies, in place"
    return f


def _set_ibm_analytic(self, axis, ana_i=None, ana_f=None):
    """analytic wall positions for iibm = 3 with ianal /= 0 (what analitic_x / analitic_y return for xi / xf,
    src/ibm.f90:412-417); None, None switches back to ianal = 0"""
    fn = self._L.x3d_set_ibm_analytic
    dp = C.POINTER(C.c_double)
    fn.argtypes = [C.c_void_p, C.c_int, dp, dp]
    if ana_i is None or ana_f is None:
        self._check(fn(self._h, int(axis), None, None))
        return
    ai, af = np.asfortranarray(ana_i, dtype=np.float64), np.asfortranarray(ana_f, dtype=np.float64)
    self._check(fn(self._h, int(axis), ai.ctypes.data_as(dp), af.ctypes.data_as(dp)))


def _make_cubspl(ax):
    def f(self, u, lind, nx=None, ny=None, nz=None):
        keep = []
        shape = list(u.shape) if isinstance(u, np.ndarray) else list(reversed(u.shape))
        n = [C.c_int(int(v)) for v in (shape if nx is None else (nx, ny, nz))]
        ld = C.c_double(float(lind))
        fn = getattr(self._L, "x3d_cubspl" + ax)
        fn.argtypes = [C.c_void_p, C.c_void_p] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_double)]
        self._check(fn(self._h, _addr(u, keep), *[C.byref(v) for v in n], C.byref(ld)))
    f.__name__ = "cubspl" + ax
    f.__doc__ = f"reference procedure `cubspl{ax}(u,lind)` (src/ibm.f90): u is rebuilt inside the bodies by cubic splines, in place"
    return f


def _solver_init_channel(self):
    fn = self._L.x3d_solver_init_channel
    fn.argtypes = [C.c_void_p]
    self._check(fn(self._h))


def _solver_step(self, nsteps=1):
    fn = self._L.x3d_solver_step
    fn.argtypes = [C.c_void_p, C.c_int]
    self._check(fn(self._h, int(nsteps)))


def _solver_diag(self):
    """-> dict(eek, eps, eps2, enst, divmax) as postprocess_tgv / divergence(nlock=2) report them"""
    out = (C.c_double * 5)()
    fn = self._L.x3d_solver_diagnostics_tgv
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    self._check(fn(self._h, out))
    return dict(eek=out[0], eps=out[1], eps2=out[2], enst=out[3], divmax=out[4])


def _solver_divergence(self):
    a, b = C.c_double(), C.c_double()
    fn = self._L.x3d_solver_divergence
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    self._check(fn(self._h, C.byref(a), C.byref(b)))
    return a.value, b.value


def _solver_set_velocity(self, ux, uy, uz):
    keep = []
    fn = self._L.x3d_solver_set_velocity
    fn.argtypes = [C.c_void_p] * 4
    self._check(fn(self._h, _addr(ux, keep), _addr(uy, keep), _addr(uz, keep)))


def _solver_get_velocity(self, ux=None, uy=None, uz=None):
    if ux is None:
        ux, uy, uz = (np.zeros(self._solver_shape, order="F") for _ in range(3))
    keep = []
    fn = self._L.x3d_solver_get_velocity
    fn.argtypes = [C.c_void_p] * 4
    self._check(fn(self._h, _addr(ux, keep), _addr(uy, keep), _addr(uz, keep)))
    return ux, uy, uz


def _solver_set_case(self, cpg=0, wrotation=0.0, spinup_time=0, iin=0, u1=1.0, u2=1.0, inflow_noise=0.0, iibm=0, ubc=(0.0, 0.0, 0.0)):
    """case parameters beyond mesh and schemes: channel forcing (src/Case-Channel.f90:396-420), cylinder inflow / outflow
    (src/Case-Cylinder-wake.f90:100-203), immersed boundary (iibm, body velocity)"""
    c = _lib.CaseParams(int(cpg), float(wrotation), int(spinup_time), int(iin), float(u1), float(u2), float(inflow_noise), int(iibm),
                        float(ubc[0]), float(ubc[1]), float(ubc[2]))
    fn = self._L.x3d_solver_set_case
    fn.argtypes = [C.c_void_p, C.POINTER(_lib.CaseParams)]
    self._check(fn(self._h, C.byref(c)))


def _solver_set_ibm_mask(self, ep1):
    keep = []
    fn = self._L.x3d_solver_set_ibm_mask
    fn.argtypes = [C.c_void_p, C.c_void_p]
    self._check(fn(self._h, _addr(ep1, keep)))


def _solver_set_inflow_noise(self, bxo=None, byo=None, bzo=None):
    keep = []
    fn = self._L.x3d_solver_set_inflow_noise
    fn.argtypes = [C.c_void_p] * 4
    self._check(fn(self._h, _addr(bxo, keep), _addr(byo, keep), _addr(bzo, keep)))


def _solver_wall_velocity_x(self, planes=None):
    """set (planes = 6 arrays or None entries) or get (planes = None -> list of 6 arrays) bxx1 bxy1 bxz1 bxxn bxyn bxzn"""
    keep = []
    arr = (C.c_void_p * 6)()
    if planes is None:
        out = [np.zeros((self._solver_shape[1], self._solver_shape[2]), order="F") for _ in range(6)]
        for q in range(6):
            arr[q] = out[q].ctypes.data
        fn = self._L.x3d_solver_get_wall_velocity_x
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        self._check(fn(self._h, arr))
        return out
    for q in range(6):
        a = _addr(planes[q], keep)
        arr[q] = a.value if a is not None else None
    fn = self._L.x3d_solver_set_wall_velocity_x
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    self._check(fn(self._h, arr))


def _solver_apply_spatial_filter(self, ifilter=1, af=0.45):
    """reference procedure `apply_spatial_filter(ux1,uy1,uz1,phi1)` (src/tools.f90:600-675) on the solver's velocity"""
    fn = self._L.x3d_solver_apply_spatial_filter
    fn.argtypes = [C.c_void_p, C.c_int, C.c_double]
    self._check(fn(self._h, int(ifilter), float(af)))


def _solver_init_cyl(self):
    fn = self._L.x3d_solver_init_cyl
    fn.argtypes = [C.c_void_p]
    self._check(fn(self._h))


def _solver_advance_host(self, vin, vout, nsteps=1):
    """queue one job: H2D of the host velocity `vin` (3 arrays), nsteps time steps, D2H into `vout`; asynchronous"""
    keep = []
    fn = self._L.x3d_solver_advance_host
    fn.argtypes = [C.c_void_p] * 7 + [C.c_int]
    self._check(fn(self._h, *[_addr(a, keep) for a in vin], *[_addr(a, keep) for a in vout], int(nsteps)))


def _decomp_stats(self):
    """-> (bytes sent to other ranks by transposes since decomp_init, transposed fields)"""
    a, b = C.c_ulonglong(), C.c_ulonglong()
    fn = self._L.x3d_decomp_stats
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    self._check(fn(self._h, C.byref(a), C.byref(b)))
    return a.value, b.value


def _transpose_selftest(self, decomp_id=0, whiches=(0, 1, 2, 3), modes=(-1, 0, 1, 2), kinds=(0, 1)):
    """bit-exactness of the production transposes between library-owned pencils (collective); -> number of wrong elements.
    kinds: 0 real, 1 complex"""
    fn = self._L.x3d_transpose_selftest
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    bad = 0
    for which in whiches:
        for cplx in kinds:
            for mode in modes:
                m = C.c_longlong()
                self._check(fn(self._h, int(which), int(decomp_id), cplx, int(mode), C.byref(m)))
                bad += m.value
    return bad


def _solver_host_sync(self):
    fn = self._L.x3d_solver_host_sync
    fn.argtypes = [C.c_void_p]
    self._check(fn(self._h))


def _profile_step(self, nsteps=1):
    """run `nsteps` solver steps with per-launch CUDA-event timing; -> list of kernel classes"""
    import json
    self._L.x3d_profile_begin.argtypes = [C.c_void_p]
    self._L.x3d_profile_end.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    self._check(self._L.x3d_profile_begin(self._h))
    self.solver_step(nsteps)
    buf = C.create_string_buffer(1 << 16)
    self._check(self._L.x3d_profile_end(self._h, buf, len(buf)))
    out = json.loads(buf.value.decode())
    for r in out:
        r["count"] //= nsteps
        r["total_ms"] /= nsteps
    return out


X3D.profile_step = _profile_step
X3D.poisson_init = _poisson_init
X3D.poisson = _poisson
X3D.decomp_init = _decomp_init
X3D.decomp_info_init = _decomp_info_init
X3D.decomp_info = _decomp_info
for _n in ("transpose_x_to_y", "transpose_y_to_z", "transpose_z_to_y", "transpose_y_to_x"):
    setattr(X3D, _n, _make_transpose(_n))
X3D.solver_init = _solver_init
X3D.solver_init_tgv = _solver_init_tgv
X3D.solver_init_channel = _solver_init_channel
X3D.set_ibm_geometry = _set_ibm_geometry
X3D.set_ibm_analytic = _set_ibm_analytic
for _ax in "xyz":
    setattr(X3D, "lagpol" + _ax, _make_lagpol(_ax))
    setattr(X3D, "cubspl" + _ax, _make_cubspl(_ax))
X3D.solver_step = _solver_step
X3D.solver_diagnostics_tgv = _solver_diag
X3D.solver_divergence = _solver_divergence
X3D.solver_set_velocity = _solver_set_velocity
X3D.solver_get_velocity = _solver_get_velocity
X3D.solver_advance_host = _solver_advance_host
X3D.solver_set_case = _solver_set_case
X3D.solver_set_ibm_mask = _solver_set_ibm_mask
X3D.solver_set_inflow_noise = _solver_set_inflow_noise
X3D.solver_wall_velocity_x = _solver_wall_velocity_x
X3D.solver_init_cyl = _solver_init_cyl
X3D.solver_apply_spatial_filter = _solver_apply_spatial_filter
X3D.solver_host_sync = _solver_host_sync
X3D.decomp_stats = _decomp_stats
X3D.transpose_selftest = _transpose_selftest


# ---------------------------------------------------------------------------------------
# decomposition helpers (CPU only) and the two halves of a transpose
# ---------------------------------------------------------------------------------------
TRANSPOSES = {"x_to_y": 0, "y_to_z": 1, "z_to_y": 2, "y_to_x": 3}


def decomp_compute(nx, ny, nz, p_row, p_col, rank):
    """pencil extents of `rank` (1-based inclusive starts/ends like TYPE(DECOMP_INFO)); no GPU needed"""
    L = _lib.load()
    info = _lib.DecompInfo()
    L.x3d_decomp_compute.argtypes = [C.c_int] * 6 + [C.POINTER(_lib.DecompInfo)]
    if L.x3d_decomp_compute(nx, ny, nz, p_row, p_col, rank, C.byref(info)):
        raise X3DError(L.x3d_last_error().decode())
    return {k: list(getattr(info, k)) for k, _ in _lib.DecompInfo._fields_}


def transpose_plan(nx, ny, nz, p_row, p_col, rank, which):
    """all-to-all(v) plan of one transpose for one rank; no GPU needed"""
    L = _lib.load()
    w = TRANSPOSES[which] if isinstance(which, str) else int(which)
    cap = max(p_row, p_col)
    npeers = C.c_int()
    peers = (C.c_int * cap)()
    arrs = [(C.c_longlong * cap)() for _ in range(4)]
    sd, rd = (C.c_int * 3)(), (C.c_int * 3)()
    L.x3d_transpose_plan.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.POINTER(C.c_longlong)] * 4 + [C.POINTER(C.c_int)] * 2
    if L.x3d_transpose_plan(nx, ny, nz, p_row, p_col, rank, w, C.byref(npeers), peers, *arrs, sd, rd):
        raise X3DError(L.x3d_last_error().decode())
    n = npeers.value
    return dict(peers=list(peers[:n]), scount=list(arrs[0][:n]), sdispl=list(arrs[1][:n]), rcount=list(arrs[2][:n]),
                rdispl=list(arrs[3][:n]), send_dims=list(sd), recv_dims=list(rd))


def nccl_unique_id() -> bytes:
    L = _lib.load()
    buf = C.create_string_buffer(128)
    L.x3d_nccl_unique_id.argtypes = [C.c_char_p]
    if L.x3d_nccl_unique_id(buf):
        raise X3DError(L.x3d_last_error().decode())
    return buf.raw


def _transpose_pack(self, which, src, packed, decomp_id=0, is_complex=False):
    keep = []
    fn = self._L.x3d_transpose_pack
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    self._check(fn(self._h, TRANSPOSES[which], _addr(src, keep), _addr(packed, keep), decomp_id, int(is_complex)))


def _transpose_unpack(self, which, packed, dst, decomp_id=0, is_complex=False):
    keep = []
    fn = self._L.x3d_transpose_unpack
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    self._check(fn(self._h, TRANSPOSES[which], _addr(packed, keep), _addr(dst, keep), decomp_id, int(is_complex)))


X3D.transpose_pack = _transpose_pack
X3D.transpose_unpack = _transpose_unpack
