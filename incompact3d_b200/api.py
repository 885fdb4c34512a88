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
