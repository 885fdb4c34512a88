"""ctypes binding of the CPU oracle (oracle/libx3d_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libx3d_oracle.so")

_dp = C.POINTER(C.c_double)


def _names(cls):
    return [f[0] for f in cls._fields_]


class DerivCoeffs(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "alfa1 af1 bf1 cf1 df1 alfa2 af2 alfan afn bfn cfn dfn alfam afm alfai afi bfi "
        "alsa1 as1 bs1 cs1 ds1 alsa2 as2 alsan asn bsn csn dsn alsam asm_ "
        "alsa3 as3 bs3 alsat ast bst alsa4 as4 bs4 cs4 alsatt astt bstt cstt "
        "alsai asi bsi csi dsi alcai6 aci6 bci6 ailcai6 aici6 bici6 cici6 dici6").split()]


class FilterCoeffs(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "fial1 fia1 fib1 fic1 fid1 fial2 fia2 fib2 fic2 fid2 fial3 fia3 fib3 fic3 fid3 fie3 fif3 "
        "fialn fian fibn ficn fidn fialm fiam fibm ficm fidm fialp fiap fibp ficp fidp fiep fifp "
        "fiali fiai fibi fici fidi").split()]


def build(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.x3do_last_error.restype = C.c_char_p
        L.x3do_axis_create.restype = C.c_void_p
        L.x3do_axis_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                       C.c_double, C.c_double]
        L.x3do_axis_destroy.argtypes = [C.c_void_p]
        L.x3do_axis_set_filter.argtypes = [C.c_void_p, C.c_double]
        L.x3do_axis_nm.argtypes = [C.c_void_p]
        L.x3do_axis_d.argtypes = [C.c_void_p]
        L.x3do_axis_d.restype = C.c_double
        L.x3do_axis_get_array.argtypes = [C.c_void_p, C.c_char_p, _dp, C.c_int]
        L.x3do_axis_get_coeffs.argtypes = [C.c_void_p, C.POINTER(DerivCoeffs), C.POINTER(FilterCoeffs)]
        L.x3do_op.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                              C.POINTER(DerivCoeffs), C.POINTER(FilterCoeffs), C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class Axis:
    """schemes() for one direction: coefficient scalars + LU arrays (src/schemes.f90)."""

    def __init__(self, n, ncl1, ncln, length, ifirstder=4, isecondder=4, ipinter=3, nu0nu=4.0, cnu=0.44, af=None):
        L = lib()
        self.h = L.x3do_axis_create(n, ncl1, ncln, float(length), ifirstder, isecondder, ipinter, nu0nu, cnu)
        if not self.h:
            raise RuntimeError(L.x3do_last_error().decode())
        self.n, self.ncl1, self.ncln = n, ncl1, ncln
        self.periodic = ncl1 == 0 and ncln == 0
        self.nm = L.x3do_axis_nm(self.h)
        self.d = L.x3do_axis_d(self.h)
        if af is not None:
            L.x3do_axis_set_filter(self.h, float(af))
        self.c = DerivCoeffs()
        self.fc = FilterCoeffs()
        L.x3do_axis_get_coeffs(self.h, C.byref(self.c), C.byref(self.fc))
        self._cache = {}

    def arr(self, name):
        if name not in self._cache:
            L = lib()
            n = L.x3do_axis_get_array(self.h, name.encode(), None, 0)
            if n < 0:
                raise KeyError(name)
            out = np.zeros(n)
            L.x3do_axis_get_array(self.h, name.encode(), _p(out), n)
            self._cache[name] = out
        return self._cache[name]

    def __del__(self):
        try:
            lib().x3do_axis_destroy(self.h)
        except Exception:
            pass


def op(name, u, f, s, w, c=None, fc=None, npaire=1, post=None, periodic=False, rhs_only=False, out=None):
    """Apply reference operator `name` (e.g. 'derx_11', 'interyvp') to the Fortran-ordered array u."""
    L = lib()
    u = np.asfortranarray(u, dtype=np.float64)
    axis = "xyz".index([ch for ch in name[3:] if ch in "xyz"][0]) if not name.startswith("inter") else "xyz".index(name[5])
    dims = list(u.shape)
    if name.endswith("vp") and not periodic:
        dims[axis] -= 1
    elif name.endswith("pv") and not periodic:
        dims[axis] += 1
    t = np.full(dims, -777.0, order="F") if out is None else out
    din = (C.c_int * 3)(*u.shape)
    rc = L.x3do_op(name.encode(), din, int(npaire), _p(u), _p(t), _p(np.ascontiguousarray(f)),
                   _p(np.ascontiguousarray(s)), _p(np.ascontiguousarray(w)),
                   _p(None if post is None else np.ascontiguousarray(post)),
                   C.byref(c) if c is not None else None, C.byref(fc) if fc is not None else None,
                   int(bool(periodic)), int(bool(rhs_only)))
    if rc:
        raise RuntimeError(L.x3do_last_error().decode())
    return t
