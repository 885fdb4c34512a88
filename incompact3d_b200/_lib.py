"""ctypes binding of libx3d_b200.so (the C ABI declared in include/x3d_b200.h).

Fails loudly when the CUDA library is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libx3d_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class DerivCoeffs(C.Structure):
    """x3d_deriv_coeffs -- members of modules derivX/Y/Z (src/module_param.f90:559-617)."""
    _fields_ = [(n, C.c_double) for n in (
        "alfa1 af1 bf1 cf1 df1 alfa2 af2 alfan afn bfn cfn dfn alfam afm alfai afi bfi "
        "alsa1 as1 bs1 cs1 ds1 alsa2 as2 alsan asn bsn csn dsn alsam asm_ "
        "alsa3 as3 bs3 alsat ast bst alsa4 as4 bs4 cs4 alsatt astt bstt cstt "
        "alsai asi bsi csi dsi alcai6 aci6 bci6 ailcai6 aici6 bici6 cici6 dici6").split()]


class FilterCoeffs(C.Structure):
    """x3d_filter_coeffs -- members of modules parfiX/Y/Z (src/module_param.f90:621-656)."""
    _fields_ = [(n, C.c_double) for n in (
        "fial1 fia1 fib1 fic1 fid1 fial2 fia2 fib2 fic2 fid2 fial3 fia3 fib3 fic3 fid3 fie3 fif3 "
        "fialn fian fibn ficn fidn fialm fiam fibm ficm fidm fialp fiap fibp ficp fidp fiep fifp "
        "fiali fiai fibi fici fidi").split()]


class DecompInfo(C.Structure):
    _fields_ = [(n, C.c_int * 3) for n in ("xst", "xen", "xsz", "yst", "yen", "ysz", "zst", "zen", "zsz")]


class PoissonParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("bcx", C.c_int), ("bcy", C.c_int), ("bcz", C.c_int),
                ("xlx", C.c_double), ("yly", C.c_double), ("zlz", C.c_double),
                ("istret", C.c_int), ("alpha", C.c_double), ("beta", C.c_double)]


class SolverParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("nclx1", C.c_int), ("nclxn", C.c_int), ("ncly1", C.c_int), ("nclyn", C.c_int),
                ("nclz1", C.c_int), ("nclzn", C.c_int),
                ("xlx", C.c_double), ("yly", C.c_double), ("zlz", C.c_double),
                ("re", C.c_double), ("dt", C.c_double),
                ("ifirstder", C.c_int), ("isecondder", C.c_int), ("ipinter", C.c_int), ("itimescheme", C.c_int),
                ("istret", C.c_int), ("beta", C.c_double), ("nu0nu", C.c_double), ("cnu", C.c_double),
                ("p_row", C.c_int), ("p_col", C.c_int), ("itype", C.c_int)]


class CaseParams(C.Structure):
    _fields_ = [("cpg", C.c_int), ("wrotation", C.c_double), ("spinup_time", C.c_int), ("iin", C.c_int),
                ("u1", C.c_double), ("u2", C.c_double), ("inflow_noise", C.c_double),
                ("iibm", C.c_int), ("ubcx", C.c_double), ("ubcy", C.c_double), ("ubcz", C.c_double)]


_lib = None


def load():
    """Load the CUDA library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m incompact3d_b200.build` "
            "(or __graft_entry__.build()).  incompact3d_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.x3d_last_error.restype = C.c_char_p
    L.x3d_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.x3d_destroy.argtypes = [C.c_void_p]
    L.x3d_launch_count.argtypes = [C.c_void_p]
    L.x3d_launch_count.restype = C.c_longlong
    L.x3d_sync.argtypes = [C.c_void_p]
    L.x3d_stream.argtypes = [C.c_void_p]
    L.x3d_stream.restype = C.c_ulonglong
    L.x3d_set_deriv_coeffs.argtypes = [C.c_void_p, C.c_int, C.POINTER(DerivCoeffs)]
    L.x3d_set_filter_coeffs.argtypes = [C.c_void_p, C.c_int, C.POINTER(FilterCoeffs)]
    L.x3d_set_flags.argtypes = [C.c_void_p] + [C.c_int] * 6
    _lib = L
    return L


def symbols_declared_in_header():
    """names of all functions declared in include/x3d_b200.h (for the export test)"""
    import re
    hdr = os.path.join(HERE, "..", "include", "x3d_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    stems = re.findall(r"X3D_DECL_DER[XY]\((\w+)\)", txt)
    txt = re.sub(r"#define[^\n]*(\\\n[^\n]*)*", "", txt)  # drop macro bodies
    names = set(re.findall(r"\b(x3d_[a-z0-9_]+)\s*\(", txt))
    for stem in stems:
        if stem != "name":
            names.add("x3d_" + stem)
    names.discard("x3d_")
    return sorted(n for n in names if not n.endswith("_"))
