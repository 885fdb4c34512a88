"""Shared helpers of the parity tests: call one reference-named operator through the
product C ABI (incompact3d_b200.X3D) and through the oracle with the same inputs."""
import ctypes as C
import re

import numpy as np

import oracle_lib as ol


def to_prod(struct, cls):
    out = cls()
    C.memmove(C.byref(out), C.byref(struct), C.sizeof(cls))
    return out


def parse(name):
    """-> (family, axis letter, bc or None) ; family in d1,d2,fil,dvp,ivp,dpv,ipv"""
    m = re.match(r"^(der|fil)([xyz])(\2?)_(\d\d)$", name)
    if m:
        fam = "fil" if m.group(1) == "fil" else ("d2" if m.group(3) else "d1")
        return fam, m.group(2), m.group(4)
    m = re.match(r"^(der|inter)([xyz])(vp|pv)$", name)
    fam = ("d" if m.group(1) == "der" else "i") + m.group(3)
    return fam, m.group(2), None


def lu_arrays(A, fam, npaire):
    """the LU arrays the reference call sites pass for this operator / npaire"""
    p = "p" if npaire == 1 else ""
    if fam == "d1":
        return [A.arr("ff" + p), A.arr("fs" + p), A.arr("fw" + p)]
    if fam == "d2":
        return [A.arr("sf" + p), A.arr("ss" + p), A.arr("sw" + p)]
    if fam == "fil":
        return [A.arr("fiff" + p), A.arr("fifs" + p), A.arr("fifw" + p)]
    if fam == "dvp":
        return [A.arr("cfx6"), A.arr("csx6"), A.arr("cwx6")]
    if fam == "ivp":
        return [A.arr("cifxp6"), A.arr("cisxp6"), A.arr("ciwxp6")]
    if fam == "dpv":
        return [A.arr("cfip6"), A.arr("csip6"), A.arr("cwip6"), A.arr("cfx6"), A.arr("csx6"), A.arr("cwx6")]
    return [A.arr("cifip6"), A.arr("cisip6"), A.arr("ciwip6"), A.arr("cifx6"), A.arr("cisx6"), A.arr("ciwx6")]


def oracle_op(name, u, A, npaire, post=None, rhs_only=False):
    fam, ax, bc = parse(name)
    lu = lu_arrays(A, fam, npaire)
    if fam in ("dpv", "ipv"):
        lu = lu[3:] if A.periodic else lu[:3]
    return ol.op(name, u, *lu, c=A.c, fc=A.fc, npaire=npaire, post=post, periodic=A.periodic, rhs_only=rhs_only)


def configure(x3d, A, axis, istret=0, iimplicit=0):
    from incompact3d_b200 import DerivCoeffs, FilterCoeffs
    x3d.set_deriv_coeffs(axis, to_prod(A.c, DerivCoeffs))
    x3d.set_filter_coeffs(axis, to_prod(A.fc, FilterCoeffs))
    ncl = [True, True, True]
    ncl[axis] = A.periodic
    x3d.set_flags(iibm=0, istret=istret, iimplicit=iimplicit, nclx=ncl[0], ncly=ncl[1], nclz=ncl[2])


def product_op(x3d, name, u, A, npaire, post=None, t=None, lind=0.0):
    """call x3d.<name> with the reference argument list; u,t numpy (host) or torch (device)"""
    fam, ax, bc = parse(name)
    axis = "xyz".index(ax)
    lu = lu_arrays(A, fam, npaire)
    shape = list(u.shape) if isinstance(u, np.ndarray) else list(reversed(u.shape))
    n, nm = A.n, A.nm
    if fam in ("dvp", "ivp"):
        shape[axis] = nm
    elif fam in ("dpv", "ipv"):
        shape[axis] = n
    if t is None:
        if isinstance(u, np.ndarray):
            t = np.full(shape, -777.0, order="F")
        else:
            import torch
            t = torch.full(tuple(reversed(shape)), -777.0, dtype=torch.float64, device=u.device)
    dims = list(u.shape) if isinstance(u, np.ndarray) else list(reversed(u.shape))
    nx, ny, nz = dims
    fn = getattr(x3d, name)
    if fam in ("d1", "d2", "fil"):
        if fam == "d1" and ax == "y":
            pp = post if post is not None else np.ones(ny)
            fn(t, u, None, None, *lu, pp, nx, ny, nz, npaire, lind)
        else:
            fn(t, u, None, None, *lu, nx, ny, nz, npaire, lind)
    else:
        vel = [nx, ny, nz]
        if fam in ("dpv", "ipv"):
            vel[axis] = n
        # integer argument orders of the reference (src/derive.f90:3796-5615)
        if ax == "x":
            ints = [vel[0], nm, vel[1], vel[2]] if fam in ("dvp", "ivp") else [nm, vel[0], vel[1], vel[2]]
        elif ax == "y":
            ints = [vel[0], vel[1], nm, vel[2]] if fam in ("dvp", "ivp") else [vel[0], nm, vel[1], vel[2]]
        else:
            ints = [vel[0], vel[1], vel[2], nm] if fam in ("dvp", "ivp") else [vel[0], vel[1], nm, vel[2]]
        extra = []
        if name in ("deryvp", "derypv"):
            extra = [post if post is not None else np.ones(nm if name == "deryvp" else n)]
        fn(t, u, None, None, *lu, *extra, *ints, npaire)
    return t


def rel_linf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
