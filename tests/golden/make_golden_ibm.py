#!/usr/bin/env python
"""Golden vectors for the immersed-boundary pre-pass of the operators, from the reference source.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_ibm.py

Executes the reference's own statements (via f90mini.py) of src/ibm.f90 lagpolx / lagpoly / lagpolz (with polint)
-- the Lagrange reconstruction of the field inside solid bodies that derx/dery/derz and derxx/deryy/derzz run on
their input when iibm = 2 (src/derive.f90:23) -- on a seeded field with synthetic body intervals (immersed and
touching the boundaries, one or two bodies per line, fewer fluid points than npif next to a face), for izap 0 / 1,
on a uniform mesh and, for y, on a stretched one.  Writes tests/golden/ibm.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402
from make_golden_poisson import load, sub_text  # noqa: E402

SEED = 20261020
NX, NY, NZ = 26, 24, 22
LEN = (5.0, 4.0, 3.0)
NOBJMAX = 2


def geometry(rng, n_line, na, nb, length, coords, npif):
    """bodies along the lines of one direction: nobj(na,nb), xi/xf(nobjmax,na,nb), nipif/nfpif(0:nobjmax,na,nb)"""
    nobj = np.zeros((na, nb), dtype=np.int64)
    xi = np.zeros((NOBJMAX, na, nb)); xf = np.zeros((NOBJMAX, na, nb))
    nip = np.zeros((NOBJMAX + 1, na, nb), dtype=np.int64); nfp = np.zeros((NOBJMAX + 1, na, nb), dtype=np.int64)
    for b in range(nb):
        for a in range(na):
            kind = rng.integers(0, 6)
            if kind == 0:
                continue
            if kind in (1, 2):          # one immersed body
                lo = rng.uniform(0.25, 0.45) * length; hi = lo + rng.uniform(0.1, 0.25) * length
                segs = [(lo, hi)]
            elif kind == 3:             # two bodies
                lo = rng.uniform(0.2, 0.3) * length; hi = lo + 0.1 * length
                lo2 = rng.uniform(0.6, 0.7) * length; hi2 = lo2 + 0.12 * length
                segs = [(lo, hi), (lo2, hi2)]
            elif kind == 4:             # body touching the first boundary (semi-immersed)
                segs = [(0.0, rng.uniform(0.1, 0.2) * length)]
            else:                       # body touching the last boundary
                segs = [(rng.uniform(0.8, 0.9) * length, length)]
            nobj[a, b] = len(segs)
            for i, (lo, hi) in enumerate(segs):
                xi[i, a, b] = lo; xf[i, a, b] = hi
                # fluid points available on each side (the reference's genepsi3d counts them); keep within the mesh
                left = int(np.searchsorted(coords, lo)) - 1
                right = n_line - int(np.searchsorted(coords, hi, side="right")) - 1
                nip[i + 1, a, b] = max(0, min(npif, left - 1))
                nfp[i + 1, a, b] = max(0, min(npif, right - 1))
                if len(segs) == 2 and i == 0:
                    nfp[i + 1, a, b] = min(nfp[i + 1, a, b], 1)
                if len(segs) == 2 and i == 1:
                    nip[i + 1, a, b] = min(nip[i + 1, a, b], 1)
    return nobj, xi, xf, nip, nfp


def main():
    ibm, mp = load("ibm"), load("module_param")
    out = {}
    rng = np.random.default_rng(SEED)
    u0 = np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ)))
    out["u"] = u0
    d = [LEN[0] / (NX - 1), LEN[1] / (NY - 1), LEN[2] / (NZ - 1)]
    yp_uniform = np.arange(NY) * d[1]
    eta = np.linspace(0, 1, NY)
    yp_stretched = LEN[1] * (0.5 * (1 - np.cos(np.pi * eta)))    # any monotone node distribution
    arrays = {"u", "xsize", "ysize", "zsize", "nobjx", "nobjy", "nobjz", "xi", "xf", "yi", "yf", "zi", "zf", "nxipif", "nxfpif", "nyipif",
              "nyfpif", "nzipif", "nzfpif", "yp", "xa", "ya", "c", "d"}
    tr = fm.Transpiler(arrays_hint=arrays)
    for npif in (2,):
        for izap in (1, 0):
            for ax in "xyz":
                for stretched in ((False, True) if ax == "y" else (False,)):
                    ns = fm.base_namespace()
                    ns.update(fm.module_parameters(mp))
                    yp = yp_stretched if stretched else yp_uniform
                    ns.update(dict(nx=NX, ny=NY, nz=NZ, dx=d[0], dy=d[1], dz=d[2], xlx=LEN[0], yly=LEN[1], zlz=LEN[2], npif=npif, izap=izap,
                                   xsize=fm.FArr(np.array([NX, NY, NZ])), ysize=fm.FArr(np.array([NX, NY, NZ])),
                                   zsize=fm.FArr(np.array([NX, NY, NZ])), yp=fm.FArr(yp.copy()), nmax=30))  # ibm.f90:351
                    axis = "xyz".index(ax)
                    n_line = (NX, NY, NZ)[axis]
                    na, nb = [(NY, NZ), (NX, NZ), (NX, NY)][axis]
                    coords = [np.arange(NX) * d[0], yp, np.arange(NZ) * d[2]][axis]
                    nobj, xi, xf, nip, nfp = geometry(rng, n_line, na, nb, LEN[axis], coords, npif)
                    ns[f"nobj{ax}"] = fm.FArr(nobj)
                    ns[f"{ax}i"] = fm.FArr(xi); ns[f"{ax}f"] = fm.FArr(xf)
                    ns[f"n{ax}ipif"] = fm.FArr(nip, lb=(0, 1, 1)); ns[f"n{ax}fpif"] = fm.FArr(nfp, lb=(0, 1, 1))
                    _, code = tr.subroutine(sub_text(ibm, "polint"))
                    # polint returns y, dy through its dummies: make the transpiled function hand them back
                    exec(code, ns)
                    pol = ns["polint"]

                    def polint_call(xa, ya, n, x, y, dy, _pol=pol):
                        loc = _pol(xa, ya, n, x, y, dy)
                        return loc["y"], loc["dy"]
                    _, code = tr.subroutine(sub_text(ibm, f"lagpol{ax}"))
                    code = code.replace("pass  # call polint", "ypol, dypol = polint_call(xa, ya, na, xpol, 0.0, 0.0)")
                    ns["polint_call"] = polint_call
                    exec(code, ns)
                    u = fm.FArr(u0.copy(order="F"))
                    ns[f"lagpol{ax}"](u)
                    tag = f"{ax}/izap{izap}/st{int(stretched)}"
                    out[f"{tag}/out"] = u.a.copy()
                    out[f"{tag}/nobj"] = nobj; out[f"{tag}/xi"] = xi; out[f"{tag}/xf"] = xf
                    out[f"{tag}/nipif"] = nip; out[f"{tag}/nfpif"] = nfp
                    out[f"{tag}/coords"] = coords
                    changed = int((u.a != u0).sum())
                    print(tag, "points rebuilt:", changed)
                    assert changed > 0
    out["meta/n"] = np.array([NX, NY, NZ]); out["meta/len"] = np.array(LEN); out["meta/npif"] = np.int64(2)
    out["meta/nobjmax"] = np.int64(NOBJMAX)
    np.savez_compressed(os.path.join(HERE, "ibm.npz"), **out)
    print("ibm.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
