#!/usr/bin/env python
"""Golden vectors for the cubic-spline immersed-boundary pre-pass (iibm = 3), from the reference source.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_cubspl.py

Executes the reference's own statements (via f90mini.py) of src/ibm.f90 cubsplx / cubsply / cubsplz with cubic_spline
on a seeded field with synthetic body intervals (immersed bodies, two bodies per line with few fluid points between
them, bodies that start or end on the domain boundary and get ghost points), for izap 0 / 1, wall values lind = 0 and
0.7, on a uniform mesh and, for y, a stretched one; x and y also with analytic wall positions (ianal = 1: the results of
analitic_x / analitic_y are supplied as arrays).  Writes tests/golden/ibm_cubspl.npz.

Two statements are replaced in the transpiled text, both documented here because they change nothing numerically:
 * `call analitic_x/y(...)` (case geometry, out of scope) -> a lookup in the supplied array of analytic positions;
 * `(...)**2`, `(...)**3` -> repeated multiplication, which is what the Fortran compiler emits for integer powers
   (Python's float power goes through pow()).
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402
from make_golden_ibm import LEN, NOBJMAX, NX, NY, NZ, geometry  # noqa: E402
from make_golden_poisson import load, sub_text  # noqa: E402

SEED = 20261022


def ipow(x, n):
    r = x
    for _ in range(n - 1):
        r = r * x
    return r


def main():
    ibm, mp = load("ibm"), load("module_param")
    out = {}
    rng = np.random.default_rng(SEED)
    u0 = np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ)))
    out["u"] = u0
    d = [LEN[0] / (NX - 1), LEN[1] / (NY - 1), LEN[2] / (NZ - 1)]
    yp_uniform = np.arange(NY) * d[1]
    eta = np.linspace(0, 1, NY)
    yp_stretched = LEN[1] * (0.5 * (1 - np.cos(np.pi * eta)))
    arrays = {"u", "xsize", "ysize", "zsize", "nobjx", "nobjy", "nobjz", "xi", "xf", "yi", "yf", "zi", "zf", "nxipif", "nxfpif", "nyipif",
              "nyfpif", "nzipif", "nzfpif", "yp", "xa", "ya", "xaa", "yaa", "xx", "alpha", "cc", "zz", "ll", "aa", "yy", "hh", "dd", "bb",
              "mm", "ana_i", "ana_f"}
    tr = fm.Transpiler(arrays_hint=arrays)
    _, spline_code = tr.subroutine(sub_text(ibm, "cubic_spline"))
    spline_code, npow = re.subn(r"\(xcc-xx\[j-1\]\)\*\*([23])", r"ipow(xcc-xx[j-1], \1)", spline_code)
    assert npow == 2, npow
    npif = 2
    cases = []
    for izap in (1, 0):
        for ax in "xyz":
            for stretched in ((False, True) if ax == "y" else (False,)):
                for ianal in ((0, 1) if (ax != "z" and izap == 1) else (0,)):
                    cases.append((ax, izap, stretched, ianal, 0.7 if (izap == 1 and not stretched) else 0.0))
    for ax, izap, stretched, ianal, lind in cases:
        ns = fm.base_namespace()
        ns.update(fm.module_parameters(mp))
        yp = yp_stretched if stretched else yp_uniform
        ns.update(dict(nx=NX, ny=NY, nz=NZ, dx=d[0], dy=d[1], dz=d[2], xlx=LEN[0], yly=LEN[1], zlz=LEN[2], npif=npif, izap=izap, ianal=ianal,
                       xsize=fm.FArr(np.array([NX, NY, NZ])), ysize=fm.FArr(np.array([NX, NY, NZ])),
                       zsize=fm.FArr(np.array([NX, NY, NZ])), yp=fm.FArr(yp.copy()), ipow=ipow))
        axis = "xyz".index(ax)
        n_line = (NX, NY, NZ)[axis]
        na, nb = [(NY, NZ), (NX, NZ), (NX, NY)][axis]
        coords = [np.arange(NX) * d[0], yp, np.arange(NZ) * d[2]][axis]
        nobj, xi, xf, nip, nfp = geometry(rng, n_line, na, nb, LEN[axis], coords, npif)
        # bodies on a domain boundary keep the default count npif there (genepsi3d.f90:650-651: verif_epsi only counts
        # fluid points in front of a fluid -> solid transition), which is what gives them their ghost points
        for b in range(nb):
            for a in range(na):
                for i in range(nobj[a, b]):
                    if xi[i, a, b] <= 0.0:
                        nip[i + 1, a, b] = npif
                    if xf[i, a, b] >= LEN[axis]:
                        nfp[i + 1, a, b] = npif
        ana_i = xi + rng.uniform(-0.2, 0.2, xi.shape) * d[axis] * (xi > 0)
        ana_f = xf + rng.uniform(-0.2, 0.2, xf.shape) * d[axis] * (xf < LEN[axis])
        ns[f"nobj{ax}"] = fm.FArr(nobj)
        ns[f"{ax}i"] = fm.FArr(xi); ns[f"{ax}f"] = fm.FArr(xf)
        ns[f"n{ax}ipif"] = fm.FArr(nip, lb=(0, 1, 1)); ns[f"n{ax}fpif"] = fm.FArr(nfp, lb=(0, 1, 1))
        ns["ana_i"] = fm.FArr(ana_i); ns["ana_f"] = fm.FArr(ana_f)
        exec(spline_code, ns)
        spl = ns["cubic_spline"]
        unmatched = [0]

        def spline_call(xa, ya, n, x, y, _spl=spl, _um=unmatched):
            loc = _spl(xa, ya, n, x, y)
            xx, nc = loc["xx"], loc["nc"]
            if not any(xx[q - 1] <= x <= xx[q] for q in range(2, nc + 1)):
                _um[0] += 1
            return loc["y"]
        ns["spline_call"] = spline_call
        _, code = tr.subroutine(sub_text(ibm, f"cubspl{ax}"))
        assert code.count("bcimp = lind\n") == 1
        code = code.replace("bcimp = lind\n", "bcimp = lind\n    ypol = 0.0\n", 1)
        assert code.count("pass  # call cubic_spline") == 1
        code = code.replace("pass  # call cubic_spline", "ypol = spline_call(xa, ya, na, xpol, ypol)")
        if ax != "z":
            idx = {"x": "i, j, k", "y": "j, i, k"}[ax]
            assert code.count(f"pass  # call analitic_{ax}") == 2
            code = code.replace(f"pass  # call analitic_{ax}", f"ana_resi = ana_i[{idx}]", 1)
            code = code.replace(f"pass  # call analitic_{ax}", f"ana_resf = ana_f[{idx}]", 1)
        exec(code, ns)
        u = fm.FArr(u0.copy(order="F"))
        with np.errstate(all="ignore"):
            ns[f"cubspl{ax}"](u, lind)
        tag = f"{ax}/izap{izap}/st{int(stretched)}/an{ianal}"
        out[f"{tag}/out"] = u.a.copy()
        out[f"{tag}/nobj"] = nobj; out[f"{tag}/xi"] = xi; out[f"{tag}/xf"] = xf
        out[f"{tag}/nipif"] = nip; out[f"{tag}/nfpif"] = nfp
        out[f"{tag}/ana_i"] = ana_i; out[f"{tag}/ana_f"] = ana_f
        out[f"{tag}/coords"] = coords
        out[f"{tag}/lind"] = np.float64(lind)
        out[f"{tag}/unmatched"] = np.int64(unmatched[0])
        changed = int((u.a != u0).sum())
        print(tag, "points rebuilt:", changed, "spline calls that matched no interval:", unmatched[0], "finite:", bool(np.isfinite(u.a).all()))
        assert changed > 0
    out["meta/n"] = np.array([NX, NY, NZ]); out["meta/len"] = np.array(LEN); out["meta/npif"] = np.int64(npif)
    out["meta/nobjmax"] = np.int64(NOBJMAX)
    out["meta/cases"] = np.array([f"{ax}/izap{izap}/st{int(st)}/an{ia}" for ax, izap, st, ia, _ in cases])
    np.savez_compressed(os.path.join(HERE, "ibm_cubspl.npz"), **out)
    print("ibm_cubspl.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
