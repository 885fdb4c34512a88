#!/usr/bin/env python
"""Golden vectors for the Poisson-solver set-up and the stretched-mesh pieces, from the reference source.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden_poisson.py

Executes the reference's own statements (via f90mini.py) of
  * src/stretching.f90   stretching_full                       (yp, ypi, ppy, pp2y, pp4y, ppyi, pp2yi, pp4yi, alpha)
  * src/poisson.f90      abxyz, waves, matrice_refinement      (twiddles, modified wavenumbers, kxyz, a / a2 / a3)
  * src/tools.f90        inversion5_v1, inversion5_v2          (pentadiagonal solves of seeded right-hand sides)
on small meshes for the boundary-condition / istret combinations the solver supports and writes
tests/golden/poisson.npz.  The GPU box has no /root/reference: tests only read the committed file.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402
from make_golden import Ref, REF  # noqa: E402

SEED = 20261018


def load(name):
    return open(os.path.join(REF, "src", name + ".f90")).read()


def sub_text(text, name):
    """text of `subroutine name` ... `end subroutine name` (module procedures included)"""
    ms = list(re.finditer(r"^\s*(?:module\s+)?subroutine\s+%s\b.*?^\s*end\s+subroutine\s+%s\b" % (name, name), text,
                          re.S | re.M | re.I))
    if not ms:
        raise KeyError(name)
    t = max((m.group(0) for m in ms), key=len)   # the implementation, not an interface block
    return re.sub(r"^\s*module\s+subroutine", "subroutine", t, flags=re.I)


def one_config(src, nx, ny, nz, bc, istret, beta, lengths, out, tag):
    """nx,ny,nz: velocity nodes; bc = (bcx,bcy,bcz) 0 periodic / 1 not"""
    ncl = tuple((0, 0) if b == 0 else (2, 2) for b in bc)
    if bc == (1, 1, 1):
        ncl = ((1, 1), (1, 1), (1, 1))
    ref = Ref(src, nx, ny, nz, ncl, lengths=lengths, istret=istret)
    ns = ref.ns
    nxm, nym, nzm = ref.nm
    xlx, yly, zlz = lengths
    ns.update(dict(nx=nx, ny=ny, nz=nz, nxm=nxm, nym=nym, nzm=nzm, xlx=xlx, yly=yly, zlz=zlz,
                   dx=ref.d[0], dy=ref.d[1], dz=ref.d[2], bcx=bc[0], bcy=bc[1], bcz=bc[2],
                   ncly1=ncl[1][0], nclyn=ncl[1][1], beta=beta, istret=istret, nrank=0,
                   cx_one_one=complex(1.0, 1.0), twopi=2.0 * np.arccos(-1.0), pi=np.arccos(-1.0),
                   epsilon=1.e-16))   # tools.f90:1240-1244, DOUBLE_PREC build
    tr = fm.Transpiler()
    # ---- stretching ------------------------------------------------------------------------------------
    if istret:
        _, code = tr.subroutine(sub_text(src["stretching"], "stretching_full"))
        exec(code, ns)
        arrs = {k: fm.farr((ny,)) for k in ("yp", "ypi", "ppy", "pp2y", "pp4y", "ppyi", "pp2yi", "pp4yi")}
        loc = ns["stretching_full"](ny, *[arrs[k] for k in ("yp", "ypi", "ppy", "pp2y", "pp4y", "ppyi", "pp2yi", "pp4yi")], False)
        ns["alpha"] = loc["alpha"]
        for k, v in arrs.items():
            out[f"{tag}/{k}"] = v.a.copy()
            ns[k] = v
        out[f"{tag}/alpha"] = np.float64(loc["alpha"])
    # ---- abxyz / waves ---------------------------------------------------------------------------------
    px, py, pz = nxm, nym, nzm          # pressure mesh
    nzh = pz // 2 + 1
    _, code = tr.subroutine(sub_text(src["poisson"], "abxyz"))
    exec(code, ns)
    ab = {k: fm.farr((n,)) for k, n in (("ax", px), ("bx", px), ("ay", py), ("by", py), ("az", pz), ("bz", pz))}
    ns["abxyz"](ab["ax"], ab["ay"], ab["az"], ab["bx"], ab["by"], ab["bz"], px, py, pz, *bc)
    for k, v in ab.items():
        out[f"{tag}/{k}"] = v.a.copy()
        ns[k] = v
    # module arrays of decomp_2d_poisson (poisson.f90:34-47), spectral decomposition sp on one rank
    for k, n in (("xkx", nx), ("xk2", nx), ("exs", nx), ("yky", ny), ("yk2", ny), ("eys", ny),
                 ("zkz", nz // 2 + 1), ("zk2", nz // 2 + 1), ("ezs", nz // 2 + 1)):
        ns[k] = fm.farr((n,), complex)
    for nm_, st, en in (("sp", (1, 1, 1), (px, py, nzh)),):
        for pen in "xyz":
            ns[f"{nm_}__{pen}st"] = fm.FArr(np.array(st))
            ns[f"{nm_}__{pen}en"] = fm.FArr(np.array(en))
            ns[f"{nm_}__{pen}sz"] = fm.FArr(np.array(en))
    ns["kxyz"] = fm.farr((px, py, nzh), complex)
    arrays = {"xkx", "xk2", "exs", "yky", "yk2", "eys", "zkz", "zk2", "ezs", "kxyz", "sp__xst", "sp__xen", "sp__yst", "sp__yen",
              "sp__zst", "sp__zen", "ax", "bx", "ay", "by", "az", "bz", "yp", "ypi", "ppy", "pp2y", "pp4y", "ppyi", "pp2yi", "pp4yi",
              "a", "a2", "a3", "cw2", "cw22"}
    arrays |= {f"sp%{pen}{w}" for pen in "xyz" for w in ("st", "en", "sz")} | {"spI%yst", "spI%yen"}
    tr2 = fm.Transpiler(arrays_hint=arrays)
    _, code = tr2.subroutine(sub_text(src["poisson"], "waves"))
    exec(code, ns)
    ns["waves"]()
    for k in ("xkx", "xk2", "exs", "yky", "yk2", "eys", "zkz", "zk2", "ezs", "kxyz"):
        out[f"{tag}/{k}"] = ns[k].a.copy()
    # ---- matrice_refinement + inversion5 -----------------------------------------------------------------
    if istret and bc[1] == 1:
        nyh = ny // 2
        ns["a"] = fm.farr((px, nyh, nzh, 5), complex)
        ns["a2"] = fm.farr((px, nyh, nzh, 5), complex)
        ns["a3"] = fm.farr((px, nym, nzh, 5), complex)
        ns["cw2"] = fm.farr((px, py, nzh), complex)
        ns["cw22"] = fm.farr((px, py, nzh), complex)
        _, code = tr2.subroutine(sub_text(src["poisson"], "matrice_refinement"))
        exec(code, ns)
        ns["matrice_refinement"]()
        rng = np.random.default_rng(SEED + istret)
        tools = src["tools"]
        spI = None
        ns["spi__yst"] = ns["sp__yst"]; ns["spi__yen"] = ns["sp__yen"]   # the transpiler lower-cases identifiers
        tr3 = fm.Transpiler(arrays_hint=arrays | {"spi%yst", "spi%yen", "aaa", "aaa_in", "eee", "sr", "a1", "b1", "ja", "jb"})
        if istret != 3:
            out[f"{tag}/a"] = ns["a"].a.copy()
            out[f"{tag}/a2"] = ns["a2"].a.copy()
            _, code = tr3.subroutine(sub_text(tools, "inversion5_v1"))
            exec(code, ns)
            for nm_ in ("a", "a2"):
                e = rng.uniform(-1, 1, (px, nyh, nzh)) + 1j * rng.uniform(-1, 1, (px, nyh, nzh))
                out[f"{tag}/rhs_{nm_}"] = e.copy()
                ee = fm.FArr(np.asfortranarray(e.copy()))
                ns["inversion5_v1"](ns[nm_], ee, spI)
                out[f"{tag}/sol_{nm_}"] = ee.a.copy()
        else:
            out[f"{tag}/a3"] = ns["a3"].a.copy()
            _, code = tr3.subroutine(sub_text(tools, "inversion5_v2"))
            exec(code, ns)
            e = rng.uniform(-1, 1, (px, nym, nzh)) + 1j * rng.uniform(-1, 1, (px, nym, nzh))
            out[f"{tag}/rhs_a3"] = e.copy()
            ee = fm.FArr(np.asfortranarray(e.copy()))
            a3c = fm.FArr(ns["a3"].a.copy())
            ns["inversion5_v2"](a3c, ee, spI)
            out[f"{tag}/sol_a3"] = ee.a.copy()


def main():
    src = {f: load(f) for f in ("schemes", "derive", "filters", "module_param", "poisson", "tools", "stretching")}
    out = {}
    lengths = (2 * np.pi, 2.0, 1.7)
    cfgs = []
    for bc in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1)):
        cfgs.append((bc, 0))
    for bc in ((0, 1, 0), (1, 1, 0), (1, 1, 1)):
        for istret in (1, 2, 3):
            cfgs.append((bc, istret))
    names = []
    for bc, istret in cfgs:
        n = [8 if b == 0 else 9 for b in bc]
        n[1] = 12 if bc[1] == 0 else 13
        tag = f"bc{bc[0]}{bc[1]}{bc[2]}_st{istret}"
        one_config(src, n[0], n[1], n[2], bc, istret, 0.259065151, lengths, out, tag)
        out[f"{tag}/n"] = np.array(n)
        names.append(tag)
        print(tag, "ok")
    out["meta/lengths"] = np.array(lengths)
    out["meta/beta"] = np.float64(0.259065151)
    out["meta/tags"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "poisson.npz"), **out)
    print("poisson.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
