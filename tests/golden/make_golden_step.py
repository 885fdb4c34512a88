#!/usr/bin/env python
"""Golden vectors for the wall / forcing pieces of the time step, from the reference source.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_step.py

Executes the reference's own statements (via f90mini.py) of
  * src/navier.f90        pre_correc  (Dirichlet, free-slip and periodic faces; wall pressure-gradient terms)
  * src/navier.f90        gradp       (only its closing part runs: the capture of the wall pressure gradients,
                                       :439-496; the operator / transpose calls inside are skipped, px1/py1/pz1 are inputs)
  * src/Case-Channel.f90  channel_cfr (constant flow rate)
on small seeded random fields and writes tests/golden/step.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402
from make_golden_poisson import load, sub_text  # noqa: E402

SEED = 20261019
NX, NY, NZ = 7, 6, 5
WALLS = ["bxx1", "bxy1", "bxz1", "bxxn", "bxyn", "bxzn", "byx1", "byy1", "byz1", "byxn", "byyn", "byzn",
         "bzx1", "bzy1", "bzz1", "bzxn", "bzyn", "bzzn"]
DPD = ["dpdyx1", "dpdzx1", "dpdyxn", "dpdzxn", "dpdxy1", "dpdzy1", "dpdxyn", "dpdzyn", "dpdxz1", "dpdyz1", "dpdxzn", "dpdyzn"]


def plane_shape(name):
    face = name[1] if name.startswith("b") else name[4]   # bxx1 -> x face ; dpdyx1 -> x face
    return {"x": (NY, NZ), "y": (NX, NZ), "z": (NX, NY)}[face]


def main():
    nav, chan, mp = load("navier"), load("Case-Channel"), load("module_param")
    out = {}
    rng = np.random.default_rng(SEED)
    arrays = {"xsize", "xstart", "xend", "zsize", "dims", "gdt", "ux", "uy", "uz", "ep", "px1", "py1", "pz1", "pp3", "ppy",
              "dummy_coords", "dummy_periods", "ux1", "uy1", "uz1"} | set(WALLS) | set(DPD)
    tr = fm.Transpiler(arrays_hint=arrays)
    for tag, ncl in (("y22", (0, 0, 2, 2, 0, 0)), ("x22z22y11", (2, 2, 1, 1, 2, 2)), ("x21y12z11", (2, 1, 1, 2, 1, 1))):
        ns = fm.base_namespace()
        ns.update(fm.module_parameters(mp))
        ns.update(dict(nx=NX, ny=NY, nz=NZ, nym=NY - 1, nclx1=ncl[0], nclxn=ncl[1], ncly1=ncl[2], nclyn=ncl[3], nclz1=ncl[4],
                       nclzn=ncl[5], itype=ns.get("itype_tgv", 2), iforces=0, nrank=1, itime=1, ilist=10, iibm=0, mhd_active=False, itr=2,
                       xsize=fm.FArr(np.array([NX, NY, NZ])), xstart=fm.FArr(np.array([1, 1, 1])), xend=fm.FArr(np.array([NX, NY, NZ])),
                       dims=fm.FArr(np.array([1, 1])), gdt=fm.FArr(np.array([0.11, 0.07, 0.05]))))
        for nm in WALLS:
            ns[nm] = fm.farr(plane_shape(nm))          # no-slip walls: zero wall velocity
        for nm in DPD:
            ns[nm] = fm.FArr(np.asfortranarray(rng.uniform(-1, 1, plane_shape(nm))))
            out[f"pre_correc/{tag}/in/{nm}"] = ns[nm].a.copy()
        u = [np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ))) for _ in range(3)]
        for q, nm in enumerate(("ux", "uy", "uz")):
            out[f"pre_correc/{tag}/in/{nm}"] = u[q].copy()
        _, code = tr.subroutine(sub_text(nav, "pre_correc"))
        # MPI_CART_GET on one rank returns the 1 x 1 process grid (the transpiler skips external calls)
        assert "pass  # call mpi_cart_get" in code.lower()
        code = code.replace("pass  # call MPI_CART_GET", "dims[1] = 1; dims[2] = 1").replace("pass  # call mpi_cart_get", "dims[1] = 1; dims[2] = 1")
        exec(code, ns)
        fu = [fm.FArr(a.copy(order="F")) for a in u]
        ns["pre_correc"](fu[0], fu[1], fu[2], fm.farr((NX, NY, NZ)))
        for q, nm in enumerate(("ux", "uy", "uz")):
            out[f"pre_correc/{tag}/out/{nm}"] = fu[q].a.copy()
        for nm in DPD:
            out[f"pre_correc/{tag}/out/{nm}"] = ns[nm].a.copy()
        out[f"pre_correc/{tag}/ncl"] = np.array(ncl)
        # ---- gradp: wall-gradient capture
        p3 = [np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ))) for _ in range(3)]
        _, code = tr.subroutine(sub_text(nav, "gradp"))
        exec(code, ns)
        fp = [fm.FArr(a.copy(order="F")) for a in p3]
        ns["gradp"](fp[0], fp[1], fp[2], fm.farr((NX - 1, NY - 1, NZ - 1)))
        for q, nm in enumerate(("px1", "py1", "pz1")):
            out[f"gradp/{tag}/in/{nm}"] = p3[q].copy()
            assert np.array_equal(fp[q].a, p3[q])   # the operator calls are skipped: inputs unchanged
        for nm in DPD:
            out[f"gradp/{tag}/out/{nm}"] = ns[nm].a.copy()
    out["meta/gdt"] = np.array([0.11, 0.07, 0.05])
    out["meta/itr"] = np.int64(2)
    # ---- channel_cfr
    for tag, stretched in (("uniform", False), ("stretched", True)):
        ns = fm.base_namespace()
        ns.update(fm.module_parameters(mp))
        ppy = rng.uniform(0.5, 1.5, NY) if stretched else np.ones(NY)
        dy, yly = 0.37, 2.0
        ns.update(dict(dy=dy, yly=yly, ppy=fm.FArr(ppy.copy()), nrank=1, itime=1, ilist=10, ifirst=1, ilast=10,
                       xsize=fm.FArr(np.array([NX, NY, NZ])), xstart=fm.FArr(np.array([1, 1, 1])), zsize=fm.FArr(np.array([NX, NY, NZ]))))
        _, code = tr.subroutine(sub_text(chan, "channel_cfr"))
        exec(code, ns)
        u = np.asfortranarray(rng.uniform(0, 1, (NX, NY, NZ)))
        fu = fm.FArr(u.copy(order="F"))
        ns["channel_cfr"](fu, 2.0 / 3.0)
        out[f"cfr/{tag}/in"] = u
        out[f"cfr/{tag}/out"] = fu.a.copy()
        out[f"cfr/{tag}/ppy"] = ppy
        out[f"cfr/{tag}/dy_yly"] = np.array([dy, yly])
    np.savez_compressed(os.path.join(HERE, "step.npz"), **out)
    print("step.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
