#!/usr/bin/env python
"""Golden vectors for the inflow / outflow and solid-body pieces of the time step, from the reference source.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_cyl.py

Executes the reference's own statements (via f90mini.py) of
  * src/Case-Cylinder-wake.f90  inflow, outflow (convective outflow with the four choices of its celerity)
  * src/ibm.f90                 body, corgp_IBM
  * src/navier.f90              pre_correc with non-zero wall velocities (inflow / outflow planes of the cylinder case)
                                and with the inflow / outflow flow-rate correction of itype = channel, nclx = 2
  * src/Case-Channel.f90        momentum_forcing_channel (constant pressure gradient, spin-up rotation)
on small seeded random fields and writes tests/golden/cyl.npz.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402
from make_golden_poisson import load, sub_text  # noqa: E402
from make_golden_step import DPD, WALLS, plane_shape  # noqa: E402

SEED = 20261023
NX, NY, NZ = 7, 6, 5


def main():
    cyl, ibm, nav, mp = load("Case-Cylinder-wake"), load("ibm"), load("navier"), load("module_param")
    out = {}
    rng = np.random.default_rng(SEED)
    arrays = {"xsize", "xstart", "xend", "dims", "gdt", "ux", "uy", "uz", "phi", "ep", "ep1", "ux1", "uy1", "uz1", "px", "py", "pz", "bxo", "byo",
              "bzo", "cp", "dummy_coords", "dummy_periods"} | set(WALLS) | set(DPD)
    tr0 = fm.Transpiler(arrays_hint=arrays)

    class Tr:   # the scalar loop index `is` of the (inactive) scalar branches is a Python keyword
        @staticmethod
        def subroutine(text):
            name, code = tr0.subroutine(text)
            return name, re.sub(r"\bis\b", "is_", code)
    tr = Tr
    gdt = np.array([0.11, 0.07, 0.05])
    dx = 0.31

    def base(**kw):
        ns = fm.base_namespace()
        ns.update(fm.module_parameters(mp))
        ns.update(dict(nx=NX, ny=NY, nz=NZ, nym=NY - 1, dx=dx, dy=0.29, dz=0.37, iscalar=0, numscalar=0, nrank=1, itime=3, ilist=10, ifirst=1,
                       ilast=100, itr=2, iibm=2, mhd_active=False, iforces=0,
                       xsize=fm.FArr(np.array([NX, NY, NZ])), xstart=fm.FArr(np.array([1, 1, 1])), xend=fm.FArr(np.array([NX, NY, NZ])),
                       dims=fm.FArr(np.array([1, 1])), gdt=fm.FArr(gdt.copy())))
        ns.update(kw)
        return ns

    # ---- inflow
    ns = base(u1=1.0, u2=1.3, inflow_noise=0.1)
    for nm in ("bxo", "byo", "bzo"):
        ns[nm] = fm.FArr(np.asfortranarray(rng.uniform(0, 1, (NY, NZ))))
        out[f"inflow/in/{nm}"] = ns[nm].a.copy()
    for nm in ("bxx1", "bxy1", "bxz1"):
        ns[nm] = fm.farr((NY, NZ))
    _, code = tr.subroutine(sub_text(cyl, "inflow"))
    exec(code, ns)
    ns["inflow"](fm.farr((NX, NY, NZ, 1)))
    for nm in ("bxx1", "bxy1", "bxz1"):
        out[f"inflow/out/{nm}"] = ns[nm].a.copy()
    out["inflow/u1_noise"] = np.array([1.0, 0.1])

    # ---- outflow, the four branches of the celerity
    u = [np.asfortranarray(rng.uniform(0.2, 1.4, (NX, NY, NZ))) for _ in range(3)]
    for q, nm in enumerate(("ux", "uy", "uz")):
        out[f"outflow/in/{nm}"] = u[q].copy()
    for tag, u1v in (("u1_0", 0.0), ("u1_1", 1.0), ("u1_2", 2.0), ("u1_07", 0.7)):
        ns = base(u1=u1v, u2=1.3)
        for nm in ("bxxn", "bxyn", "bxzn"):
            ns[nm] = fm.farr((NY, NZ))
        _, code = tr.subroutine(sub_text(cyl, "outflow"))
        exec(code, ns)
        fu = [fm.FArr(a.copy(order="F")) for a in u]
        ns["outflow"](fu[0], fu[1], fu[2], fm.farr((NX, NY, NZ, 1)))
        for nm in ("bxxn", "bxyn", "bxzn"):
            out[f"outflow/{tag}/{nm}"] = ns[nm].a.copy()
        out[f"outflow/{tag}/u1_u2"] = np.array([u1v, 1.3])
    out["meta/gdt"] = gdt
    out["meta/itr"] = np.int64(2)
    out["meta/dx"] = np.float64(dx)

    # ---- body, corgp_IBM
    ep = np.asfortranarray((rng.uniform(0, 1, (NX, NY, NZ)) > 0.6).astype(float))
    p3 = [np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ))) for _ in range(3)]
    out["body/in/ep"] = ep
    for q, nm in enumerate(("px", "py", "pz")):
        out[f"corgp/in/{nm}"] = p3[q].copy()
    ns = base()
    _, code = tr.subroutine(sub_text(ibm, "body"))
    exec(code, ns)
    fu = [fm.FArr(a.copy(order="F")) for a in u]
    ns["body"](fu[0], fu[1], fu[2], fm.FArr(ep.copy(order="F")))
    for q, nm in enumerate(("ux", "uy", "uz")):
        out[f"body/out/{nm}"] = fu[q].a.copy()
    _, code = tr.subroutine(sub_text(ibm, "corgp_ibm"))
    exec(code, ns)
    for nlock in (1, 2):
        fu = [fm.FArr(a.copy(order="F")) for a in u]
        ns["corgp_ibm"](fu[0], fu[1], fu[2], *[fm.FArr(a.copy(order="F")) for a in p3], nlock)
        for q, nm in enumerate(("ux", "uy", "uz")):
            out[f"corgp/nlock{nlock}/{nm}"] = fu[q].a.copy()

    # ---- pre_correc with moving walls: cylinder (x: 2,2; no flow-rate correction) and channel-type (with it)
    for tag, itype_name in (("cyl", "itype_cyl"), ("channel", "itype_channel")):
        ns = base()
        ns.update(dict(nclx1=2, nclxn=2, ncly1=0, nclyn=0, nclz1=0, nclzn=0, itype=ns[itype_name], iibm=0))
        for nm in WALLS:
            ns[nm] = fm.FArr(np.asfortranarray(rng.uniform(0.5, 1.5, plane_shape(nm))))
            out[f"pre_correc/{tag}/in/{nm}"] = ns[nm].a.copy()
        for nm in DPD:
            ns[nm] = fm.FArr(np.asfortranarray(rng.uniform(-1, 1, plane_shape(nm))))
            out[f"pre_correc/{tag}/in/{nm}"] = ns[nm].a.copy()
        _, code = tr.subroutine(sub_text(nav, "pre_correc"))
        assert "pass  # call mpi_cart_get" in code.lower()
        code = code.replace("pass  # call MPI_CART_GET", "dims[1] = 1; dims[2] = 1").replace("pass  # call mpi_cart_get", "dims[1] = 1; dims[2] = 1")
        exec(code, ns)
        fu = [fm.FArr(a.copy(order="F")) for a in u]
        ns["pre_correc"](fu[0], fu[1], fu[2], fm.farr((NX, NY, NZ)))
        for q, nm in enumerate(("ux", "uy", "uz")):
            out[f"pre_correc/{tag}/out/{nm}"] = fu[q].a.copy()
        for nm in ("bxxn",):
            out[f"pre_correc/{tag}/out/{nm}"] = ns[nm].a.copy()
    # ---- momentum_forcing_channel: constant pressure gradient and spin-up rotation (Case-Channel.f90:396-420)
    chan = load("Case-Channel")
    d = [np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ, 2))) for _ in range(3)]
    for q, nm in enumerate(("dux", "duy", "duz")):
        out[f"forcing/in/{nm}"] = d[q][..., 0].copy()
    for tag, cpg, itime, spin in (("cpg", True, 5, 0), ("rot", False, 5, 10), ("rot_over", False, 12, 10), ("both", True, 3, 10)):
        ns = base()
        ns.update(dict(cpg=cpg, fcpg=0.0123, idir_stream=1, itime=itime, spinup_time=spin, iin=1, wrotation=0.37, ntime=2))
        _, code = tr.subroutine(sub_text(chan, "momentum_forcing_channel"))
        exec(code, ns)
        fd = [fm.FArr(a.copy(order="F")) for a in d]
        ns["momentum_forcing_channel"](fd[0], fd[1], fd[2], *[fm.FArr(a.copy(order="F")) for a in u])
        for q, nm in enumerate(("dux", "duy", "duz")):
            out[f"forcing/{tag}/{nm}"] = fd[q].a[..., 0].copy()
            assert np.array_equal(fd[q].a[..., 1], d[q][..., 1])
        out[f"forcing/{tag}/par"] = np.array([float(cpg), 0.0123, float(itime), float(spin), 0.37])
    np.savez_compressed(os.path.join(HERE, "cyl.npz"), **out)
    print("cyl.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
