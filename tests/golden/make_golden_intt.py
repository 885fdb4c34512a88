#!/usr/bin/env python
"""Golden vectors for the explicit time integrators, from the reference source.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_intt.py

Executes the reference's own statements (via f90mini.py) of
  * src/variables.f90         the adt/bdt/cdt/gdt/ntime/iadvance_time block of init_variables (:1340-1423)
  * src/time_integrators.f90  intt (:20-190), explicit branch (iimplicit = 0)
for itimescheme = 1 (Euler), 2 (AB2), 3 (AB3), 5 (RK3) over four time steps, feeding a fresh seeded random right-hand
side dvar1(:,:,:,1) at every call, and writes tests/golden/intt.npz: the coefficients and var1 / dvar1 after every call.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90mini as fm  # noqa: E402
from make_golden_poisson import load, sub_text  # noqa: E402

SEED = 20261021
NX, NY, NZ = 5, 4, 3
DT = 0.0125
NSTEPS = 4


def coefficient_block(var_text):
    """the statements between `adt=zero` and the allocation of dux1, wrapped as a subroutine"""
    m = re.search(r"^\s*adt=zero.*?^\s*endif\s*$(?=\s*allocate\(dux1)", var_text, re.S | re.M | re.I)
    assert m, "coefficient block not found"
    return "subroutine time_coefficients()\n" + m.group(0) + "\nend subroutine time_coefficients\n"


def main():
    ti, var, mp = load("time_integrators"), load("variables"), load("module_param")
    out = {}
    rng = np.random.default_rng(SEED)
    tr = fm.Transpiler(arrays_hint={"var1", "dvar1", "forcing1", "xsize", "adt", "bdt", "cdt", "ddt", "gdt"})
    _, coef_code = tr.subroutine(coefficient_block(var))
    _, intt_code = tr.subroutine(sub_text(ti, "intt"))
    # the local `is` of the (skipped) implicit branch is a Python keyword
    intt_code = re.sub(r"^(\s*)is = ", r"\1is_ = ", intt_code, flags=re.M)
    for scheme in (1, 2, 3, 5):
        ns = fm.base_namespace()
        ns.update(fm.module_parameters(mp))
        ns.update(dict(itimescheme=scheme, dt=DT, iimplicit=0, irestart=0, nrank=1, itime=0, itr=1,
                       xsize=fm.FArr(np.array([NX, NY, NZ])), iadvance_time=0, ntime=0, nrhotime=0))
        for nm in ("adt", "bdt", "cdt", "ddt", "gdt"):
            ns[nm] = fm.farr((5,))
        exec(coef_code, ns)
        loc = ns["time_coefficients"]()
        ntime, iadv = int(loc["ntime"]), int(loc["iadvance_time"])
        for nm in ("adt", "bdt", "cdt", "gdt"):
            out[f"s{scheme}/{nm}"] = ns[nm].a.copy()
        out[f"s{scheme}/ntime_iadvance"] = np.array([ntime, iadv])
        exec(intt_code, ns)
        v = fm.FArr(np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ))))
        d = fm.FArr(np.zeros((NX, NY, NZ, max(ntime, 1)), order="F"))
        out[f"s{scheme}/var0"] = v.a.copy()
        call = 0
        for itime in range(1, NSTEPS + 1):
            for itr in range(1, iadv + 1):
                ns["itime"], ns["itr"] = itime, itr
                rhs = np.asfortranarray(rng.uniform(-1, 1, (NX, NY, NZ)))
                d.a[:, :, :, 0] = rhs
                ns["intt"](v, d, None, None, None)
                out[f"s{scheme}/c{call}/rhs"] = rhs
                out[f"s{scheme}/c{call}/var"] = v.a.copy()
                out[f"s{scheme}/c{call}/dvar"] = d.a.copy()
                out[f"s{scheme}/c{call}/itime_itr"] = np.array([itime, itr])
                call += 1
        out[f"s{scheme}/ncalls"] = np.int64(call)
    out["meta/dt"] = np.float64(DT)
    np.savez_compressed(os.path.join(HERE, "intt.npz"), **out)
    print("intt.npz:", len(out), "entries")


if __name__ == "__main__":
    main()
