"""Host-side helpers for the cases the device step runs (set-up only, nothing here is on the per-time-step path).

`cylinder_geometry` produces what the reference's `genepsi3d` (src/genepsi3d.f90, geometry of
src/Case-Cylinder-wake.f90:46-82) hands to the immersed-boundary pre-pass for a circular cylinder along z: the mask
`ep1` and, per direction, `nobj`, `xi`, `xf`, `nxipif`, `nxfpif` (src/module_param.f90:546-556), with the wall positions
taken from the analytic circle instead of the reference's refined-mesh search."""
from __future__ import annotations

import numpy as np

NOBJMAX, NPIF, IZAP = 1, 2, 1


def cylinder_geometry(nn, lens, cex, cey, ra, nclx=True):
    """-> (ep1, [per axis: nobj, xi, xf, nipif, nfpif], (dx, dy, dz)); nclx: x has nx-1 intervals (inflow / outflow)"""
    nx, ny, nz = nn
    dx = lens[0] / (nx - 1 if nclx else nx)
    dy, dz = lens[1] / ny, lens[2] / nz
    xs, ys = np.arange(nx) * dx, np.arange(ny) * dy
    ep = np.zeros(nn, order="F")
    geo = []
    for na, nb in ((ny, nz), (nx, nz), (nx, ny)):
        nobj = np.zeros((na, nb), dtype=np.int32, order="F")
        xi = np.zeros((NOBJMAX, na, nb), order="F")
        xf = np.zeros((NOBJMAX, na, nb), order="F")
        nip = np.full((NOBJMAX + 1, na, nb), NPIF, dtype=np.int32, order="F")
        geo.append([nobj, xi, xf, nip, nip.copy(order="F")])
    inside = (xs[:, None] - cex) ** 2 + (ys[None, :] - cey) ** 2 <= ra ** 2
    ep[inside, :] = 1.0
    jj = np.nonzero(np.abs(ys - cey) < ra)[0]
    half = np.sqrt(ra ** 2 - (ys[jj] - cey) ** 2)
    geo[0][0][jj, :] = 1
    geo[0][1][0, jj, :] = (cex - half)[:, None]
    geo[0][2][0, jj, :] = (cex + half)[:, None]
    ii = np.nonzero(np.abs(xs - cex) < ra)[0]
    half = np.sqrt(ra ** 2 - (xs[ii] - cex) ** 2)
    geo[1][0][ii, :] = 1
    geo[1][1][0, ii, :] = (cey - half)[:, None]
    geo[1][2][0, ii, :] = (cey + half)[:, None]
    geo[2][0][inside] = 1
    geo[2][2][0][inside] = lens[2]
    return ep, geo, (dx, dy, dz)


def apply_cylinder(x, nn, lens, cex, cey, ra, ubc=(0.0, 0.0, 0.0), u1=1.0, u2=1.0, inflow_noise=0.0, iibm=2):
    """configure an initialised X3D solver (itype = 5) for the flow past a circular cylinder"""
    ep, geo, d = cylinder_geometry(nn, lens, cex, cey, ra)
    x.solver_set_case(u1=u1, u2=u2, inflow_noise=inflow_noise, iibm=iibm, ubc=ubc)
    x.solver_set_ibm_mask(ep)
    for axis, (nobj, xi, xf, nip, nfp) in enumerate(geo):
        x.set_ibm_geometry(axis, NOBJMAX, NPIF, IZAP, nobj, xi, xf, nip, nfp, d[axis], lens[axis],
                           coords=np.arange(nn[1]) * d[1] if axis == 1 else None)
    return ep, geo
