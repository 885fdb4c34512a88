#!/usr/bin/env python
"""Per-kernel roofline micro-benchmark of the compact operators (GPU box only).

For each operator: warm-up, then N timed launches bracketed by CUDA events on the
context stream; achieved GB/s = 16 B x points / time.  Inputs (1 GiB
per 512^3 field) are far larger than the 126 MB L2, and input/output are distinct
buffers, so nothing is served from cache between iterations.
"""
import argparse
import json
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(name):
    """-> (family, axis letter, bc) ; family in d1, d2, fil, dvp, ivp, dpv, ipv"""
    m = re.match(r"^der([xyz])(\1?)_(\d\d)$", name)
    if m:
        return ("d2" if m.group(2) else "d1"), m.group(1), m.group(3)
    m = re.match(r"^fil([xyz])_(\d\d)$", name)
    if m:
        return "fil", m.group(1), m.group(2)
    m = re.match(r"^(der|inter)([xyz])(vp|pv)$", name)
    return ("d" if m.group(1) == "der" else "i") + m.group(3), m.group(2), "00"


def call_op(x3d, A, name, fam, ax, u, t, dims, npaire):
    """the reference argument lists (src/module_param.f90:136-167, src/derive.f90:3796-5615)"""
    nx, ny, nz = dims
    axis = "xyz".index(ax)
    n, nm = A.n, A.nm
    p = "p" if npaire == 1 else ""
    fn = getattr(x3d, name)
    if fam == "fil":
        c, (f, s, w) = A.filter(0.45, p=(npaire == 1))
        x3d.set_filter_coeffs(axis, c)
        fn(t, u, None, None, f, s, w, nx, ny, nz, npaire, 0.0)
        return
    if fam in ("d1", "d2"):
        f, s, w = A.lu(fam + p)
        if fam == "d1" and ax == "y":
            fn(t, u, None, None, f, s, w, np.ones(ny), nx, ny, nz, npaire, 0.0)
        else:
            fn(t, u, None, None, f, s, w, nx, ny, nz, npaire, 0.0)
        return
    if fam == "dvp":
        lu = list(A.lu("vp"))
    elif fam == "ivp":
        lu = list(A.lu("ivpp"))
    elif fam == "dpv":
        lu = list(A.lu("pvp")) + list(A.lu("vp"))
    else:
        lu = list(A.lu("ipvp")) + list(A.lu("ivp"))
    vel = [nx, ny, nz]
    to_p = fam in ("dvp", "ivp")
    if ax == "x":
        ints = [vel[0], nm, vel[1], vel[2]] if to_p else [nm, vel[0], vel[1], vel[2]]
    elif ax == "y":
        ints = [vel[0], vel[1], nm, vel[2]] if to_p else [vel[0], nm, vel[1], vel[2]]
    else:
        ints = [vel[0], vel[1], vel[2], nm] if to_p else [vel[0], vel[1], nm, vel[2]]
    extra = [np.ones(nm if name == "deryvp" else n)] if name in ("deryvp", "derypv") else []
    fn(t, u, None, None, *lu, *extra, *ints, npaire)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ops", default="derx_00,dery_00,derz_00,derxx_00,deryy_00,derzz_00,"
                    "interxvp,deryvp,derzpv,derx_11,dery_11,derz_11")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from incompact3d_b200 import X3D, AxisSchemes   # AxisSchemes: the library's host-side schemes() (plays the Fortran host)

    nx, ny, nz = args.n
    x3d = X3D(0)
    stream = torch.cuda.ExternalStream(x3d.stream)
    peak = None
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs")
    res = []
    with torch.cuda.stream(stream):
        xs = torch.arange(nx, dtype=torch.float64, device="cuda") * (2 * np.pi / nx)
        u = (torch.sin(xs)[None, None, :] * torch.cos(torch.arange(ny, device="cuda", dtype=torch.float64) * (2 * np.pi / ny))[None, :, None]
             * torch.cos(torch.arange(nz, device="cuda", dtype=torch.float64) * (2 * np.pi / nz))[:, None, None]).contiguous()
        u += 0.1 * (torch.rand_like(u) * 2 - 1)
        t = torch.empty_like(u)
        for name in args.ops.split(","):
            fam, ax, bc = parse(name)
            axis = "xyz".index(ax)
            n = (nx, ny, nz)[axis]
            A = AxisSchemes(n, int(bc[0]), int(bc[1]), 2 * np.pi)
            x3d.set_deriv_coeffs(axis, A.c)
            ncl = [True, True, True]
            ncl[axis] = A.periodic
            x3d.set_flags(iibm=0, istret=0, iimplicit=0, nclx=ncl[0], ncly=ncl[1], nclz=ncl[2])
            if fam in ("dpv", "ipv") and not A.periodic:
                continue
            npaire = 1 if fam != "dvp" else 0
            for _ in range(args.warmup):
                call_op(x3d, A, name, fam, ax, u, t, (nx, ny, nz), npaire)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            for _ in range(args.iters):
                call_op(x3d, A, name, fam, ax, u, t, (nx, ny, nz), npaire)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            gbs = 16.0 * nx * ny * nz / (ms * 1e-3) / 1e9
            r = dict(op=name, n=[nx, ny, nz], ms=ms, gbs=gbs, frac=(gbs / peak if peak else None))
            res.append(r)
            print(json.dumps(r), flush=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
