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
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ops", default="derx_00,dery_00,derz_00,derxx_00,deryy_00,derzz_00,"
                    "interxvp,deryvp,derzpv,derx_11,dery_11,derz_11")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from incompact3d_b200 import X3D
    import helpers as H
    import oracle_lib as ol  # coefficients only (plays the Fortran host's schemes()); not timed

    nx, ny, nz = args.n
    x3d = X3D(0)
    stream = torch.cuda.ExternalStream(x3d.stream)
    peak = None
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs")
    rng = np.random.default_rng(20261017)
    res = []
    with torch.cuda.stream(stream):
        xs = torch.arange(nx, dtype=torch.float64, device="cuda") * (2 * np.pi / nx)
        u = (torch.sin(xs)[None, None, :] * torch.cos(torch.arange(ny, device="cuda", dtype=torch.float64) * (2 * np.pi / ny))[None, :, None]
             * torch.cos(torch.arange(nz, device="cuda", dtype=torch.float64) * (2 * np.pi / nz))[:, None, None]).contiguous()
        u += 0.1 * (torch.rand_like(u) * 2 - 1)
        t = torch.empty_like(u)
        for name in args.ops.split(","):
            fam, ax, bc = H.parse(name)
            axis = "xyz".index(ax)
            n = (nx, ny, nz)[axis]
            if bc is None:
                bc = "00"
            A = ol.Axis(n, int(bc[0]), int(bc[1]), 2 * np.pi, af=0.45)
            H.configure(x3d, A, axis)
            uin = u
            if fam in ("dpv", "ipv") and not A.periodic:
                continue
            npaire = 1 if fam != "dvp" else 0
            for _ in range(args.warmup):
                H.product_op(x3d, name, uin, A, npaire, t=t)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            for _ in range(args.iters):
                H.product_op(x3d, name, uin, A, npaire, t=t)
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
