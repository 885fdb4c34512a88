#!/usr/bin/env python
"""bench.py -- TGV grid-point-steps/s of the B200-native Xcompact3d hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # CPU reference arm (oracle port)

A "step" is one full RK3 time step (3 sub-steps: momentum RHS -> intt -> pre_correc ->
divergence -> spectral Poisson -> gradp -> cor_vel) of the periodic Taylor-Green vortex,
Re=1600, on 512^3 nodes (BASELINE.json configs[1]).  Synthetic data: the analytic TGV field.
`value` times the device-resident solver (fields in HBM); `e2e` drives the same step through
the C ABI with HOST (pinned) velocity arrays, H2D + D2H inside the timed region.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TGV grid-point-steps/sec"
UNIT = "grid-point-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=512, help="nodes per direction (periodic TGV)")
    ap.add_argument("--cpu-n", type=int, default=256, help="box size of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def cpu_reference_run(n, steps, warmup):
    """time the oracle port (C++/OpenMP restatement of the reference step) on the host cores"""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    import numpy as np
    L = ol.lib()
    L.x3do_solver_create.restype = C.c_void_p
    L.x3do_solver_create.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double]
    length = 2 * np.pi
    dt = 0.005 * 64.0 / n  # CFL kept as the mesh is refined
    s = L.x3do_solver_create(n, n, n, (C.c_int * 6)(0, 0, 0, 0, 0, 0), length, length, length, 1600.0, dt, 5, 4, 4, 3, 0, 0.0)
    if not s:
        raise RuntimeError(L.x3do_last_error().decode())
    s = C.c_void_p(s)
    L.x3do_solver_init_tgv(s)
    for _ in range(warmup):
        L.x3do_solver_step(s, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        L.x3do_solver_step(s, 1)
    dtw = time.perf_counter() - t0
    out = (C.c_double * 4)()
    L.x3do_solver_postprocess_tgv(s, out)
    L.x3do_solver_destroy(s)
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return dict(value=n ** 3 * steps / dtw, seconds=dtw, cores=cores, eek=out[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    steps, warmup = args.steps, args.warmup
    r = cpu_reference_run(n, steps, warmup)
    sample = (f"periodic TGV {n}^3 (same Re, RK3, CFL-scaled dt), {steps} full RK3 steps after {warmup} warm-up, "
              f"oracle C++/OpenMP restatement of the reference step (the Fortran/MPI build cannot be built here: "
              f"no Fortran compiler, no MPI, 2DECOMP&FFT un-vendored)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TGV periodic {args.n}^3 Re=1600 RK3 (BASELINE configs[1]); CPU arm runs a bounded {n}^3 sample"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    from incompact3d_b200 import X3D

    n = args.n
    length = 2 * np.pi
    dt = 0.005 * 64.0 / n
    x = X3D(local)
    if world > 1:
        # 2-D pencil decomposition with p_row=1, p_col=N (slabs): on NVSwitch every byte costs the same
        # whichever peer it goes to, and 1xN moves the fewest bytes (x<->y transposes become local)
        from incompact3d_b200 import nccl_unique_id
        obj = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        x.decomp_init(n, n, n, 1, world, rank, world, obj[0])
    x.solver_init(n, n, n, ncl=(0,) * 6, xlx=length, yly=length, zlz=length, re=1600.0, dt=dt, p_row=1, p_col=world)
    x.solver_init_tgv()
    stream = torch.cuda.ExternalStream(x.stream)
    npts = n ** 3

    def barrier():
        x.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        x.solver_step(1)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = x.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        x.solver_step(1)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:  # device time, max over ranks
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = x.launch_count - l0
    clk = clocks.stop()
    value = npts * args.steps / (ms * 1e-3)
    diag = x.solver_diagnostics_tgv()

    # ---- roofline of the dominant kernel, timed live in one extra instrumented step -------------
    roof = x.profile_step() if hasattr(x, "profile_step") else None
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("hbm_gbs", 6650.0)
    roofline = None
    if roof:
        comp = [r for r in roof if r["name"].startswith("compact_")]
        k = max(comp, key=lambda r: r["total_ms"])
        achieved = 16.0 * (npts / world) / (k["avg_ms"] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel class at 512^3 on one GPU, from the
        # `ncu --set full` capture summarised in profiles/r1f_ops_ncu_summary.txt (k_pair: 1.0738 GB + 1.0290 GB;
        # k_contig: 1.0739 + 1.0279).  It equals the algorithmic 2 x 8 B x 512^3 = 2.147 GB to within the write-back
        # still in L2 at kernel end: no re-reads.
        traffic = 2.1029e9 if (n == 512 and world == 1) else None
        roofline = {"bound": "hbm", "kernel": k["name"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": "ncu --set full, profiles/r1f_ops_ncu_summary.txt" if traffic else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                    "algorithmic_bytes_per_launch": 16.0 * npts / world,
                    "note": "dominant compact-operator kernel class; 16 B per output point (SURVEY 8d); per-launch CUDA events "
                            "on the launching stream during one extra instrumented step", "launches_per_step": k["count"],
                    "share_of_step": k["total_ms"] / sum(r["total_ms"] for r in roof),
                    "classes": roof}

    # ---- e2e: the same step driven with HOST velocity arrays through the C ABI -------------------------
    e2e = None
    if not args.no_e2e:
        nzl = x._solver_shape[2]
        hu, hv, hw = (torch.empty((max(nzl, 1), n, n), dtype=torch.float64).pin_memory() for _ in range(3))
        x.solver_get_velocity(hu, hv, hw)
        k2 = max(2, min(args.steps, 5))
        for _ in range(1):
            x.solver_set_velocity(hu, hv, hw); x.solver_step(1); x.solver_get_velocity(hu, hv, hw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k2):
            x.solver_set_velocity(hu, hv, hw)   # H2D of the step's inputs (pinned host memory)
            x.solver_step(1)
            x.solver_get_velocity(hu, hv, hw)   # D2H of the step's result
        barrier()
        te = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        e2e = {"value": npts * k2 / te, "unit": UNIT, "h2d_bytes_per_step": 3 * npts * 8, "d2h_bytes_per_step": 3 * npts * 8,
               "steps": k2, "note": "x3d_solver_set_velocity(host) + x3d_solver_step + x3d_solver_get_velocity(host) per step"}

    cpu = None
    if not args.no_cpu_baseline and rank == 0:
        try:
            steps_cpu = 4
            r = cpu_reference_run(args.cpu_n, steps_cpu, 1)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": f"periodic TGV {args.cpu_n}^3, {steps_cpu} RK3 steps after 1 warm-up ({r['seconds']:.1f} s), oracle "
                             f"C++/OpenMP restatement of the reference step (no Fortran/MPI toolchain on the box)"}
        except Exception as e:  # the baseline is reported, never fatal
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TGV periodic {n}^3 Re=1600 RK3 dt={dt:g} (BASELINE configs[1])",
                   "parallelism": f"{world} GPU" + ("" if world == 1 else f", 2-D pencil decomposition p_row=1 x p_col={world} (slabs), transposes = one kernel storing into peer pencils over NVLink (CUDA IPC), device-side flag barriers"), "l2": "fields are 1 GiB each, far larger than the 126 MB L2; no flush needed",
                   "diagnostics_after_run": diag},
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    x.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
