#!/usr/bin/env python
"""bench.py -- TGV grid-point-steps/s of the B200-native Xcompact3d hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # CPU reference arm (oracle port), SAME configuration

A "step" is one full RK3 time step (3 sub-steps: momentum RHS -> intt -> pre_correc ->
divergence -> spectral Poisson -> gradp -> cor_vel) of the periodic Taylor-Green vortex,
Re=1600, on 512^3 nodes (BASELINE.json configs[1]).  Synthetic data: the analytic TGV field.

  value     device-resident solver (fields in HBM), CUDA events on the library's stream, max over ranks
  e2e       the same step as host jobs through the C ABI (x3d_solver_advance_host): every step's 3 velocity arrays
            are copied H2D from pinned host memory and the result D2H inside the timed region; consecutive jobs are
            independent (three host-resident ensemble members advanced in turn), so the copies of neighbouring jobs
            overlap the kernels.  `e2e.serial` is the strictly serial H2D -> step -> D2H figure.
  parity    after the timed region the GPU solver (same rank layout) and the CPU oracle are stepped from the same
            TGV state at --parity-n^3 and compared (max|du|/max|u|, diagnostics); the run FAILS above 1e-11.
  roofline  the kernel class with the largest share of the step (per-launch CUDA events, one instrumented step)
  nvlink    N > 1: bytes each GPU sends to its peers per step through the transposes / their device time
Prints ONE JSON line (rank 0).  Other cases: --case channel (BASELINE configs[2], 256x129x128 stretched).
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
PARITY_TOL = 1e-11


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--case", default="tgv", choices=["tgv", "channel", "cylinder"])
    # --size: the same under torch.distributed.run, whose own parser takes "--n" for an abbreviation of --nnodes / --nproc-per-node
    ap.add_argument("--n", "--size", dest="n", type=int, default=512, help="nodes per direction (periodic TGV)")
    ap.add_argument("--parity-n", type=int, default=256, help="box size of the in-bench GPU-vs-oracle parity run (also the "
                    "bounded CPU sample of cpu_baseline)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="wall-clock budget of the reference arm's timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle legs (parity + cpu_baseline)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def host_threads():
    """cores this process may use -- NOT $OMP_NUM_THREADS (torch.distributed.run exports OMP_NUM_THREADS=1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


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
# workloads
def workload(args):
    import numpy as np
    if args.case == "tgv":
        n = args.n
        length = 2 * np.pi
        return dict(name=f"TGV periodic {n}^3 Re=1600 RK3 dt={0.005 * 64.0 / n:g} (BASELINE configs[1])", dims=(n, n, n), ncl=(0,) * 6,
                    lens=(length,) * 3, re=1600.0, dt=0.005 * 64.0 / n, itimescheme=5, isecondder=4, istret=0, beta=0.0, itype=0)
    if args.case == "cylinder":
        # BASELINE configs[3]: cylinder wake 768x256x32 (769 nodes in x), iibm = 2, AB3 (examples/Cylinder-wake/input_DNS300_LR.i3d, scaled)
        return dict(name="Cylinder wake 769x256x32 Re=300 AB3 dt=0.0025 iibm=2 inflow/outflow (BASELINE configs[3])", dims=(769, 256, 32),
                    ncl=(2, 2, 0, 0, 0, 0), lens=(20.0, 12.0, 6.0), re=300.0, dt=0.0025, itimescheme=3, isecondder=4, istret=0, beta=0.0,
                    itype=5, cyl=(5.0, 6.0, 0.5))
    # BASELINE configs[2]: channel Re_tau=180, 256x129x128, stretched y (examples/Channel/input_DNS_Re180_LR_explicittime.i3d x2)
    return dict(name="Channel 256x129x128 istret=2 beta=0.259065151 Re=4200 RK3 dt=0.005 isecondder=5, constant flow rate (BASELINE configs[2])",
                dims=(256, 129, 128), ncl=(0, 0, 2, 2, 0, 0), lens=(8.0, 2.0, 4.0), re=4200.0, dt=0.005, itimescheme=5, isecondder=5,
                istret=2, beta=0.259065151, itype=3)


class Oracle:
    """the CPU restatement of the reference step (oracle/, C++/OpenMP) -- the checker and the CPU baseline"""

    def __init__(self, w, dims=None):
        import ctypes as C
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        self.C = C
        self.L = L = ol.lib()
        L.x3do_set_threads.argtypes = [C.c_int]
        self.cores = L.x3do_set_threads(host_threads())
        nx, ny, nz = dims or w["dims"]
        self.dims = (nx, ny, nz)
        L.x3do_solver_create_case.restype = C.c_void_p
        L.x3do_solver_create_case.argtypes = ([C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double]
                                              + [C.c_int, C.c_double, C.c_double])
        s = L.x3do_solver_create_case(nx, ny, nz, (C.c_int * 6)(*w["ncl"]), *[float(v) for v in w["lens"]], float(w["re"]), float(w["dt"]),
                                      w["itimescheme"], 4, w["isecondder"], 3, w["istret"], float(w["beta"]), w["itype"], 4.0, 0.44)
        if not s:
            raise RuntimeError(L.x3do_last_error().decode())
        self.s = C.c_void_p(s)
        self.itype = w["itype"]
        self.w = w

    def init(self):
        C, L = self.C, self.L
        if self.itype == 3:
            L.x3do_solver_init_channel(self.s)
        elif self.itype == 5:
            from incompact3d_b200.cases import cylinder_geometry, NOBJMAX, NPIF, IZAP
            import numpy as np
            dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
            L.x3do_solver_init_cyl.argtypes = [C.c_void_p, C.c_double, C.c_double]
            L.x3do_solver_set_ibm.argtypes = [C.c_void_p, C.c_int, dp, dp]
            L.x3do_solver_set_ibm_geometry.argtypes = [C.c_void_p] + [C.c_int] * 4 + [ip, dp, dp, ip, ip]
            L.x3do_solver_init_cyl(self.s, 1.0, 1.0)
            ep, geo, _ = cylinder_geometry(self.dims, self.w["lens"], *self.w["cyl"])
            ubc = np.zeros(3)
            L.x3do_solver_set_ibm(self.s, 2, ep.ctypes.data_as(dp), ubc.ctypes.data_as(dp))
            for axis, (nobj, xi, xf, nip, nfp) in enumerate(geo):
                L.x3do_solver_set_ibm_geometry(self.s, axis, NOBJMAX, NPIF, IZAP, nobj.ctypes.data_as(ip), xi.ctypes.data_as(dp),
                                               xf.ctypes.data_as(dp), nip.ctypes.data_as(ip), nfp.ctypes.data_as(ip))
        else:
            L.x3do_solver_init_tgv(self.s)

    def step(self, k=1):
        if self.L.x3do_solver_step(self.s, k):
            raise RuntimeError(self.L.x3do_last_error().decode())

    def velocity(self):
        import numpy as np
        dp = self.C.POINTER(self.C.c_double)
        out = [np.zeros(self.dims, order="F") for _ in range(3)]
        self.L.x3do_solver_get_velocity(self.s, *[a.ctypes.data_as(dp) for a in out])
        return out

    def diagnostics(self):
        out = (self.C.c_double * 4)()
        self.L.x3do_solver_postprocess_tgv(self.s, out)
        return dict(eek=out[0], eps=out[1], eps2=out[2], enst=out[3])

    def close(self):
        self.L.x3do_solver_destroy(self.s)


ORACLE_WHAT = ("oracle C++/OpenMP restatement of the reference step (the Fortran/MPI build cannot be built on this image: no Fortran "
               "compiler, no MPI, 2DECOMP&FFT un-vendored)")


def run_reference(args):
    """the reference's CPU implementation of the path, all host cores, SAME workload as the GPU arm; every timed step is
    a full step of that workload, the number of timed steps is bounded by --ref-budget-s"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    o = Oracle(w)
    o.init()
    t0 = time.perf_counter()
    o.step(1)                                   # warm-up (first touch of every array)
    t_warm = time.perf_counter() - t0
    steps = int(max(1, min(args.steps, (args.ref_budget_s - t_warm) // max(t_warm, 1e-9))))
    t0 = time.perf_counter()
    o.step(steps)
    dtw = time.perf_counter() - t0
    diag = o.diagnostics() if w["itype"] == 0 else None
    o.close()
    npts = w["dims"][0] * w["dims"][1] * w["dims"][2]
    value = npts * steps / dtw
    sample = (f"{w['name']}: {steps} full steps timed after 1 warm-up step ({dtw:.1f} s; --steps {args.steps} capped by the "
              f"{args.ref_budget_s:.0f} s budget), {o.cores} OpenMP threads; {ORACLE_WHAT}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "steps_requested": args.steps, "ms_per_step": 1e3 * dtw / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": o.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "diagnostics_after_run": diag,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def alg_bytes(name, npts, nsp):
    """ALGORITHMIC bytes of one launch of a kernel class: npts = local grid points,
    nsp = local complex spectral modes"""
    if name.startswith("compact_"):
        return 16.0 * npts                      # read u, write t
    if name.startswith("accumulate_"):
        return 24.0 * npts                      # read u, read-modify-write t
    if name.startswith("momentum_z_face"):
        return None                             # k_zfix: rows next to the slab faces only
    if name.startswith("momentum_fused"):
        # 3 velocities in, 3 results out.  With the time integration folded in: u, v, w and the running sum in, the stored
        # right-hand side in (2 of 3 RK3 sub-steps), u, v, w out, the stored right-hand side out (2 of 3) = 104 B on average
        return (104.0 if "+intt" in name else 48.0) * npts
    if name.startswith("staggered_"):
        return 24.0 * npts                      # two inputs + one output, or one input + two outputs
    if name.startswith("elementwise"):
        return 104.0 * npts                     # intt of RK3: 13 array passes on average over the sub-steps
    if name.startswith("fft_z"):
        return 8.0 * npts + 16.0 * nsp
    if name.startswith("fft_y(") or name.startswith("fft_x_fwd+spectral"):
        return 32.0 * nsp                       # one read + one write of the spectral array (the x pass does two transforms and the factor)
    if name.startswith("fft_y_fwd+fft_x_fwd") or name.startswith("fft_x_inv+fft_y_inv"):
        return 64.0 * nsp                       # two passes
    if name.startswith("fft_xy") or name.startswith("poisson_spectral"):
        return 32.0 * nsp
    if name.startswith("transpose"):
        return 16.0 * npts
    return None


def _zext(x, i):
    try:
        return x.decomp_info(i)["zsz"][2]
    except Exception:
        return None


def make_solver(X3D, local, w, dims, world, rank, dist, nccl_unique_id):
    x = X3D(local)
    nx, ny, nz = dims
    if world > 1:
        # 2-D pencil decomposition with p_row=1, p_col=N (slabs): on NVSwitch every byte costs the same
        # whichever peer it goes to, and 1xN moves the fewest bytes (x<->y transposes become local)
        obj = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        x.decomp_init(nx, ny, nz, 1, world, rank, world, obj[0])
    x.solver_init(nx, ny, nz, ncl=w["ncl"], xlx=w["lens"][0], yly=w["lens"][1], zlz=w["lens"][2], re=w["re"], dt=w["dt"],
                  itimescheme=w["itimescheme"], isecondder=w["isecondder"], istret=w["istret"], beta=w["beta"], itype=w["itype"],
                  p_row=1, p_col=world)
    if w["itype"] == 3:
        x.solver_init_channel()
    elif w["itype"] == 5:
        from incompact3d_b200.cases import apply_cylinder
        apply_cylinder(x, dims, w["lens"], *w["cyl"])
        x.solver_init_cyl()
    else:
        x.solver_init_tgv()
    return x


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
    from incompact3d_b200 import X3D, nccl_unique_id

    w = workload(args)
    nx, ny, nz = w["dims"]
    npts = nx * ny * nz

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    x = make_solver(X3D, local, w, w["dims"], world, rank, dist, nccl_unique_id)
    stream = torch.cuda.ExternalStream(x.stream)

    def barrier():
        x.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- N > 1: the production transposes (peer stores over NVLink between library-owned pencils, every copy mode)
    #      checked bit for bit on the decomposition of this run before anything is timed
    bit_exact = None
    if world > 1:
        # complex on the spectral decomposition of the Poisson solver (nz/2+1 planes; registered by solver_init), real on the
        # main one; y<->z only: with slabs (p_row = 1) the x<->y transposes are local copies.  The larger buffers go first.
        sp = next((i for i in range(1, 8) if _zext(x, i) == nz // 2 + 1), None)
        bad = x.transpose_selftest(sp, whiches=(1, 2), kinds=(1,)) if sp is not None else 0
        bad += x.transpose_selftest(0, whiches=(1, 2), kinds=(0,))
        bit_exact = allmax(bad) == 0.0

    for _ in range(args.warmup):
        x.solver_step(1)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = x.launch_count
    b0 = x.decomp_stats()[0] if world > 1 else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        x.solver_step(1)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = allmax(e0.elapsed_time(e1))   # device time, max over ranks
    launches = x.launch_count - l0
    nv_bytes = (x.decomp_stats()[0] - b0) / args.steps if world > 1 else 0
    clk = clocks.stop()
    value = npts * args.steps / (ms * 1e-3)
    diag = x.solver_diagnostics_tgv() if w["itype"] == 0 else None

    # ---- roofline of the dominant kernel class, timed live in one extra instrumented step -------------
    roof = x.profile_step()
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("hbm_gbs", 6650.0)
    nzl = x._solver_shape[2]
    npl = nx * ny * max(nzl, 0)
    nspl = nx * ny * (nz // 2 + 1) / world
    for r in roof:
        ab = alg_bytes(r["name"], npl, nspl)
        r["alg_bytes_per_launch"] = ab
        r["frac_of_hbm_peak"] = (ab / (r["avg_ms"] * 1e-3) / 1e9 / peak) if (ab and r["avg_ms"] > 0 and not r["name"].startswith("transpose_p2p") and not r["name"].startswith("slab_ring")) else None
    tot_ms = sum(r["total_ms"] for r in roof)
    # dominant KERNEL: classes are per direction / role; group them by the kernel that runs (the name in parentheses)
    def kern(r):
        return r["name"][r["name"].index("(") + 1:].split(",")[0].rstrip(")") if "(" in r["name"] else r["name"]
    groups = {}
    for r in roof:
        if not r["name"].startswith("transpose_p2p") and r["alg_bytes_per_launch"]:
            groups.setdefault(kern(r), []).append(r)
    kname, members = max(groups.items(), key=lambda kv: sum(r["total_ms"] for r in kv[1]))
    g_ms = sum(r["total_ms"] for r in members)
    g_bytes = sum(r["alg_bytes_per_launch"] * r["count"] for r in members)
    k = {"name": f"{kname}: " + " + ".join(r["name"] for r in members), "count": sum(r["count"] for r in members), "total_ms": g_ms,
         "avg_ms": g_ms / sum(r["count"] for r in members), "alg_bytes_per_launch": g_bytes / sum(r["count"] for r in members)}
    achieved = g_bytes / (g_ms * 1e-3) / 1e9
    traffic, tsrc = None, None
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tj) and world == 1 and args.case == "tgv":
        T = json.load(open(tj))
        ent = T.get("kernels", {}).get(kname)
        if ent and T.get("n") == args.n:
            traffic, tsrc = ent["dram_bytes_per_launch"], f"profiles/ncu_traffic.json ({T.get('source')})"
    roofline = {"bound": "hbm", "kernel": k["name"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": k["alg_bytes_per_launch"],
                "note": "kernel with the largest share of the step (all classes considered, grouped by the kernel that runs; average over its launches); algorithmic bytes per class from "
                        "DESIGN.md section 4; per-launch CUDA events on the launching stream during one extra instrumented step",
                "launches_per_step": k["count"], "share_of_step": k["total_ms"] / tot_ms, "classes": roof}
    nvlink = None
    if world > 1:
        tms = sum(r["total_ms"] for r in roof if r["name"].startswith("transpose_") or r["name"].startswith("slab_ring_exchange"))
        gbs = nv_bytes / (tms * 1e-3) / 1e9 if tms > 0 else None
        nvlink = {"bytes_per_gpu_per_step": nv_bytes, "ms": tms, "gbs_per_direction": gbs, "frac_of_900": gbs / 900.0 if gbs else None,
                  "share_of_step": tms / tot_ms,
                  "note": "bytes each GPU stores into its peers' pencils per step (transposes, and the halo / carry planes of the slab z kernels) / device "
                          "time of those scopes (between their two flag barriers) in the instrumented step; 900 GB/s = NVLink 5 per direction"}

    # ---- e2e: host jobs through the C ABI, H2D + D2H of every step inside the timed region ---------------------
    e2e = None
    if not args.no_e2e:
        shape = (max(nzl, 1), ny, nx)
        sets = [[torch.empty(shape, dtype=torch.float64).pin_memory() for _ in range(3)] for _ in range(3)]
        x.solver_get_velocity(*sets[0])
        for s in sets[1:]:
            for a, b in zip(s, sets[0]):
                a.copy_(b)
        # strictly serial: H2D -> step -> D2H, one job at a time
        x.solver_set_velocity(*sets[0]); x.solver_step(1); x.solver_get_velocity(*sets[0])
        barrier()
        ks = 2
        t0 = time.perf_counter()
        for _ in range(ks):
            x.solver_set_velocity(*sets[0])
            x.solver_step(1)
            x.solver_get_velocity(*sets[0])
        barrier()
        t_serial = allmax(time.perf_counter() - t0) / ks
        serial = {"value": npts / t_serial, "ms_per_step": 1e3 * t_serial,
                  "note": "x3d_solver_set_velocity(host) + x3d_solver_step + x3d_solver_get_velocity(host), one after the other"}
        if w["itimescheme"] not in (1, 5):
            # Adams-Bashforth history lives on the device: host jobs are not independent, the serial figure is the e2e figure
            e2e = {"value": serial["value"], "unit": UNIT, "h2d_bytes_per_step": 3 * npts * 8, "d2h_bytes_per_step": 3 * npts * 8, "steps": ks,
                   "ms_per_step": serial["ms_per_step"], "note": serial["note"]}
    if e2e is None and not args.no_e2e:
        # pipelined: three host-resident members advanced in turn, copies overlap the neighbouring jobs' kernels
        k2 = max(6, min(args.steps, 12))
        for j in range(3):
            x.solver_advance_host(sets[j % 3], sets[j % 3], 1)
        x.solver_host_sync()
        barrier()
        t0 = time.perf_counter()
        for j in range(k2):
            x.solver_advance_host(sets[j % 3], sets[j % 3], 1)
        x.solver_host_sync()
        barrier()
        te = allmax(time.perf_counter() - t0)
        e2e = {"value": npts * k2 / te, "unit": UNIT, "h2d_bytes_per_step": 3 * npts * 8, "d2h_bytes_per_step": 3 * npts * 8,
               "steps": k2, "ms_per_step": 1e3 * te / k2,
               "serial": serial,
               "note": "x3d_solver_advance_host(host in, host out) per step: H2D of the step's three velocity arrays from pinned host memory, "
                       "the step, D2H of the result, all inside the timed region; three host-resident ensemble members are advanced in "
                       "turn (job j reads and writes member j mod 3), so the copies run on their own streams beside the kernels of the "
                       "neighbouring jobs; wall clock between barriers, max over ranks"}
        del sets
    free_b, total_b = torch.cuda.mem_get_info()
    hbm_used = round(allmax((total_b - free_b) / 1e9), 2)   # memory budget: device memory in use on the fullest GPU while the solver exists
    x.close()

    # ---- parity + cpu_baseline: GPU solver (same rank layout) vs the oracle from the same state, bounded size ----
    parity, cpu = None, None
    if not args.no_cpu_baseline:
        pd = (args.parity_n,) * 3 if args.case == "tgv" else w["dims"]
        wp = dict(w)
        if args.case == "tgv":
            wp["dt"] = 0.005 * 64.0 / args.parity_n
        nps = 3
        xp = make_solver(X3D, local, wp, pd, world, rank, dist, nccl_unique_id)
        xp.solver_step(nps)
        gu = xp.solver_get_velocity()
        gd = xp.solver_diagnostics_tgv() if w["itype"] == 0 else {}
        z0, nzlp = xp.solver_zstart, xp._solver_shape[2]
        xp.close()
        ref = [torch.empty((pd[2], pd[1], pd[0]), dtype=torch.float64, device="cuda") for _ in range(3)]
        rd = torch.zeros(4, dtype=torch.float64, device="cuda")
        if rank == 0:
            try:
                o = Oracle(wp, pd)
                o.init()
                t0 = time.perf_counter(); o.step(1); t1 = time.perf_counter()
                o.step(nps - 1)
                t2 = time.perf_counter()
                for r_, a in zip(ref, o.velocity()):
                    r_.copy_(torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))))
                if w["itype"] == 0:
                    od = o.diagnostics()
                    rd.copy_(torch.tensor([od["eek"], od["eps"], od["eps2"], od["enst"]], dtype=torch.float64))
                o.close()
                pn = pd[0] * pd[1] * pd[2]
                cpu = {"value": pn * (nps - 1) / (t2 - t1), "unit": UNIT, "cores": o.cores, "kind": "port",
                       "sample": f"{wp['name'] if args.case != 'tgv' else f'periodic TGV {args.parity_n}^3 (same Re, RK3, CFL-scaled dt)'}: "
                                 f"{nps - 1} full steps timed ({t2 - t1:.1f} s) after 1 warm-up step ({t1 - t0:.1f} s), {o.cores} OpenMP threads; "
                                 f"{ORACLE_WHAT}.  The reference arm (--impl reference) times the full-size workload."}
            except Exception as e:  # the baseline is reported, never fatal; parity then stays unproven
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        if world > 1:
            for r_ in ref:
                dist.broadcast(r_, src=0)
            dist.broadcast(rd, src=0)
        scale = max(float(r_.abs().max()) for r_ in ref)
        err = 0.0
        if nzlp > 0 and scale > 0:
            for g, r_ in zip(gu, ref):
                gt = torch.from_numpy(np.ascontiguousarray(g.transpose(2, 1, 0))).cuda()
                err = max(err, float((gt - r_[z0:z0 + nzlp]).abs().max()) / scale)
        err = allmax(err)
        drel = None
        if w["itype"] == 0 and scale > 0:
            got = np.array([gd["eek"], gd["eps"], gd["eps2"], gd["enst"]])
            drel = float(np.abs(got / rd.cpu().numpy() - 1).max())
        ok = scale > 0 and err < PARITY_TOL and (drel is None or drel < 1e-9) and (not gd or abs(gd["divmax"]) < 1e-10)
        parity = {"workload": f"{pd[0]}x{pd[1]}x{pd[2]} {args.case}, {nps} steps from the initial field, GPU solver on {world} rank(s) vs the CPU oracle",
                  "max_du_over_max_u": err, "tol": PARITY_TOL, "diagnostics_rel_err": drel, "gpu_divmax": gd.get("divmax") if gd else None,
                  "ok": bool(ok)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"],
                   "parallelism": f"{world} GPU" + ("" if world == 1 else f", 2-D pencil decomposition p_row=1 x p_col={world} (slabs), transposes = peer stores into the owners' pencils over NVLink (CUDA IPC) between device-side flag barriers; z part of the momentum terms on the slabs themselves (halo + carry planes) when the slabs are equal and >= 64 planes"),
                   "l2": "fields are far larger than the 126 MB L2 (1 GiB each at 512^3); no flush needed" if npts >= 2 ** 26 else "fields fit the 126 MB L2: an L2-resident, launch-bound configuration",
                   "hbm_gb_in_use_per_gpu": hbm_used,
                   "diagnostics_after_run": diag},
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "transposes_bit_exact": bit_exact, "nvlink": nvlink,
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if (parity and not parity["ok"]) or bit_exact is False:
        sys.exit(3)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
