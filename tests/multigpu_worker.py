"""Run under torchrun (one rank per GPU): NCCL pencil transposes bit-exact against slicing the global
array, and the slab-decomposed solver against the single-rank oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from incompact3d_b200 import X3D, decomp_compute, nccl_unique_id
    from test_decomp_cpu import pencil

    def fresh_id():
        obj = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()

    # ---- transposes on every process grid with p_row*p_col == world -------------------------------
    dims = (21, 18, 20)
    nx, ny, nz = dims
    G = np.arange(nx * ny * nz, dtype=np.float64).reshape(dims, order="F") + 0.25
    Gc = G + 1j * (G[::-1] * 0.5 + 1.0)
    grids = [(pr, world // pr) for pr in (1, 2, 4, 8) if world % pr == 0 and pr <= world]
    for p_row, p_col in grids:
        x = X3D(local)
        x.decomp_init(nx, ny, nz, p_row, p_col, rank, world, fresh_id())
        info = decomp_compute(nx, ny, nz, p_row, p_col, rank)
        # the production data plane: peer stores / block copies / copy engines between library-owned pencils
        assert x.transpose_selftest() == 0, ("selftest", p_row, p_col, rank)
        # what bench.py does on the spectral decomposition: a second decomposition that splits unevenly (ranks hold pencils of
        # different sizes: their self-test buffers must still be (re)allocated by all ranks together), complex first, y<->z only
        sp = x.decomp_info_init(nx, ny, nz // 2 + 1)
        assert x.transpose_selftest(sp, whiches=(1, 2), kinds=(1,)) == 0, ("selftest sp", p_row, p_col, rank)
        assert x.transpose_selftest(0, whiches=(1, 2), kinds=(0,)) == 0, ("selftest after sp", p_row, p_col, rank)
        for arr, cplx in ((G, False), (Gc, True)):
            tdt = torch.complex128 if cplx else torch.float64
            for name, s, d in (("transpose_x_to_y", "x", "y"), ("transpose_y_to_z", "y", "z"),
                               ("transpose_z_to_y", "z", "y"), ("transpose_y_to_x", "y", "x")):
                src = dev(pencil(arr, info, s))
                dst = torch.zeros(tuple(reversed(info[d + "sz"])), dtype=tdt, device="cuda")
                getattr(x, name)(src, dst)
                x.sync()
                got = dst.cpu().numpy().transpose(2, 1, 0)
                assert np.array_equal(got, pencil(arr, info, d)), (p_row, p_col, name, cplx, rank)
        x.close()
        dist.barrier()

    # ---- slab-decomposed solver vs the single-rank oracle -----------------------------------------
    import oracle_lib as ol
    from test_oracle_tgv import make_solver
    # torch.distributed.run exports OMP_NUM_THREADS=1: give the oracle this rank's share of the host cores instead
    Lo = ol.lib()
    Lo.x3do_set_threads.argtypes = [C.c_int]
    Lo.x3do_set_threads(max(1, len(os.sched_getaffinity(0)) // world))
    # the last case has line lengths for which the fused momentum kernels (and their reduce-add accumulation across
    # the z -> y transposes) run
    # X3D_P2P_MODE (read by x3d_decomp_init): the block copies of the y<->z transposes through the vector-copy kernel
    # (what these small pencils take by default) and through the copy engines (what 512^3 pencils take)
    for nn, ncl, p2p_mode, overlap in (((32, 24, 40), (0,) * 6, None, "0"),
                                       ((64, 64, 128), (0,) * 6, None, "0"),   # power-of-two periodic mesh: the hand-written FFT passes, 65 spectral planes split unevenly ((33, 25, 33), (1,) * 6, "1", "0"), ((24, 176, 168), (0,) * 6, None, "0"),
                                       ((24, 176, 168), (0,) * 6, "1", "1"), ((176, 176, 168), (0,) * 6, None, "2"),
                                       # equal slabs of >= 64 planes: the z part of the momentum terms runs on the slabs themselves
                                       # (k_mom_slab + k_zfix, halo and carry planes through ring_exchange) instead of through transposes
                                       ((32, 168, 256), (0,) * 6, None, "0"), ((176, 168, 256), (0,) * 6, None, "0"),
                                       ((32, 168, 512), (0,) * 6, None, "0")):   # slabs of 256 / 128 / 64 planes on 2 / 4 / 8 ranks
        length = 2 * np.pi
        if p2p_mode is None:
            os.environ.pop("X3D_P2P_MODE", None)
        else:
            os.environ["X3D_P2P_MODE"] = p2p_mode
        # X3D_OVERLAP (read by x3d_solver_init): forward velocity transposes on a second stream, for the last case
        os.environ["X3D_OVERLAP"] = overlap   # 2: the y kernel beside the forward transposes, x + intt with the z part as an extra term
        x = X3D(local)
        x.decomp_init(*nn, 1, world, rank, world, fresh_id())
        x.solver_init(*nn, ncl=ncl, xlx=length, yly=length, zlz=length, re=1600.0, dt=0.002, p_row=1, p_col=world)
        x.solver_init_tgv()
        z0, nzl = x.solver_zstart, x._solver_shape[2]
        Ls, s = make_solver(n=nn, ncl=ncl, length=length, re=1600.0, dt=0.002)
        Ls.x3do_solver_init_tgv(s)
        x.solver_step(2)
        assert Ls.x3do_solver_step(s, 2) == 0
        dp = C.POINTER(C.c_double)
        ru, rv, rw = (np.zeros(nn, order="F") for _ in range(3))
        Ls.x3do_solver_get_velocity(s, ru.ctypes.data_as(dp), rv.ctypes.data_as(dp), rw.ctypes.data_as(dp))
        gu, gv, gw = x.solver_get_velocity()
        scale = max(np.abs(ru).max(), np.abs(rv).max(), np.abs(rw).max())
        for a, b in ((gu, ru), (gv, rv), (gw, rw)):
            err = np.abs(a - b[:, :, z0:z0 + nzl]).max() / scale if nzl else 0.0
            assert err < 1e-11, (nn, ncl, rank, err)
        out = (C.c_double * 4)()
        Ls.x3do_solver_postprocess_tgv(s, out)
        d = x.solver_diagnostics_tgv()
        got = np.array([d["eek"], d["eps"], d["eps2"], d["enst"]])
        assert np.abs(got / np.array(out[:]) - 1).max() < 1e-10, (got, out[:])
        assert abs(d["divmax"]) < 1e-11
        if nn[1] >= 168 and os.environ.get("X3D_FUSED", "1") != "0":
            names = {r["name"] for r in x.profile_step(1)}
            slab = nn[2] % world == 0 and (nn[2] // world) % 8 == 0 and 64 <= nn[2] // world <= 544 and os.environ.get("X3D_SLABZ", "1") != "0"
            zname = "momentum_fused_z_slab(k_mom_slab)" if slab else "momentum_fused_z(k_mom_pair)"
            assert "momentum_fused_y(k_mom_pair)" in names and zname in names, names
            assert ("slab_ring_exchange(k_p2p_blocks)" in names) == slab, names
        x.close()
        dist.barrier()
    if rank == 0:
        print("MULTIGPU OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
