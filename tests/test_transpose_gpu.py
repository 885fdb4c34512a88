"""GPU tests of the pencil transposes.

 * pack / unpack CUDA kernels for every transpose on simulated process grids (all ranks emulated on
   one GPU, the all-to-all done by copying the packed blocks): bit-exact against slicing the global
   array, real and complex, uneven splits;
 * single-rank x3d_transpose_* is a bit-exact copy;
 * with >= 2 GPUs: the real NCCL exchange and the slab-decomposed solver (tests/multigpu_worker.py
   under torchrun)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from incompact3d_b200 import decomp_compute, transpose_plan
from test_decomp_cpu import DST, SRC, pencil

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("grid", [(1, 2), (2, 1), (2, 2), (1, 4), (2, 4), (4, 2), (1, 8), (3, 2)])
@pytest.mark.parametrize("cplx", [False, True])
def test_pack_unpack_kernels_simulated_ranks(grid, cplx):
    import torch
    from incompact3d_b200 import X3D
    p_row, p_col = grid
    n = p_row * p_col
    dims = (21, 14, 18)
    nx, ny, nz = dims
    G = np.arange(nx * ny * nz, dtype=np.float64).reshape(dims, order="F") + 0.25
    if cplx:
        G = G + 1j * (G[::-1, :, :] * 0.5 + 3.0)
    ctxs = []
    for r in range(n):
        x = X3D(0)
        x.decomp_init(nx, ny, nz, p_row, p_col, r, n, None)   # no NCCL id: pack/unpack only
        ctxs.append(x)
    infos = [decomp_compute(nx, ny, nz, p_row, p_col, r) for r in range(n)]
    tdt = torch.complex128 if cplx else torch.float64

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 1, 0))).cuda()

    for which in ("x_to_y", "y_to_z", "z_to_y", "y_to_x"):
        plans = [transpose_plan(nx, ny, nz, p_row, p_col, r, which) for r in range(n)]
        sends = []
        for r in range(n):
            src = dev(pencil(G, infos[r], SRC[which]))
            sb = torch.zeros(sum(plans[r]["scount"]), dtype=tdt, device="cuda")
            ctxs[r].transpose_pack(which, src, sb, 0, cplx)
            ctxs[r].sync()
            sends.append(sb)
        for r in range(n):
            P = plans[r]
            rb = torch.zeros(sum(P["rcount"]), dtype=tdt, device="cuda")
            for m, p in enumerate(P["peers"]):
                Q = plans[p]
                mm = Q["peers"].index(r)
                rb[P["rdispl"][m]:P["rdispl"][m] + P["rcount"][m]] = sends[p][Q["sdispl"][mm]:Q["sdispl"][mm] + Q["scount"][mm]]
            shape = infos[r][DST[which] + "sz"]
            dst = torch.zeros(tuple(reversed(shape)), dtype=tdt, device="cuda")
            ctxs[r].transpose_unpack(which, rb, dst, 0, cplx)
            ctxs[r].sync()
            got = dst.cpu().numpy().transpose(2, 1, 0)
            assert np.array_equal(got, pencil(G, infos[r], DST[which])), (grid, which, r)
    for x in ctxs:
        x.close()


def test_single_rank_transposes_are_exact_copies():
    from incompact3d_b200 import X3D
    x = X3D(0)
    x.decomp_init(12, 10, 8)
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.uniform(-1, 1, (12, 10, 8)))
    for name in ("transpose_x_to_y", "transpose_y_to_z", "transpose_z_to_y", "transpose_y_to_x"):
        b = np.zeros_like(a)
        getattr(x, name)(a, b)
        assert np.array_equal(a, b)
    c = np.asfortranarray(a + 1j * a[::-1])
    d = np.zeros_like(c)
    x.transpose_x_to_y(c, d)
    assert np.array_equal(c, d)
    x.close()


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_nccl_transposes_and_solver(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU OK" in r.stdout, r.stdout[-3000:]
