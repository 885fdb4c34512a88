"""Host logic of the pencil decomposition and of the transpose exchange plans (CPU only):
 * block distribution rule of 2DECOMP&FFT (the LAST mod(n,p) ranks get one extra point);
 * every transpose, every process grid: pack -> all-to-all(v) -> unpack driven by the library's plan
   reproduces the destination pencils of a global array bit-exactly (simulated ranks);
 * the same exchange run for real on 2 ranks over torch.distributed/gloo."""
import os
import sys

import numpy as np
import pytest

from incompact3d_b200 import decomp_compute, transpose_plan

GRIDS = [(1, 1), (1, 2), (2, 1), (2, 2), (1, 4), (2, 4), (4, 2), (1, 8), (3, 2)]
DIMS = [(16, 12, 20), (13, 11, 9), (17, 8, 10)]


def pencil(G, info, which):
    st, en = info[which + "st"], info[which + "en"]
    return np.asfortranarray(G[st[0] - 1:en[0], st[1] - 1:en[1], st[2] - 1:en[2]])


def test_distribution_rule():
    info = [decomp_compute(10, 11, 9, 1, 4, r) for r in range(4)]
    assert [i["xsz"][2] for i in info] == [2, 2, 2, 3]          # 9 over 4: last rank gets the extra
    assert [i["zsz"][1] for i in info] == [2, 3, 3, 3]          # 11 over 4: last three
    assert [i["xst"][2] for i in info] == [1, 3, 5, 7]
    for i in info:
        assert i["xsz"][0] == 10 and i["ysz"][1] == 11 and i["zsz"][2] == 9


SRC = {"x_to_y": "x", "y_to_z": "y", "z_to_y": "z", "y_to_x": "y"}
DST = {"x_to_y": "y", "y_to_z": "z", "z_to_y": "y", "y_to_x": "x"}
# (axis cut on the send side, axis cut on the receive side)
AXES = {"x_to_y": (0, 1), "y_to_x": (1, 0), "y_to_z": (1, 2), "z_to_y": (2, 1)}


def blocks(infos, peers, key, axis):
    return [(infos[p][key + "st"][axis] - 1, infos[p][key + "en"][axis]) for p in peers]


def pack(local, which, rank, infos, plan):
    sa, _ = AXES[which]
    my = infos[rank]
    # block m of the send array = the range the peer owns in the DESTINATION pencil, in my local coordinates
    dst = DST[which]
    out = np.zeros(sum(plan["scount"]))
    for m, p in enumerate(plan["peers"]):
        lo, hi = infos[p][dst + "st"][sa] - 1, infos[p][dst + "en"][sa]
        off = my[SRC[which] + "st"][sa] - 1
        sl = [slice(None)] * 3
        sl[sa] = slice(lo - off, hi - off)
        blk = local[tuple(sl)]
        assert blk.size == plan["scount"][m]
        out[plan["sdispl"][m]:plan["sdispl"][m] + blk.size] = blk.ravel(order="F")
    return out


def unpack(buf, which, rank, infos, plan):
    _, ra = AXES[which]
    my = infos[rank]
    dst, src = DST[which], SRC[which]
    shape = my[dst + "sz"]
    out = np.full(shape, np.nan, order="F")
    for m, p in enumerate(plan["peers"]):
        lo, hi = infos[p][src + "st"][ra] - 1, infos[p][src + "en"][ra]
        off = my[dst + "st"][ra] - 1
        sl = [slice(None)] * 3
        sl[ra] = slice(lo - off, hi - off)
        bshape = list(shape)
        bshape[ra] = hi - lo
        n = int(np.prod(bshape))
        assert n == plan["rcount"][m]
        out[tuple(sl)] = buf[plan["rdispl"][m]:plan["rdispl"][m] + n].reshape(bshape, order="F")
    return out


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("dims", DIMS)
def test_transposes_with_simulated_ranks(grid, dims):
    p_row, p_col = grid
    n = p_row * p_col
    nx, ny, nz = dims
    G = np.arange(nx * ny * nz, dtype=np.float64).reshape(dims, order="F") + 0.25
    infos = [decomp_compute(nx, ny, nz, p_row, p_col, r) for r in range(n)]
    for which in ("x_to_y", "y_to_z", "z_to_y", "y_to_x"):
        plans = [transpose_plan(nx, ny, nz, p_row, p_col, r, which) for r in range(n)]
        sends = [pack(pencil(G, infos[r], SRC[which]), which, r, infos, plans[r]) for r in range(n)]
        for r in range(n):
            P = plans[r]
            assert list(P["send_dims"]) == infos[r][SRC[which] + "sz"] and list(P["recv_dims"]) == infos[r][DST[which] + "sz"]
            recv = np.zeros(sum(P["rcount"]))
            for m, p in enumerate(P["peers"]):
                Q = plans[p]
                mm = Q["peers"].index(r)                       # my slot in the peer's plan
                assert Q["scount"][mm] == P["rcount"][m]
                recv[P["rdispl"][m]:P["rdispl"][m] + P["rcount"][m]] = sends[p][Q["sdispl"][mm]:Q["sdispl"][mm] + Q["scount"][mm]]
            got = unpack(recv, which, r, infos, P)
            assert np.array_equal(got, pencil(G, infos[r], DST[which])), (grid, dims, which, r)


def _gloo_worker(rank, world, port, dims, grid, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny, nz = dims
        p_row, p_col = grid
        G = np.arange(nx * ny * nz, dtype=np.float64).reshape(dims, order="F") + 0.5
        infos = [decomp_compute(nx, ny, nz, p_row, p_col, r) for r in range(world)]
        ok = True
        for which in ("x_to_y", "y_to_z", "z_to_y", "y_to_x"):
            P = transpose_plan(nx, ny, nz, p_row, p_col, rank, which)
            send = pack(pencil(G, infos[rank], SRC[which]), which, rank, infos, P)
            # all-to-all(v) over the whole world (ranks outside the group exchange nothing)
            ins = [torch.zeros(0, dtype=torch.float64) for _ in range(world)]
            outs = [torch.zeros(0, dtype=torch.float64) for _ in range(world)]
            for m, p in enumerate(P["peers"]):
                ins[p] = torch.from_numpy(send[P["sdispl"][m]:P["sdispl"][m] + P["scount"][m]].copy())
                outs[p] = torch.zeros(P["rcount"][m], dtype=torch.float64)
            reqs = []
            for p in range(world):
                if p == rank:
                    outs[p].copy_(ins[p])
                    continue
                if ins[p].numel():
                    reqs.append(dist.isend(ins[p], p))
                if outs[p].numel():
                    reqs.append(dist.irecv(outs[p], p))
            for r_ in reqs:
                r_.wait()
            recv = np.zeros(sum(P["rcount"]))
            for m, p in enumerate(P["peers"]):
                recv[P["rdispl"][m]:P["rdispl"][m] + P["rcount"][m]] = outs[p].numpy()
            got = unpack(recv, which, rank, infos, P)
            ok = ok and np.array_equal(got, pencil(G, infos[rank], DST[which]))
            dist.barrier()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid", [(1, 2), (2, 1)])
def test_transposes_over_gloo_world2(grid):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if grid == (1, 2) else 1)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, (13, 11, 9), grid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
