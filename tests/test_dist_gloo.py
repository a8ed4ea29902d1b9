"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the uniform row partition, the column
remapping and the halo plan that ga_set_csr_dist consumes.  Each rank builds its plan with
torch.distributed collectives; the sharded SpMV is then emulated in numpy exactly as the library
performs it (pack -> exchange -> local rows over [x_local | halo]) and compared with the global A x."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, q):
    sys.path.insert(0, ROOT)
    import importlib

    import torch.distributed as dist

    import __graft_entry__ as entry

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    if kind == "lap":
        m = 37
        full = syn.laplacian2d(m)
    else:
        full = syn.sprand_sym(1500, 4.0, seed=5)
    A = syn.to_scipy(full)
    n = A.shape[0]
    blk, parts = gd.block_partition(n, world)
    r0, r1 = parts[rank]
    Aloc = A[r0:r1].tocsr()
    Aloc.sort_indices()
    rows = (Aloc.indptr.astype(np.int64), Aloc.indices.astype(np.int32), Aloc.data, (r1 - r0, n))
    plan = gd.halo_plan(rows[1], n, rank, world)
    # consistency of the plan
    assert plan["recv_counts"].sum() == plan["nhalo"] and plan["send_counts"][rank] == 0 and plan["recv_counts"][rank] == 0
    counts = [None] * world
    dist.all_gather_object(counts, (plan["send_counts"].tolist(), plan["recv_counts"].tolist()))
    for a in range(world):
        for b in range(world):
            assert counts[a][0][b] == counts[b][1][a]          # what a sends to b == what b receives from a
    x = np.cos(np.arange(n) * 0.37) + 0.1
    xl = x[r0:r1]
    # "pack": per destination peer
    off = np.concatenate([[0], np.cumsum(plan["send_counts"])])
    sendbuf = {p: xl[plan["send_idx"][off[p]:off[p + 1]]] for p in range(world)}
    blocks = [None] * world
    dist.all_gather_object(blocks, {"x": xl, "sendbuf": sendbuf})
    y = gd.emulate_sharded_spmv(rows, plan, blocks, rank)
    ok = np.allclose(y, (A @ x)[r0:r1], rtol=1e-13, atol=1e-13)
    if kind == "lap" and world == 2:
        # a stencil only needs one grid line from each neighbour
        assert plan["nhalo"] == 37
    res = [None] * world
    dist.all_gather_object(res, bool(ok))
    if rank == 0:
        q.put(all(res))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "lap"), (3, "lap"), (2, "sprand"), (3, "sprand")])
def test_halo_plan_gloo(world, kind):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world * 7 + (0 if kind == "lap" else 3)
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True


def _worker_normal(rank, world, port, wide, q):
    sys.path.insert(0, ROOT)
    import importlib

    import torch.distributed as dist

    import __graft_entry__ as entry

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    A = syn.to_scipy(syn.tall_sparse(2300, 170, 6))
    if wide:
        A = A.T.tocsr()                      # m <= n: OP = A A^H, T = A^H
    rows, plan, n, (t0, t1) = gd.shard_normal(A, rank, world)
    assert n == 170 and rows[3] == (t1 - t0, 170)
    x = np.sin(np.arange(n) * 0.11) + 0.3
    r0, r1 = plan["row_range"]
    xl = x[r0:r1]
    off = np.concatenate([[0], np.cumsum(plan["send_counts"])])
    sendbuf = {p: xl[plan["send_idx"][off[p]:off[p + 1]]] for p in range(world)}
    gathered = [None] * world
    dist.all_gather_object(gathered, {"x": xl, "sendbuf": sendbuf})
    _, z = gd.emulate_sharded_normal(rows, plan, gathered, rank)
    zs = [None] * world
    dist.all_gather_object(zs, z)
    plans = [None] * world
    dist.all_gather_object(plans, {k: plan[k] for k in ("send_counts", "send_idx", "send_dest_off", "n_pad", "row_range")})
    y = gd.reduce_scatter_emulated(plan, zs, rank, plans)
    ref = (A @ (A.T @ x)) if wide else (A.T @ (A @ x))
    ok = np.allclose(y, ref[r0:r1], rtol=1e-12, atol=1e-12)
    res = [None] * world
    dist.all_gather_object(res, bool(ok))
    if rank == 0:
        q.put(all(res))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,wide", [(2, False), (3, False), (2, True)])
def test_sharded_normal_operator_gloo(world, wide):
    """plan + data flow of ga_set_normal_csr_dist: halo of x, T_loc, T_loc^H, rank-ordered reduce-scatter"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world * 5 + (1 if wide else 0)
    procs = [ctx.Process(target=_worker_normal, args=(r, world, port, wide, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) is True


def test_block_partition():
    sys.path.insert(0, ROOT)
    import importlib

    import __graft_entry__ as entry

    entry.load_package()
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    blk, parts = gd.block_partition(10, 3)
    assert blk == 4 and parts == [(0, 4), (4, 8), (8, 10)]
    assert gd.padded_block(16_000_000, 8) == 2_000_128 - 2_000_128 % 256 or gd.padded_block(16_000_000, 8) % 256 == 0
    need = gd.local_need(np.array([0, 5, 9, 4, 7]), 4, 8)
    assert list(need) == [0, 9]


def test_float32_blocks_are_padded_to_512_rows():
    """Float32 contexts pad the local block to whole 2 KiB column segments = 512 rows (row_pad in csrc/ga_api.cu); the
    halo plan numbers remote columns from that padded size."""
    sys.path.insert(0, ROOT)
    import importlib

    import __graft_entry__ as entry

    pkg = entry.load_package()
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    assert gd.row_pad_for(None) == gd.row_pad_for(np.float64) == gd.row_pad_for(pkg.GA_C64) == 256
    assert gd.row_pad_for(np.float32) == gd.row_pad_for("f32") == gd.row_pad_for(pkg.GA_F32) == 512
    assert gd.padded_block(22500, 2) == 11264 and gd.padded_block(22500, 2, np.float32) == 11264
    assert gd.padded_block(1300, 2) == 768 and gd.padded_block(1300, 2, np.float32) == 1024
    # a two-rank plan computed serially (all_gather_object replaced by the precomputed needs)
    m, world = 36, 2
    n = m * m
    blk, parts = gd.block_partition(n, world)
    cols = [syn.laplacian2d(m, row_range=parts[r])[1] for r in range(world)]
    needs = [gd.local_need(cols[r], *parts[r]) for r in range(world)]
    for dt, pad in ((None, 768), (np.float32, 1024)):
        for r in range(world):
            plan = gd.halo_plan(cols[r], n, r, world, all_gather_object=lambda obj: needs, dtype=dt)
            assert plan["n_pad"] == pad and plan["nhalo"] == m
            remote = (cols[r] < parts[r][0]) | (cols[r] >= parts[r][1])
            assert np.all(plan["col_local"][remote] >= pad) and np.all(plan["col_local"][~remote] < blk)


@pytest.mark.parametrize("world", [2, 3])
def test_shard_generalized_common_halo_plan(world):
    """dist.shard_generalized: A and B share one halo plan (the union of their remote columns).  Emulate the exchange
    serially: every rank's halo = the entries the owners' send lists deliver, and the [own | halo] products of both
    matrices must reproduce A x and B x."""
    sys.path.insert(0, ROOT)
    import importlib

    import scipy.sparse as sp

    import __graft_entry__ as entry

    entry.load_package()
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    m = 23
    A = syn.to_scipy(syn.laplacian2d(m))
    n = m * m
    rng = np.random.default_rng(5)
    E = sp.random(n, n, density=3.0 / n, random_state=7, format="csr")
    B = (sp.identity(n) * 4.0 + 0.1 * (E + E.T)).tocsr()          # a different pattern than A's: the union matters
    blk, parts = gd.block_partition(n, world)
    needs = []
    for r in range(world):
        r0, r1 = parts[r]
        cols = np.concatenate([A[r0:r1].indices, B[r0:r1].indices])
        needs.append(gd.local_need(cols, r0, r1))
    out = [gd.shard_generalized(A, B, r, world, all_gather_object=lambda obj: needs) for r in range(world)]
    x = rng.standard_normal(n)
    for r in range(world):
        ra, rb, plan, (r0, r1) = out[r]
        assert (r0, r1) == parts[r] and ra[3] == rb[3] == (r1 - r0, n)
        assert len(plan["col_local"]) == len(ra[1]) and len(plan["col_local_b"]) == len(rb[1])
        # what the peers deliver, owner by owner, in the order of their send lists for this rank
        halo = []
        for q in range(world):
            pq = out[q][2]
            off = int(np.sum(pq["send_counts"][:r]))
            idx = pq["send_idx"][off: off + int(pq["send_counts"][r])]
            assert int(pq["send_dest_off"][r]) == sum(len(h) for h in halo)
            halo.append(x[parts[q][0] + idx])
        halo = np.concatenate(halo)
        assert len(halo) == plan["nhalo"] and np.array_equal(halo, x[plan["need"]])
        xext = np.concatenate([x[r0:r1], np.zeros(plan["n_pad"] - (r1 - r0)), halo])
        for (rowptr, col, val, _), cl, M in ((ra, plan["col_local"], A), (rb, plan["col_local_b"], B)):
            y = np.zeros(r1 - r0)
            np.add.at(y, np.repeat(np.arange(r1 - r0), np.diff(rowptr)), val * xext[cl])
            assert np.allclose(y, (M @ x)[r0:r1], rtol=1e-14, atol=1e-14)
