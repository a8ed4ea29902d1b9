"""Row sharding of the Lanczos problem over one process per GPU (host-side plumbing).

The data distribution is P_ARPACK's (reference comments, src/dsaup2.jl:911-926): rows of A, V,
resid are block-partitioned, the small ncv-sized arrays are replicated.  This module computes,
with torch.distributed (any backend: NCCL on the GPUs, gloo in the CPU tests), what the C library
needs for the sharded SpMV: the uniform block partition, the column remapping and the halo plan
(who needs which x entries from whom).  The per-step exchange itself runs inside the library.
"""
import numpy as np

ROW_PAD = 256  # must match kRowPad in csrc/ga_ctx.cuh


def block_partition(n, world):
    """Uniform contiguous row blocks: rank r owns [r*blk, min(n, (r+1)*blk))."""
    blk = (n + world - 1) // world
    return blk, [(r * blk, min(n, (r + 1) * blk)) for r in range(world)]


def padded_block(n, world):
    blk = (n + world - 1) // world
    return (blk + ROW_PAD - 1) // ROW_PAD * ROW_PAD


def local_need(col, r0, r1):
    """Sorted unique global columns referenced by this rank's rows that it does not own."""
    col = np.asarray(col, dtype=np.int64)
    remote = col[(col < r0) | (col >= r1)]
    return np.unique(remote)


def halo_plan(col, n, rank, world, all_gather_object=None):
    """Halo plan for this rank's rows (global column indices `col`).

    all_gather_object(obj) -> list of every rank's obj (default: torch.distributed.all_gather_object).
    Returns dict(col_local int32, nhalo, recv_counts int32[world], send_counts int32[world],
    send_idx int32[sum(send_counts)], need int64[nhalo])."""
    blk, parts = block_partition(n, world)
    r0, r1 = parts[rank]
    n_pad = padded_block(n, world)
    col = np.asarray(col, dtype=np.int64)
    need = local_need(col, r0, r1)
    owner = need // blk
    recv_counts = np.bincount(owner, minlength=world).astype(np.int32)
    # what every rank needs, so that each owner learns what to send
    if all_gather_object is None:
        import torch.distributed as dist

        def all_gather_object(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    needs = all_gather_object(need)
    send_counts = np.zeros(world, dtype=np.int32)
    send_dest_off = np.zeros(world, dtype=np.int32)   # where my entries start in rank q's halo
    send_idx = []
    for q in range(world):
        nq = np.asarray(needs[q], dtype=np.int64)
        mine = nq[(nq >= r0) & (nq < r1)]
        send_counts[q] = len(mine)
        send_dest_off[q] = int(np.searchsorted(nq, r0))
        send_idx.append((mine - r0).astype(np.int32))
    send_idx = np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.int32)
    # column remap: own -> g - r0 ; remote -> n_pad + position in `need`
    own = (col >= r0) & (col < r1)
    col_local = np.empty(len(col), dtype=np.int64)
    col_local[own] = col[own] - r0
    col_local[~own] = n_pad + np.searchsorted(need, col[~own])
    return {"col_local": np.ascontiguousarray(col_local, dtype=np.int32), "nhalo": int(len(need)),
            "recv_counts": recv_counts, "send_counts": send_counts, "send_dest_off": send_dest_off,
            "send_idx": np.ascontiguousarray(send_idx, dtype=np.int32), "need": need, "n_pad": n_pad,
            "row_range": (r0, r1)}


def setup_comm(ctx, comm_unique_id, rank, world, fused=True):
    """Rendezvous for one context (torch.distributed must be initialised): broadcast the NCCL unique
    id, then (fused=True) all-gather the cudaIpc handles so that the per-step collectives run inside
    the compute kernels over NVSwitch peer memory.  `comm_unique_id` = the package's function."""
    import torch.distributed as dist

    ids = [comm_unique_id().tobytes() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(np.frombuffer(ids[0], dtype=np.uint8).copy())
    if fused:
        mine = ctx.ipc_export().tobytes()
        allh = [None] * world
        dist.all_gather_object(allh, mine)
        ctx.ipc_import(np.frombuffer(b"".join(allh), dtype=np.uint8).copy())
    return ctx


def emulate_sharded_spmv(csr_rows, plan, x_blocks, rank):
    """numpy emulation of what the library does per OP*x on one rank (used by the CPU tests):
    peers' send buffers -> halo, then the local SpMV over [x_local | halo]."""
    rowptr, _, val, _ = csr_rows
    world = len(x_blocks)
    halo = []
    for q in range(world):
        if plan["recv_counts"][q]:
            halo.append(x_blocks[q]["sendbuf"][rank])
    halo = np.concatenate(halo) if halo else np.zeros(0, dtype=val.dtype)
    xl = x_blocks[rank]["x"]
    xext = np.concatenate([xl, np.zeros(plan["n_pad"] - len(xl), dtype=xl.dtype), halo])
    y = np.zeros(len(rowptr) - 1, dtype=np.result_type(val.dtype, xl.dtype))
    prod = val * xext[plan["col_local"]]
    np.add.at(y, np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr)), prod)
    return y
