"""Row sharding of the Lanczos problem over one process per GPU (host-side plumbing).

The data distribution is P_ARPACK's (reference comments, src/dsaup2.jl:911-926): rows of A, V,
resid are block-partitioned, the small ncv-sized arrays are replicated.  This module computes,
with torch.distributed (any backend: NCCL on the GPUs, gloo in the CPU tests), what the C library
needs for the sharded SpMV: the uniform block partition, the column remapping and the halo plan
(who needs which x entries from whom).  The per-step exchange itself runs inside the library.
"""
import numpy as np

ROW_PAD = 256  # must match kRowPad in csrc/ga_ctx.cuh


def row_pad_for(dtype=None):
    """Row padding of a context's local block: whole 2 KiB Gram-Schmidt column segments, i.e. 256 rows, 512 for Float32
    vectors (row_pad in csrc/ga_api.cu).  dtype: a numpy dtype / GA_* code / name as accepted by Context, or None."""
    if dtype is None:
        return ROW_PAD
    if dtype in (3, "f32", "float32") or (not isinstance(dtype, (int, str)) and np.dtype(dtype) == np.float32):
        return 2 * ROW_PAD
    return ROW_PAD


def block_partition(n, world):
    """Uniform contiguous row blocks: rank r owns [r*blk, min(n, (r+1)*blk))."""
    blk = (n + world - 1) // world
    return blk, [(r * blk, min(n, (r + 1) * blk)) for r in range(world)]


def padded_block(n, world, dtype=None):
    blk = (n + world - 1) // world
    pad = row_pad_for(dtype)
    return (blk + pad - 1) // pad * pad


def local_need(col, r0, r1):
    """Sorted unique global columns referenced by this rank's rows that it does not own."""
    col = np.asarray(col, dtype=np.int64)
    remote = col[(col < r0) | (col >= r1)]
    return np.unique(remote)


def halo_plan(col, n, rank, world, all_gather_object=None, dtype=None):
    """Halo plan for this rank's rows (global column indices `col`).  dtype: the context's element type (only Float32
    matters: its local blocks are padded to 512 rows, and remote columns are numbered from the padded block size).

    all_gather_object(obj) -> list of every rank's obj (default: torch.distributed.all_gather_object).
    Returns dict(col_local int32, nhalo, recv_counts int32[world], send_counts int32[world],
    send_idx int32[sum(send_counts)], need int64[nhalo])."""
    blk, parts = block_partition(n, world)
    r0, r1 = parts[rank]
    n_pad = padded_block(n, world, dtype)
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


def shard_normal(A, rank, world, all_gather_object=None, dtype=None):
    """Row sharding of ArpackNormalOp(A) (svds): T = A for m > n (At_side = :left), T = A^H otherwise;
    rank r owns a uniform block of T's rows and the usual block of the length-min(m, n) Lanczos vectors.
    A: scipy sparse matrix (every rank may pass the full matrix, or just its rows via `rows_of_T`); dtype: the context's
    element type when it is not A's (Float32 blocks are padded to 512 rows, see halo_plan).
    Returns (csr_rows, plan, n, (t0, t1)): csr_rows = (rowptr int64, col int32 global, val, (mt_local, n))."""
    import scipy.sparse as sp

    A = sp.csr_matrix(A)
    m, n0 = A.shape
    T = A if m > n0 else A.conj().T.tocsr()
    T.sort_indices()
    mt, n = T.shape
    tblk, tparts = block_partition(mt, world)
    t0, t1 = tparts[rank]
    Tl = T[t0:t1]
    rowptr = np.ascontiguousarray(Tl.indptr, dtype=np.int64)
    col = np.ascontiguousarray(Tl.indices, dtype=np.int32)
    plan = halo_plan(col, n, rank, world, all_gather_object, dtype=A.dtype if dtype is None else dtype)
    return (rowptr, col, Tl.data, (t1 - t0, n)), plan, n, (t0, t1)


def shard_generalized(A, B, rank, world, all_gather_object=None):
    """Row sharding of the generalised problem A x = lambda B x (ArpackSymmetricGeneralizedOp): rank r owns a uniform
    block of the rows of A and of B; both matrices share ONE halo plan, built from the union of their remote columns.
    A, B: scipy sparse matrices (n x n).  Returns (rows_a, rows_b, plan, (r0, r1)) with rows_* = (rowptr int64,
    col int32 global, val, (n_local, n)) and plan = halo_plan(...) whose "col_local" belongs to A's entries and
    "col_local_b" to B's."""
    import scipy.sparse as sp

    A = sp.csr_matrix(A); B = sp.csr_matrix(B)
    n = A.shape[0]
    if A.shape != (n, n) or B.shape != (n, n):
        raise ValueError("A and B must be square matrices of the same size")
    blk, parts = block_partition(n, world)
    r0, r1 = parts[rank]
    Al, Bl = A[r0:r1], B[r0:r1]
    Al.sort_indices(); Bl.sort_indices()
    ca = np.ascontiguousarray(Al.indices, dtype=np.int32)
    cb = np.ascontiguousarray(Bl.indices, dtype=np.int32)
    plan = halo_plan(np.concatenate([ca, cb]), n, rank, world, all_gather_object)
    both = plan["col_local"]
    plan["col_local"] = np.ascontiguousarray(both[: len(ca)])
    plan["col_local_b"] = np.ascontiguousarray(both[len(ca):])
    rows_a = (np.ascontiguousarray(Al.indptr, dtype=np.int64), ca, Al.data, (r1 - r0, n))
    rows_b = (np.ascontiguousarray(Bl.indptr, dtype=np.int64), cb, Bl.data, (r1 - r0, n))
    return rows_a, rows_b, plan, (r0, r1)


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


def emulate_sharded_normal(csr_rows, plan, gathered, rank):
    """numpy emulation of one OP*x = T^H (T x) of the row-sharded normal operator on one rank, step by
    step as the library performs it (used by the CPU tests).  `gathered[q]` = dict(x=x block of rank q,
    sendbuf={peer: entries}, zhalo=None) from every rank; returns (y_loc, z) where z is this rank's
    partial result in the [own | halo] layout.  Call twice: first to get z of every rank, then
    reduce_scatter_emulated to finish."""
    rowptr, _, val, _ = csr_rows
    world = len(gathered)
    halo = [gathered[q]["sendbuf"][rank] for q in range(world) if plan["recv_counts"][q]]
    halo = np.concatenate(halo) if halo else np.zeros(0, dtype=val.dtype)
    xl = gathered[rank]["x"]
    n_pad = plan["n_pad"]
    xext = np.concatenate([xl, np.zeros(n_pad - len(xl), dtype=xl.dtype), halo])
    nrows = len(rowptr) - 1
    rows = np.repeat(np.arange(nrows), np.diff(rowptr))
    y = np.zeros(nrows, dtype=np.result_type(val.dtype, xl.dtype))
    np.add.at(y, rows, val * xext[plan["col_local"]])                       # y_loc = T_loc [x_own | halo]
    z = np.zeros(n_pad + plan["nhalo"], dtype=y.dtype)
    np.add.at(z, plan["col_local"], np.conj(val) * y[rows])                 # z = T_loc^H y_loc
    return y, z


def reduce_scatter_emulated(plan, zs, rank, plans):
    """rank-ordered sum of the peers' halo segments that belong to this rank's rows (what
    k_reduce_scatter does): zs[q] = rank q's partial result, plans[q] = its plan."""
    world = len(zs)
    r0, r1 = plan["row_range"]
    n_pad = plan["n_pad"]
    out = zs[rank][: r1 - r0].copy()
    off = np.concatenate([[0], np.cumsum(plan["send_counts"])])
    for q in range(world):
        if q == rank or plan["send_counts"][q] == 0:
            continue
        rows = plan["send_idx"][off[q]:off[q + 1]]                          # my rows that rank q references
        seg0 = n_pad + int(plan["send_dest_off"][q])                        # where they start in q's halo part
        out[rows] += zs[q][seg0: seg0 + len(rows)]
    return out
