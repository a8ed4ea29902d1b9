"""Multi-GPU parity check, launched with torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py [m]

Row-sharded eigs(A, 10; ncv=40) on the 2-D Laplacian (m x m grid) against the analytic spectrum,
the CPU oracle (start vector bit-exact across the shards, OP*x with halo exchange) and a
single-GPU run of the same problem.  Prints one JSON line on rank 0; exit code 0 = parity green.
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    m = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    O = entry.load_oracle()
    n, nev, ncv = m * m, 10, 40
    blk, parts = gd.block_partition(n, world)
    r0, r1 = parts[rank]
    csr = syn.laplacian2d(m, row_range=(r0, r1))
    op = pkg.ArpackSimpleOp((csr[0], csr[1], csr[2], (r1 - r0, n)))
    plan = gd.halo_plan(csr[1], n, rank, world)
    ctx = pkg.Context(op, ncv, device=local, rank=rank, world_size=world, n_global=n, row_offset=r0, plan=plan)
    fused = os.environ.get("GA_NCCL_ONLY", "0") != "1"
    gd.setup_comm(ctx, pkg.comm_unique_id, rank, world, fused=fused)
    ok = {}
    # 1. start vector: every shard is its slice of the reference's dlarnv stream, bit for bit
    ierr, rn = ctx.getv0()
    ref, seed = O.dlarnv(n)
    ok["start_vector_bit_exact"] = bool(np.array_equal(ctx.download_resid(), ref[r0:r1]) and tuple(ctx.iseed) == seed)
    ok["start_norm"] = bool(abs(rn - O.norm2(ref)) <= 2 * np.spacing(rn))
    # 2. OP*x with halo exchange
    A = syn.to_scipy(syn.laplacian2d(m))
    y = ctx.opx(ref[r0:r1])
    ok["opx"] = bool(np.allclose(y, (A @ ref)[r0:r1], rtol=1e-13, atol=1e-13))
    # 3. whole solve
    rc = ctx.symeigs_begin(nev, "LM", 0.0, max(300, int(np.ceil(np.sqrt(n)))), None)
    rcit = ctx.symeigs_iterate(-1)
    info, values, bounds, vecs, iparam = ctx.symeigs_end(nev, ritzvec=True)
    exact = syn.laplacian2d_eigs(m, nev)
    ok["solve_rc"] = bool(rc == 0 and rcit == 99 and info == 0)
    ok["eigenvalues_1e-10"] = bool(len(values) == nev and np.max(np.abs(values - exact) / exact) <= 1e-10)
    # Ritz vectors: gather the shards and check residuals on rank 0
    shards = [None] * world
    dist.all_gather_object(shards, vecs)
    restarts = [None] * world
    dist.all_gather_object(restarts, (int(iparam[2]), [float(v) for v in values]))
    ok["ranks_agree_bitwise"] = bool(all(r == restarts[0] for r in restarts))
    out = None
    if rank == 0:
        Z = np.concatenate(shards, axis=0)
        R = A @ Z - Z * values
        ok["ritz_residuals"] = bool(np.max(np.linalg.norm(R, axis=0) / np.abs(values)) <= 1e-10)
        ok["ritz_orthonormal"] = bool(np.allclose(Z.T @ Z, np.eye(nev), atol=1e-12))
        single = pkg.eigs(pkg.ArpackSimpleOp(syn.laplacian2d(m)), nev, ncv=ncv, device=local)
        ok["vs_single_gpu_values"] = bool(np.max(np.abs(single.values - values) / np.abs(values)) <= 1e-12)
        ok["vs_single_gpu_restarts"] = bool(abs(int(single.iparam[2]) - int(iparam[2])) <= 1)   # north_star: +-1 (Dot2 sweeps: independent of the row partition)
        out = {"world": world, "n": n, "fused_peer_memory": fused, "checks": ok, "restarts": int(iparam[2]), "restarts_single_gpu": int(single.iparam[2]),
               "nopx": int(ctx.stats()["nopx"]), "halo": int(plan["nhalo"])}
        print(json.dumps(out))
    allok = [None] * world
    dist.all_gather_object(allok, all(ok.values()))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if all(allok) else 1)


if __name__ == "__main__":
    main()
