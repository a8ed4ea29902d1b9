"""Multi-GPU parity check of the row-sharded normal operator / svds (torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29513 tests/dist_gpu_check_normal.py [m n]

svds(A, 6; ncv=24) of a tall sparse A (C5 shape, reduced) with rows of A and of the Lanczos vectors
block-partitioned: OP*x = A'(A x) against scipy, singular values against a dense SVD and a single-GPU
run, U/V orthonormality and residuals (check_svd, reference test/utility.jl:244-253).
GA_NCCL_ONLY=1 selects ncclSend/ncclRecv instead of the peer-memory path.  Exit code 0 = parity green.
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

    m = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    A = syn.to_scipy(syn.tall_sparse(m, n, 8))
    k, ncv = 6, 24
    rows, plan, nn, (t0, t1) = gd.shard_normal(A, rank, world)
    r0, r1 = plan["row_range"]
    op = pkg.ArpackNormalOp(rows)
    ctx = pkg.Context(op, ncv, device=local, rank=rank, world_size=world, n_global=nn, row_offset=r0, plan=plan)
    fused = os.environ.get("GA_NCCL_ONLY", "0") != "1"
    gd.setup_comm(ctx, pkg.comm_unique_id, rank, world, fused=fused)
    ok = {}
    # 1. OP*x
    x = np.cos(np.arange(nn) * 0.013) + 0.2
    y = ctx.opx(x[r0:r1])
    ref = A.T @ (A @ x)
    ok["opx"] = bool(np.allclose(y, ref[r0:r1], rtol=1e-12, atol=1e-12))
    y2 = ctx.opx(x[r0:r1])
    ok["opx_repeatable_bitwise"] = bool(np.array_equal(y, y2))
    # 2. svds: Lanczos on the sharded operator + device post-processing
    rc = ctx.symeigs_begin(k, "LM", 0.0, 300, None)
    rcit = ctx.symeigs_iterate(-1)
    info, values, bounds, vecs, iparam = ctx.symeigs_end(k, ritzvec=True)
    ok["solve_rc"] = bool(rc == 0 and rcit == 99 and info == 0)
    S = np.sqrt(np.maximum(values, 0.0))
    order = np.argsort(-S, kind="stable")
    S = S[order]
    rcv, W, norms = ctx.svd_vectors(order)
    ok["svd_vectors_rc"] = bool(rcv == 0)
    Vs = [None] * world; Us = [None] * world
    dist.all_gather_object(Vs, vecs[:, order])
    dist.all_gather_object(Us, W)
    agree = [None] * world
    dist.all_gather_object(agree, ([float(s) for s in S], int(iparam[2])))
    ok["ranks_agree_bitwise"] = bool(all(a == agree[0] for a in agree))
    out = None
    if rank == 0:
        V = np.concatenate(Vs, axis=0); U = np.concatenate(Us, axis=0)
        sd = np.linalg.svd(A.toarray(), compute_uv=False)[:k]
        ok["singular_values_vs_dense"] = bool(np.allclose(S, sd, rtol=1e-10))
        ok["V_orthonormal"] = bool(np.allclose(V.T @ V, np.eye(k), atol=1e-12))
        ok["U_orthonormal"] = bool(np.allclose(U.T @ U, np.eye(k), atol=1e-12))
        ok["svd_residuals"] = bool(np.max(np.linalg.norm(A @ V - U * S, axis=0) / S) <= 1e-12)
        single = pkg.svds(A, k, ncv=ncv, device=local)
        ok["vs_single_gpu_values"] = bool(np.allclose(single.S, S, rtol=1e-12))
        sg = np.sign(np.sum(single.V * V, axis=0))
        ok["vs_single_gpu_vectors"] = bool(np.allclose(single.V * sg, V, atol=1e-8) and np.allclose(single.U * sg, U, atol=1e-8))
        out = {"world": world, "m": m, "n": nn, "fused_peer_memory": fused, "checks": ok, "restarts": int(iparam[2]),
               "nopx": int(ctx.stats()["nopx"]), "halo": int(plan["nhalo"]), "singular_values": [float(s) for s in S]}
        print(json.dumps(out))
    allok = [None] * world
    dist.all_gather_object(allok, all(ok.values()))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if all(allok) else 1)


if __name__ == "__main__":
    main()
