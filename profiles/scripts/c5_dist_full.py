"""C5 at its stated size sharded over N GPUs (torchrun, one process per GPU): svds of the synthetic 50M x 2M matrix (nnz = 5e8) via
the row-sharded normal operator, k = 20, ncv = 60.  Every rank generates only its rows; checks are the size-independent ones
(ranks agree bitwise, V'V = I, U'U = I through an all-reduce, ||A V - U S|| row-block by row-block) plus the single-GPU
singular values when a file with them is given.  usage: torchrun ... c5_dist_full.py [scale] [out.json]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    m, n, k, ncv = int(50_000_000 * scale), int(2_000_000 * scale), 20, 60
    tblk, tparts = gd.block_partition(m, world)
    t0, t1 = tparts[rank]
    tg = time.time()
    rowptr, col, val, shp = syn.tall_sparse_stream(m, n, 10, row_range=(t0, t1))
    plan = gd.halo_plan(col, n, rank, world)
    r0, r1 = plan["row_range"]
    tgen = time.time() - tg
    tu = time.time()
    ctx = pkg.Context(pkg.ArpackNormalOp((rowptr, col, val, (t1 - t0, n))), ncv, device=local, rank=rank, world_size=world,
                      n_global=n, row_offset=r0, plan=plan)
    gd.setup_comm(ctx, pkg.comm_unique_id, rank, world, fused=True)
    tup = time.time() - tu
    ok = {}
    torch.cuda.synchronize(); dist.barrier()
    ts = time.time()
    rc = ctx.symeigs_begin(k, "LM", 0.0, 300, None)
    rcit = ctx.symeigs_iterate(-1)
    info, values, bounds, vecs, iparam = ctx.symeigs_end(k, ritzvec=True)
    torch.cuda.synchronize(); dist.barrier()
    tsolve = time.time() - ts
    st = ctx.stats()
    ok["solve_rc"] = bool(rc == 0 and rcit == 99 and info == 0)
    S = np.sqrt(np.maximum(values, 0.0))
    order = np.argsort(-S, kind="stable")
    S = S[order]
    tv = time.time()
    rcv, W, norms = ctx.svd_vectors(order)              # this rank's rows of U
    tvec = time.time() - tv
    ok["svd_vectors_rc"] = bool(rcv == 0)
    agree = [None] * world
    dist.all_gather_object(agree, ([float(s) for s in S], int(iparam[2])))
    ok["ranks_agree_bitwise"] = bool(all(a == agree[0] for a in agree))
    Vs = [None] * world
    dist.all_gather_object(Vs, np.ascontiguousarray(vecs[:, order]))
    V = np.concatenate(Vs, axis=0)
    ok["V_orthonormal"] = bool(np.max(np.abs(V.T @ V - np.eye(k))) <= 1e-12)
    G = torch.from_numpy(np.ascontiguousarray(W.T @ W)).cuda()
    dist.all_reduce(G)
    ok["U_orthonormal"] = bool(np.max(np.abs(G.cpu().numpy() - np.eye(k))) <= 1e-12)
    import scipy.sparse as sp
    Tl = sp.csr_matrix((val, col, rowptr.astype(np.int32)), shape=(t1 - t0, n))
    res = torch.tensor([float(np.max(np.linalg.norm(Tl @ V[:, [0, k // 2, k - 1]] - W[:, [0, k // 2, k - 1]] * S[[0, k // 2, k - 1]], axis=0)))], device="cuda", dtype=torch.float64)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    ok["svd_residuals_le_1e-11_sigma1"] = bool(float(res.item()) <= 1e-11 * S[0])
    allok = [None] * world
    dist.all_gather_object(allok, all(ok.values()))
    if rank == 0:
        out = {"world": world, "m": m, "n": n, "nnz": int(10 * m), "k": k, "ncv": ncv, "checks": ok, "all_ranks_ok": bool(all(allok)),
               "restarts": int(iparam[2]), "lanczos_steps": int(st["nopx"]), "solve_s": round(tsolve, 3),
               "steps_per_s": round(st["nopx"] / max(st["taitr"], 1e-9), 1), "taitr_s": round(st["taitr"], 3), "svd_vectors_s": round(tvec, 2),
               "gen_s": round(tgen, 1), "upload_and_setup_s": round(tup, 1), "halo_entries": int(plan["nhalo"]),
               "singular_values": [float(s) for s in S]}
        print(json.dumps(out))
        if len(sys.argv) > 2:
            json.dump(out, open(sys.argv[2], "w"), indent=1)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if all(allok) else 1)


if __name__ == "__main__":
    main()
