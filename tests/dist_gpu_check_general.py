"""Multi-GPU parity check of the row-sharded generalised problem A x = lambda B x (mode 2, bmat = :G;
ArpackSymmetricGeneralizedOp, reference src/idonow_ops.jl:462-492) under torchrun, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tests/dist_gpu_check_general.py [m]

  - the operator pair on the reference's dsdrv3 pencil (test/interface_high.jl:31-45): OP x = inv(B) (A x) by the sharded
    conjugate gradients against a dense solve, B x against the CSR product;
  - dgetv0! (bmat = :G) + 12 dsaitr! steps in lock step with a single-GPU run of the same problem: H, rnorm and every
    counter (nopx, nbx, nrorth, nitref);
  - a whole solve on the pencil (2-D Laplacian, 2 I + A / 4) with an analytic spectrum: eigenvalues, B-orthonormal Ritz
    vectors, true residuals, all ranks bit-identical, and the single-GPU run's values / restart count.
Exit code 0 = parity green.
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
    import scipy.linalg as sl
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist

    from conftest import DSDRV3_GOLDEN, dsdrv3

    m = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    gd = importlib.import_module(entry.PKG_NAME + ".dist")
    ok, info_out = {}, {}

    def sharded_ctx(A, B, ncv):
        ra, rb, plan, (r0, r1) = gd.shard_generalized(A, B, rank, world)
        op = pkg.ArpackSymmetricGeneralizedOp(ra, rb, sharded=True)
        ctx = pkg.Context(op, ncv, device=local, rank=rank, world_size=world, n_global=A.shape[0], row_offset=r0, plan=plan)
        gd.setup_comm(ctx, pkg.comm_unique_id, rank, world, fused=True)
        return ctx, plan, r0, r1

    # ---- 1. operator pair + lock step on dsdrv3 (n = 100 per the reference's test; 300 here so that every rank has rows)
    n1, ncv = 300, 12
    A, B = dsdrv3(n1)
    ctx, plan, r0, r1 = sharded_ctx(A, B, ncv)
    rng = np.random.default_rng(0)
    good_op, good_bx = True, True
    for _ in range(3):
        x = rng.standard_normal(n1)
        y = np.asarray(ctx.opx(x[r0:r1]))
        ref = sl.solve(B.toarray(), A @ x)[r0:r1]
        good_op = good_op and np.linalg.norm(y - ref) <= 1e-12 * np.linalg.norm(ref)
        bx = np.asarray(ctx.bx(x[r0:r1]))
        good_bx = good_bx and np.linalg.norm(bx - (B @ x)[r0:r1]) <= 1e-15 * np.linalg.norm((B @ x)[r0:r1])
    ok["opx_vs_dense_solve"] = bool(good_op)
    ok["bx"] = bool(good_bx)
    nsolve, niter, relres = ctx.solve_stats()
    ok["cg_stats"] = bool(nsolve == 3 and niter > 0 and relres <= 1e-14)
    ctx.reset_stats()
    rc, rn = ctx.getv0()
    H = np.zeros((ncv, 2))
    info, rn2 = ctx.lanczos(0, ncv, H, rn)
    st = ctx.stats()
    Vs = [None] * world
    dist.all_gather_object(Vs, np.asarray(ctx.download_v(0, ncv)))
    ctx.close()
    single = pkg.Context(pkg.ArpackSymmetricGeneralizedOp(A, B), ncv, device=local)
    rc1, rn1 = single.getv0()
    H1 = np.zeros((ncv, 2))
    info1, rn21 = single.lanczos(0, ncv, H1, rn1)
    st1 = single.stats()
    single.close()
    V = np.concatenate(Vs, axis=0)
    ok["lockstep_rc"] = bool(rc == 0 and info == 0 and rc1 == 0 and info1 == 0)
    ok["lockstep_start_norm"] = bool(abs(rn - rn1) <= 1e-12 * rn1)
    ok["lockstep_H_vs_single_gpu"] = bool(np.max(np.abs(H - H1)) <= 1e-11 * np.max(np.abs(H1)))
    ok["lockstep_rnorm"] = bool(abs(rn2 - rn21) <= 1e-9 * max(1.0, np.max(np.abs(H1))))
    ok["lockstep_V_B_orthonormal"] = bool(np.max(np.abs(V.T @ (B @ V) - np.eye(ncv))) <= 1e-13)
    ok["lockstep_counters"] = bool(all(st[k] == st1[k] for k in ("nopx", "nbx", "nrorth", "nitref", "nrstrt")))
    info_out["counters"] = {k: int(st[k]) for k in ("nopx", "nbx", "nrorth", "nitref", "nrstrt")}

    # ---- 2. dsdrv3 golden eigenvalues (n = 100: test/interface_high.jl:45) on the sharded path
    A, B = dsdrv3(100)
    ctx, plan, r0, r1 = sharded_ctx(A, B, 20)
    rc = ctx.symeigs_begin(4, "LM", 0.0, 300, None)
    rcit = ctx.symeigs_iterate(-1)
    info, values, bounds, vecs, iparam = ctx.symeigs_end(4, ritzvec=True)
    ok["dsdrv3_rc_mode2"] = bool(rc == 0 and rcit == 99 and info == 0 and int(iparam[6]) == 2)
    ok["dsdrv3_golden_1e-12"] = bool(np.allclose(values, DSDRV3_GOLDEN, rtol=1e-12, atol=0))
    info_out["dsdrv3_restarts"] = int(iparam[2])
    ctx.close()

    # ---- 3. whole solve with an analytic spectrum: A = 2-D Laplacian, B = 2 I + A / 4 (same eigenvectors)
    A = syn.to_scipy(syn.laplacian2d(m))
    n = m * m
    B = (sp.identity(n) * 2.0 + 0.25 * A).tocsr()
    nev, ncv = 6, 24
    ctx, plan, r0, r1 = sharded_ctx(A, B, ncv)
    rc = ctx.symeigs_begin(nev, "LM", 0.0, 1000, None)
    rcit = ctx.symeigs_iterate(-1)
    info, values, bounds, vecs, iparam = ctx.symeigs_end(nev, ritzvec=True)
    k = np.arange(1, m + 1)
    mu1 = 2 - 2 * np.cos(k * np.pi / (m + 1))
    mu = np.sort((mu1[:, None] + mu1[None, :]).ravel())[-nev:]
    exact = mu / (2 + mu / 4)
    ok["solve_rc"] = bool(rc == 0 and rcit == 99 and info == 0)
    ok["eigenvalues_1e-10"] = bool(len(values) == nev and np.allclose(np.sort(values), exact, rtol=1e-10, atol=0))
    shards = [None] * world
    dist.all_gather_object(shards, np.asarray(vecs))
    agree = [None] * world
    dist.all_gather_object(agree, (int(iparam[2]), [float(v) for v in values]))
    ok["ranks_agree_bitwise"] = bool(all(a == agree[0] for a in agree))
    nsolve, niter, relres = ctx.solve_stats()
    st = ctx.stats()
    if rank == 0:
        Z = np.concatenate(shards, axis=0)
        ok["ritz_B_orthonormal"] = bool(np.max(np.abs(Z.T @ (B @ Z) - np.eye(nev))) <= 1e-10)
        ok["ritz_residuals"] = bool(all(np.linalg.norm(A @ Z[:, i] - values[i] * (B @ Z[:, i])) <= 1e-9 * np.linalg.norm(A @ Z[:, i])
                                        for i in range(nev)))
        sg = pkg.eigs(A, B, nev, ncv=ncv, maxiter=1000, device=local)
        ok["vs_single_gpu_values"] = bool(np.max(np.abs(np.sort(sg.values) - np.sort(values)) / np.abs(np.sort(values))) <= 1e-12)
        ok["vs_single_gpu_restarts"] = bool(abs(int(sg.iparam[2]) - int(iparam[2])) <= 1)
        info_out.update(world=world, n=n, restarts=int(iparam[2]), restarts_single_gpu=int(sg.iparam[2]), nopx=int(st["nopx"]),
                        nbx=int(st["nbx"]), cg_solves=int(nsolve), cg_iterations=int(niter), cg_max_relres=float(relres),
                        halo=int(plan["nhalo"]))
        print(json.dumps({"checks": ok, **info_out}))
    allok = [None] * world
    dist.all_gather_object(allok, all(ok.values()))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if all(allok) else 1)


if __name__ == "__main__":
    main()
