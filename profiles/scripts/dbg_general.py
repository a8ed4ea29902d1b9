"""First GPU run of the generalised problem (mode 2, bmat = :G): every check prints one JSON line."""
import importlib, json, sys, time, traceback
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sl
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
O = entry.load_oracle()
GOLD = [121003.49732902342, 121616.60247324049, 122057.49457079473, 122323.22366457577]


def dsdrv3(n):
    A = (sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * (n + 1)).tocsr()
    B = (sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]) * (1.0 / (6 * (n + 1)))).tocsr()
    return A, B


def run(name, fn):
    t0 = time.time()
    try:
        out = fn()
    except Exception as e:
        out = {"err": str(e)[:300], "tb": traceback.format_exc()[-600:]}
    out["check"] = name
    out["sec"] = round(time.time() - t0, 2)
    print(json.dumps(out), flush=True)


def c_ops():
    A, B = dsdrv3(100)
    ctx = pkg.Context(pkg.ArpackSymmetricGeneralizedOp(A, B), 20)
    x = np.random.default_rng(0).standard_normal(100)
    y = np.asarray(ctx.opx(x)); ref = sl.solve(B.toarray(), A @ x)
    bx = np.asarray(ctx.bx(x))
    out = {"opx_relerr": float(np.linalg.norm(y - ref) / np.linalg.norm(ref)), "bx_relerr": float(np.linalg.norm(bx - B @ x) / np.linalg.norm(B @ x)),
           "solve_stats": ctx.solve_stats()}
    ctx.close()
    return out


def c_lockstep():
    A, B = dsdrv3(100)
    ncv = 12
    ctx = pkg.Context(pkg.ArpackSymmetricGeneralizedOp(A, B), ncv)
    o = O.Oracle(A, ncv, B=B)
    rc, rn = ctx.getv0()
    o.dgetv0(1, False, 1)
    r_gpu = ctx.download_resid()
    out = {"getv0_rc": rc, "rnorm_gpu": float(rn), "rnorm_ora": float(o.rnorm), "resid_diff": float(np.max(np.abs(r_gpu - o.resid)))}
    H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, ncv, H, rn)
    o.dsaitr(0, ncv)
    V = ctx.download_v(0, ncv)
    out.update(info=info, H_diff=float(np.max(np.abs(H - np.asarray(o.H)))), H_scale=float(np.max(np.abs(H))),
               V_diff=float(np.max(np.abs(V - np.swapaxes(o.Vt, 0, 1)))), rnorm_diff=float(abs(rn - o.rnorm)),
               VBV=float(np.max(np.abs(V.T @ (B @ V) - np.eye(ncv)))), stats=dict(ctx.stats()), ora_stats={k: v for k, v in o.stats().items() if k.startswith("n")})
    ctx.close()
    return out


def c_eigs(which="LM", ncv=20, maxiter=300):
    A, B = dsdrv3(100)
    st = {}
    res = pkg.eigs(A, B, 4, ncv=ncv, which=which, maxiter=maxiter, stats=st)
    w = sl.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    ref = np.sort(w)[-4:] if which == "LM" else np.sort(w)[:4]
    ora = O.Oracle(A, ncv, B=B).symeigs(4, which=which, maxiter=maxiter)
    Z = res.vectors
    return {"values": [float(v) for v in res.values], "rel_vs_dense": float(np.max(np.abs(np.sort(res.values) - ref) / np.abs(ref))),
            "rel_vs_golden": float(np.max(np.abs(np.array(res.values) - GOLD) / np.array(GOLD))) if which == "LM" else None,
            "restarts": int(res.iparam[2]), "restarts_oracle": ora["niter"], "mode": int(res.iparam[6]),
            "nopx": st.get("nopx"), "nbx": st.get("nbx"), "ora_nopx": ora["stats"]["nopx"], "ora_nbx": ora["stats"]["nbx"],
            "ZBZ": float(np.max(np.abs(Z.T @ (B @ Z) - np.eye(4))))}


def c_dd():
    A, B = dsdrv3(100)
    res = pkg.eigs(A, B, 4, ncv=20, dtype="dd")
    ora = O.Oracle(A, 20, dtype=O.DD, B=B).symeigs(4)
    v = np.asarray(res.values); ov = np.asarray(ora["values"])
    d = (v[:, 0] - ov[:, 0]) + (v[:, 1] - ov[:, 1])
    return {"values_hi": [float(x) for x in v[:, 0]], "rel_vs_oracle": float(np.max(np.abs(d) / np.abs(ov[:, 0]))),
            "restarts": int(res.iparam[2]), "restarts_oracle": ora["niter"]}


def c_complex():
    rng = np.random.default_rng(3)
    n = 60
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, n - 1))
    A = sp.diags([np.conj(ph), rng.uniform(1, 3, n), ph], [-1, 0, 1]).tocsr()
    pb = 0.3 * np.exp(1j * rng.uniform(0, 2 * np.pi, n - 1))
    B = sp.diags([np.conj(pb), rng.uniform(1.5, 2.5, n), pb], [-1, 0, 1]).tocsr()
    w = sl.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    res = pkg.hermeigs(A, B, 4, ncv=16)
    ref = np.sort(w[np.argsort(-np.abs(w))[:4]])
    Z = res.vectors
    return {"rel_vs_dense": float(np.max(np.abs(np.sort(res.values) - ref) / np.abs(ref))), "restarts": int(res.iparam[2]),
            "ZBZ": float(np.max(np.abs(Z.conj().T @ (B @ Z) - np.eye(4))))}


def c_large():
    m = 300
    A = syn.to_scipy(syn.laplacian2d(m))
    n = m * m
    B = (sp.identity(n) * 2.0 + 0.25 * A).tocsr()     # SPD "mass-like" matrix with the stencil's pattern
    st = {}
    t0 = time.time()
    res = pkg.eigs(A, B, 6, ncv=24, maxiter=2000, stats=st, keep_context=True)
    t = time.time() - t0
    Z = res.vectors
    rr = [float(np.linalg.norm(A @ Z[:, i] - res.values[i] * (B @ Z[:, i])) / np.linalg.norm(A @ Z[:, i])) for i in range(6)]
    # analytic: A and B share eigenvectors: lambda = mu / (2 + mu/4), mu in (0, 8) -> largest ~ 8 / 4 = 2
    out = {"n": n, "values": [float(v) for v in res.values], "max_rel_residual": max(rr), "restarts": int(res.iparam[2]),
           "ZBZ": float(np.max(np.abs(Z.T @ (B @ Z) - np.eye(6)))), "solve_s": round(t, 3), "nopx": st.get("nopx"), "nbx": st.get("nbx"),
           "solve_stats": res.ctx.solve_stats()}
    res.ctx.close()
    return out


run("ops", c_ops)
run("lockstep", c_lockstep)
run("eigs_LM", lambda: c_eigs("LM", 20, 300))
run("eigs_LM_ncv10", lambda: c_eigs("LM", 10, 300))
run("eigs_SA_ncv10", lambda: c_eigs("SA", 10, 400))
run("dd", c_dd)
run("complex", c_complex)
run("large", c_large)
