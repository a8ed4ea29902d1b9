#!/usr/bin/env python
"""Time to solution of BASELINE config C2 on one B200: eigs(A, 10; ncv=40) Float64 on the 2-D Laplacian with
m x m grid points (default m = 4000, n = 16M), whole solve through the library's own IRL driver, then checks
against the analytic spectrum and true Ritz residuals.  Progress is written after every chunk of restarts so a
cut-off run still leaves evidence.   python profiles/scripts/c2_tts.py [m] [budget_seconds]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 200.0
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
out_path = os.path.join(ROOT, "gpurun_out", "c2_tts_m%d.json" % m)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
nev, ncv = 10, 40
t0 = time.perf_counter()
csr = syn.laplacian2d(m)
t_gen = time.perf_counter() - t0
n = m * m
t0 = time.perf_counter()
ctx = pkg.Context(pkg.ArpackSimpleOp(csr), ncv)
t_up = time.perf_counter() - t0
maxiter = max(300, int(np.ceil(np.sqrt(n))))
res = {"workload": "C2 eigs(A,10;ncv=40) Float64 2-D Laplacian n=%d, 1 GPU, whole solve" % n, "maxiter": maxiter,
       "generate_s": round(t_gen, 2), "context_and_upload_s": round(t_up, 3)}
t0 = time.perf_counter()
rc = ctx.symeigs_begin(nev, "LM", 0.0, maxiter, None)
assert rc == 0, rc
done = 0
while True:
    rcit = ctx.symeigs_iterate(100)
    done += 100
    el = time.perf_counter() - t0
    st = ctx.stats()
    res.update(progress={"restarts_upper": done, "elapsed_s": round(el, 2), "lanczos_steps": int(st["nopx"])})
    json.dump(res, open(out_path, "w"), indent=1)
    if rcit == 99 or el > budget:
        break
info, values, bounds, vecs, iparam = ctx.symeigs_end(nev, ritzvec=True)
t_solve = time.perf_counter() - t0
st = ctx.stats()
exact = syn.laplacian2d_eigs(m, nev)
res.update(finished=bool(rcit == 99), info=int(info), solve_s=round(t_solve, 2), restarts=int(iparam[2]), nconv=int(iparam[4]),
           lanczos_steps=int(st["nopx"]), nrorth=int(st["nrorth"]), nitref=int(st["nitref"]),
           steps_per_s=round(st["nopx"] / max(st["taitr"], 1e-9), 1),
           stats_s={k: round(float(st[k]), 3) for k in ("taitr", "tapps", "teigt", "tgets", "tconv", "trvec")},
           values=[float(v) for v in values])
if len(values) == nev:
    res["max_rel_err_vs_analytic"] = float(np.max(np.abs(values - exact) / exact))
    A = syn.to_scipy(csr)
    R = A @ vecs - vecs * values
    res["max_ritz_residual_rel"] = float(np.max(np.linalg.norm(R, axis=0) / np.abs(values)))
    res["ZtZ_minus_I"] = float(np.max(np.abs(vecs.T @ vecs - np.eye(nev))))
    res["max_bound_over_lambda"] = float(np.max(bounds / np.abs(values)))
ctx.close()
json.dump(res, open(out_path, "w"), indent=1)
print(json.dumps(res))
