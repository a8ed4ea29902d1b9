"""Generalised problem A x = lambda B x (mode 2, device CG with B) at n = 90 000: whole-solve time and ms per Lanczos step.
A = 2-D Laplacian, B = 2 I + A / 4 (SPD, same pattern), nev = 6, ncv = 24.  usage: python profiles/microbench/general_bench.py"""
import importlib, json, os, sys, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
m = 300
A = syn.to_scipy(syn.laplacian2d(m)); n = m * m
B = (sp.identity(n) * 2.0 + 0.25 * A).tocsr()
out = {"n": n, "env": {k: os.environ[k] for k in ("GA_CG_GRAPH",) if k in os.environ}}
for rep in range(2):      # second solve: warm (kernels loaded, graph built per context again)
    st = {}
    t0 = time.time()
    res = pkg.eigs(A, B, 6, ncv=24, maxiter=2000, stats=st, keep_context=True)
    t = time.time() - t0
    Z = res.vectors
    rr = max(float(np.linalg.norm(A @ Z[:, i] - res.values[i] * (B @ Z[:, i])) / np.linalg.norm(A @ Z[:, i])) for i in range(6))
    out["solve_%d" % rep] = {"seconds": round(t, 3), "restarts": int(res.iparam[2]), "nopx": st.get("nopx"), "ms_per_lanczos_step": round(t * 1e3 / st.get("nopx"), 4),
                             "max_rel_residual": rr, "solve_stats": [float(v) for v in res.ctx.solve_stats()], "values": [float(v) for v in res.values]}
    res.ctx.close()
print(json.dumps(out))
