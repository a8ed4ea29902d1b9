"""Bit-level comparison of the B200 path with the oracle in its exact-arithmetic mode (development probe).

    python profiles/scripts/parity_probe.py [out.json]

Lock step: dgetv0 + ncv Lanczos steps on 2-D Laplacians, H / V / resid compared bit for bit; then whole
solves: restart counts, Lanczos step counts and eigenvalues of the device vs the oracle."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
O = entry.load_oracle()
out = {"lib": os.environ.get("GA_B200_LIB", "default")}


def ulps(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64).astype(np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64).astype(np.float64)
    return float(np.max(np.abs(a - b))) if a.size else 0.0


lock = []
for m, ncv in ((48, 24), (300, 40)):
    A = syn.to_scipy(syn.laplacian2d(m))
    for mode in ("exact", "plain"):
        with O.arith(mode):
            ctx = pkg.Context(pkg.ArpackSimpleOp(A), ncv)
            ora = O.Oracle(A, ncv)
            ctx.getv0(); ora.dgetv0()
            rn = ctx.norm2_resid()
            H = np.zeros((ncv, 2))
            info, rn = ctx.lanczos(0, ncv, H, rn)
            ora.dsaitr(0, ncv)
            V = np.asarray(ctx.download_v()); Vo = np.swapaxes(np.asarray(ora.Vt), 0, 1)
            r = np.asarray(ctx.download_resid()); ro = np.asarray(ora.resid)
            Ho = np.asarray(ora.H)
            ncols_equal = int(sum(np.array_equal(V[:, c], Vo[:, c]) for c in range(ncv)))
            lock.append({"m": m, "ncv": ncv, "oracle_arith": mode, "H_bitwise": bool(np.array_equal(H, Ho)),
                         "H_max_ulps": ulps(H, Ho), "V_cols_bitwise": ncols_equal, "V_maxabs": float(np.max(np.abs(V - Vo))),
                         "resid_bitwise": bool(np.array_equal(r, ro)), "rnorm_equal": bool(rn == ora.rnorm),
                         "nrorth": (ctx.stats()["nrorth"], ora.stats()["nrorth"])})
            ctx.close()
out["lockstep"] = lock
print(json.dumps(lock), flush=True)

solves = []
for m, nev, ncv in ((64, 4, 20), (100, 4, 20), (100, 10, 40), (300, 10, 40)):
    csr = syn.laplacian2d(m)
    A = syn.to_scipy(csr)
    t0 = time.time()
    res = pkg.eigs(pkg.ArpackSimpleOp(csr), nev, ncv=ncv)
    tg = time.time() - t0
    ex = syn.laplacian2d_eigs(m, nev)
    row = {"m": m, "nev": nev, "ncv": ncv, "gpu_restarts": int(res.iparam[2]), "gpu_nopx": int(res.iparam[8]),
           "gpu_err_vs_analytic": float(np.max(np.abs(res.values - ex) / ex)), "gpu_s": round(tg, 3)}
    for mode in ("exact", "plain"):
        with O.arith(mode):
            t0 = time.time()
            r = O.Oracle(A, ncv).symeigs(nev)
            row["oracle_%s_restarts" % mode] = int(r["niter"])
            row["oracle_%s_values_bitwise" % mode] = bool(np.array_equal(np.asarray(r["values"]), np.asarray(res.values)))
            row["oracle_%s_maxrel" % mode] = float(np.max(np.abs(np.asarray(r["values"]) - res.values) / np.abs(res.values)))
            row["oracle_%s_s" % mode] = round(time.time() - t0, 2)
    solves.append(row)
    print(json.dumps(row), flush=True)
out["solves"] = solves
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
