"""C5 at full size (50M x 2M, nnz = 5e8) on one GPU: time of the normal-operator product y = A'(A x) with and without the
column-blocked second product, and the Lanczos step rate of svds' iteration (k = 20, ncv = 60).
usage: python profiles/microbench/c5_normal_bench.py [scale]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
m, n, ncv, k = int(50_000_000 * scale), int(2_000_000 * scale), 60, 20
t0 = time.time(); csr = syn.tall_sparse_stream(m, n, 10); tgen = time.time() - t0
nnz = int(csr[0][-1])
alg_bytes = 2 * (nnz * 12 + (m + 1) * 8) + 8 * (2 * n + 2 * m)          # normal-operator apply
out = {"m": m, "n": n, "nnz": nnz, "gen_s": round(tgen, 1), "algorithmic_GB_per_apply": round(alg_bytes / 1e9, 3)}
sweep = [("unblocked", str(1 << 40)), ("blocked_32MiB", None)]
if len(sys.argv) > 2 and sys.argv[2] == "sweep":      # block width in elements of the long vector
    sweep = [("blocked_%dMiB" % (w * 8 >> 20), str(w)) for w in (1 << 20, 1 << 21, 1 << 22, 1 << 23)]
if len(sys.argv) > 2 and sys.argv[2] == "blocked":
    sweep = [("blocked_32MiB", None)]
for label, env in sweep:
    if env is None:
        os.environ.pop("GA_NORMAL_COLBLOCK", None)
    else:
        os.environ["GA_NORMAL_COLBLOCK"] = env
    t0 = time.time()
    ctx = pkg.Context(pkg.ArpackNormalOp(csr), ncv)
    tup = time.time() - t0
    ms = ctx.bench_opx(10)
    ctx.getv0(); rn = ctx.norm2_resid()
    H = np.zeros((ncv, 2))
    ctx.timer_start()
    info, rn = ctx.lanczos(0, ncv, H, rn)
    tl = ctx.timer_stop()
    out[label] = {"setup_s": round(tup, 1), "opx_ms": round(ms, 3), "opx_GBps": round(alg_bytes / ms / 1e6, 1),
                  "lanczos_60_steps_ms": round(tl, 2), "steps_per_s": round(ncv / (tl / 1e3), 1)}
    y = ctx.download_resid()[:4].tolist()
    out[label]["resid_head"] = y
    ctx.close()
    print(json.dumps(out), flush=True)
