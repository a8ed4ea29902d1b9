"""Time of Context creation (ga_create + ga_set_csr: allocation, CSR upload from pinned memory, host index passes, SpMV autotune)
for the C2 operator.  usage: python profiles/microbench/upload_bench.py [m]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry, bench
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
csr = syn.laplacian2d(m)
pins = [bench.pinned_copy(a) for a in csr[:3]]
op = pkg.ArpackSimpleOp((pins[0][0], pins[1][0], pins[2][0], csr[3]))
out = {}
for thr in ("1", "2", "4", "8", "16"):
    os.environ["GA_UPLOAD_THREADS"] = thr
    ts = []
    for rep in range(3):
        t0 = time.perf_counter(); ctx = pkg.Context(op, 40); ts.append(time.perf_counter() - t0); ctx.close()
    out["threads_" + thr] = [round(t, 4) for t in ts]
print(json.dumps(out))
