import importlib, os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package(); O = entry.load_oracle()
rng = np.random.default_rng(5)
n = 7001
A = sp.random(n, n, density=3.0 / n, random_state=rng, format="lil", dtype=np.float64)
A[17, :] = rng.standard_normal(n)
A[4000, : 2049] = 1.5
A[4001, : 2048] = -0.5
for r in range(100, 160):
    A[r, :] = 0.0
for r in range(5000, 5040):
    A[r, 100:700] = rng.standard_normal(600)
A = A.tocsr(); A.eliminate_zeros()
x, _ = O.dlarnv(n)
outs = []
for v1 in ("1", "0"):
    os.environ["GA_SPMV_V1"] = v1
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), 2)
    outs.append(np.array(ctx.opx(x))); ctx.close()
d = np.nonzero(outs[0] != outs[1])[0]
nnzr = np.diff(A.indptr)
yref = A @ x
print("differing rows:", len(d))
for r in d[:40]:
    print(r, nnzr[r], outs[0][r], outs[1][r], yref[r])
