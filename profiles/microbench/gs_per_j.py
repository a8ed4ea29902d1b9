"""GB/s of the Gram-Schmidt step kernel per Lanczos step j (one launch = all sweeps of step j).
usage: python profiles/microbench/gs_per_j.py [f64|c64|dd] [m] [ncv]"""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
m = int(sys.argv[2]) if len(sys.argv) > 2 else (4000 if dt == "f64" else 2000)
ncv = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if dt == "c64":
    op = pkg.ArpackSimpleOp(syn.hermitian_stencil(m), dtype="c64"); S = 16
elif dt == "dd":
    op = pkg.ArpackSimpleOp(syn.laplacian2d(m), dtype="dd"); S = 16
else:
    op = pkg.ArpackSimpleOp(syn.laplacian2d(m)); S = 8
n = m * m; n_pad = (n + 255) // 256 * 256
ctx = pkg.Context(op, ncv, dots=os.environ.get('GA_DOTS') if dt == 'f64' else None)
ctx.getv0(); rn = ctx.norm2_resid()
H = np.zeros((ncv, 2) + ((2,) if dt == "dd" else ()))
info, rn = ctx.lanczos(0, ncv, H, rn)       # warm-up pass over all j
ctx.getv0(); rn = ctx.norm2_resid()
ctx.profile(True)
st0 = ctx.stats()
info, rn = ctx.lanczos(0, ncv, H, rn)
t = ctx.profile_gs_times(ncv)
st = ctx.stats()
nro = st["nrorth"] - st0["nrorth"]
rows = []
for j in range(1, ncv + 1):
    sweeps = 3 if nro >= ncv - 1 else 2     # DGKS fires on (almost) every step for these matrices
    b = ((j + 1) * sweeps + 2) * n_pad * S
    rows.append((j, round(float(t[j - 1]) * 1e3, 1), int(b / float(t[j - 1]) / 1e6)))
print(json.dumps({"dtype": dt, "n": n, "ncv": ncv, "nrorth": int(nro), "per_j_us_GBps": rows}))
ctx.close()
