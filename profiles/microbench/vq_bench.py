"""Time of the implicit-restart basis update V <- V*Q (ga_apply_shifts -> k_vq) per call, with the HBM traffic it stands for.
usage: python profiles/microbench/vq_bench.py [f64|c64|dd|f32] [m] [ncv] [kev]"""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
dt = sys.argv[1] if len(sys.argv) > 1 else "f64"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
ncv = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kev = int(sys.argv[4]) if len(sys.argv) > 4 else 10
S = {"f64": 8, "c64": 16, "dd": 16, "f32": 4}[dt]
csr = syn.hermitian_stencil(m) if dt == "c64" else syn.laplacian2d(m)
op = pkg.ArpackSimpleOp(csr, dtype=dt)
n = m * m
ctx = pkg.Context(op, ncv)
ctx.getv0(); rn = ctx.norm2_resid()
H = np.zeros((ncv, 2) + ((2,) if dt == "dd" else ()))
info, rn = ctx.lanczos(0, ncv, H, rn)
rng = np.random.default_rng(0)
Q, _ = np.linalg.qr(rng.standard_normal((ncv, ncv)))
Q = np.asfortranarray(Q)
if dt == "dd":
    Q = np.stack([Q, np.zeros_like(Q)], axis=-1)
np_ = ncv - kev
beta = 0.5
reps = 10
ctx.apply_shifts(kev, np_, Q, beta)       # warm-up
ctx.timer_start()
for _ in range(reps):
    ctx.apply_shifts(kev, np_, Q, beta)
ms = ctx.timer_stop() / reps
n_pad = (n + 511) // 512 * 512
bytes_ = n_pad * S * (ncv + (kev + 1) + 2)   # read kplusp columns, write kev+1, read+write resid
print(json.dumps({"dtype": dt, "n": n, "ncv": ncv, "kev": kev, "ms_per_call": round(ms, 4), "GB": round(bytes_ / 1e9, 3),
                  "GBps": round(bytes_ / ms / 1e6, 1), "env": {k: os.environ[k] for k in ("GA_VQ_MIN2K",) if k in os.environ},
                  "note": "includes the host->device copy of Q and one host sync per call"}))
ctx.close()
