"""OP*x alone on device-resident vectors: GB/s of the CSR SpMV for the BASELINE matrices.
usage: python profiles/microbench/spmv_bench.py [lap|c4|c5|c1] ; env GA_SPMV_V1 / GA_SPMV_STAGES select the kernel."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
which = sys.argv[1] if len(sys.argv) > 1 else "lap"
if which == "lap":
    csr = syn.laplacian2d(4000); S = 8; op = pkg.ArpackSimpleOp(csr)
elif which == "c4":
    csr = syn.hermitian_stencil(2000); S = 16; op = pkg.ArpackSimpleOp(csr, dtype="c64")
elif which == "c1":
    csr = syn.sprand_sym(100000); S = 8; op = pkg.ArpackSimpleOp(csr)
elif which == "c5":
    csr = syn.tall_sparse(5000000, 200000, 10); S = 8; op = pkg.ArpackNormalOp(csr)
rowptr, col, val, shape = csr
nnz = int(rowptr[-1]); m, n = shape
ctx = pkg.Context(op, 2)
ctx.getv0(); rn = ctx.norm2_resid(); H = np.zeros((2, 2)); ctx.lanczos(0, 1, H, rn)   # V[:,0] = unit vector
ms = ctx.bench_opx(20)
if which == "c5":
    bytes_ = 2 * (nnz * (S + 4) + (m + 1) * 4) + S * (2 * n + 2 * m)
else:
    bytes_ = nnz * (S + 4) + (m + 1) * 4 + 2 * m * S
print(json.dumps({"matrix": which, "n": m, "nnz": nnz, "ms": round(ms, 4), "GBps": round(bytes_ / ms / 1e6, 1),
                  "v1": os.environ.get("GA_SPMV_V1", "0"), "stages": os.environ.get("GA_SPMV_STAGES", "default")}))
ctx.close()
