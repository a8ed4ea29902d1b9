"""In-kernel timeline (globaltimer) of the persistent Gram-Schmidt step kernel, CTA 0."""
import sys, importlib, numpy as np
sys.path.insert(0, '/root/repo')
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + '.synthetic')
names = ["first tile", "stream end", "partials written", "barrier A done", "CTA0 reduced", "decided", "barrier B done"]
for m in (1414, 4000):
    csr = syn.laplacian2d(m); ncv = 40
    ctx = pkg.Context(pkg.ArpackSimpleOp(csr), ncv)
    ctx.getv0(); rn = ctx.norm2_resid(); H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, 26, H, rn)
    ctx.debug_trace()
    info, rn = ctx.lanczos(26, 1, H, rn)
    t = ctx.debug_trace(read=True).astype(np.int64)
    print("m=%d j=27: us since kernel entry" % m)
    for ph in range(3):
        print("  phase", ph, {n: round((t[1 + 8 * ph + i] - t[0]) / 1e3, 1) for i, n in enumerate(names)})
    ctx.close()
