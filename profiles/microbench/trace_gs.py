import sys, importlib, numpy as np
sys.path.insert(0,'/root/repo')
import __graft_entry__ as entry
pkg=entry.load_package(); syn=importlib.import_module(entry.PKG_NAME+'.synthetic')
for m in (1414, 4000):
    csr=syn.laplacian2d(m); ncv=40
    ctx=pkg.Context(pkg.ArpackSimpleOp(csr), ncv)
    ctx.getv0(); rn=ctx.norm2_resid(); H=np.zeros((ncv,2))
    info,rn=ctx.lanczos(0,26,H,rn)
    ctx.debug_trace()
    for k in (26,27,28):
        info,rn=ctx.lanczos(k,1,H,rn)
        t=ctx.debug_trace(read=True).astype(np.int64)
        print(m, "j=%d"%(k+1), "last sweep timeline (us from entry):", [round((x-t[0])/1e3,1) for x in t[1:8]])
    ctx.close()
