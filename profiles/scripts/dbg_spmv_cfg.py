"""Streamed-SpMV configuration sweep on the complex stencil (C4): for every (groups, stages) pair
  A. repeated OP*x on a fixed random x, bitwise against the first (non-streamed) kernel and against SciPy
  B. one restart-free Lanczos run of ncv steps: three-term relation per step
  C. restart cycles of the whole solve: digest of resid after every cycle, against the first kernel's
and the in-kernel stale-stage self-check counter (ga_debug_spmv_check).  Output: one JSON line per config.
argv: m [reps] [cycles]"""
import hashlib, importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import __graft_entry__ as entry
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
import pickle, subprocess
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 10
nev, ncv = 6, 30
CACHE = "/tmp/dbg_spmv_m%d.pkl" % m
BASE = "/tmp/dbg_spmv_base_m%d.pkl" % m
if len(sys.argv) <= 4:
    # driver: build the matrix once, one worker process per configuration (a faulting kernel kills its CUDA context)
    csr = syn.hermitian_stencil(m); A = syn.to_scipy(csr); n = A.shape[0]
    rng = np.random.default_rng(1); x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    pickle.dump({"csr": csr, "x": x, "ref": A @ x}, open(CACHE, "wb"), protocol=4)
    if os.path.exists(BASE):
        os.remove(BASE)
    cfgs = os.environ.get("DBG_CFGS", "0,0,1 4,6,1 4,5,1 6,8,1 4,8,1 6,6,1 4,4,1 0,0,0 4,8,0 6,8,0").split()
    for cfg in cfgs:
        subprocess.run([sys.executable, __file__, str(m), str(reps), str(cycles), cfg], timeout=120)
    sys.exit(0)
dat = pickle.load(open(CACHE, "rb")); csr, x, ref = dat["csr"], dat["x"], dat["ref"]
A = syn.to_scipy(csr); n = A.shape[0]
base = pickle.load(open(BASE, "rb")) if os.path.exists(BASE) else {}
configs = [tuple(int(v) for v in sys.argv[4].split(","))]
for g, s, step in configs:
    t0 = time.time()
    out = {"groups": g, "stages": s, "step_kernel": step}
    try:
        os.environ["GA_NO_STEP_KERNEL"] = "0" if step else "1"
        ctx = pkg.Context(pkg.ArpackSimpleOp(csr, dtype="c64"), ncv)
        ctx.tune_spmv(g, s)
        ctx.debug_spmv_check()
        if step:
            # A
            nbad_rep, worst, firstbad = 0, 0.0, []
            for r in range(reps):
                y = np.asarray(ctx.opx(x))
                if "y" not in base:
                    base["y"] = y.copy()
                d = np.abs(y - ref)
                if not np.array_equal(y, base["y"]):
                    nbad_rep += 1
                    bad = np.nonzero(y != base["y"])[0]
                    if len(firstbad) < 3:
                        firstbad.append([int(bad[0]), int(bad[-1]), int(len(bad))])
                worst = max(worst, float(d.max()))
            out["A"] = {"reps": reps, "nbad_rep": nbad_rep, "max_err_vs_scipy": worst, "bad_rows_first_last_count": firstbad}
            # B
            ctx.getv0(); rn = ctx.norm2_resid(); H = np.zeros((ncv, 2))
            info, rn = ctx.lanczos(0, ncv, H, rn)
            V = ctx.download_v(0, ncv)
            badsteps = []
            for j in range(1, ncv - 1):
                lhs = A @ V[:, j]
                rhs = H[j, 0] * V[:, j - 1] + H[j, 1] * V[:, j] + H[j + 1, 0] * V[:, j + 1]
                d = np.abs(lhs - rhs)
                e = float(np.linalg.norm(lhs - rhs) / np.linalg.norm(lhs))
                if e > 1e-10:
                    idx = np.nonzero(d > 1e-8)[0]
                    badsteps.append([j, e, int(idx[0]) if len(idx) else -1, int(idx[-1]) if len(idx) else -1, int(len(idx))])
            out["B"] = {"nbad_steps": len(badsteps), "bad": badsteps[:4]}
            ctx.close()
            ctx = pkg.Context(pkg.ArpackSimpleOp(csr, dtype="c64"), ncv)
            ctx.tune_spmv(g, s)
            ctx.debug_spmv_check()
        # C
        rc = ctx.symeigs_begin(nev, "LM", 0.0, 300, None)
        assert rc == 0, rc
        key = "C%d" % step
        digs, firstdiff = [], -1
        for cyc in range(cycles):
            r = ctx.symeigs_iterate(1)
            dg = hashlib.sha1(ctx.download_resid().tobytes()).hexdigest()[:12]
            digs.append(dg)
            if r != 0:
                break
        if key not in base:
            base[key] = digs
        for i, (a, b) in enumerate(zip(digs, base[key])):
            if a != b:
                firstdiff = i
                break
        st = ctx.stats()
        out["C"] = {"cycles": len(digs), "first_cycle_differing_from_v1": firstdiff, "nopx": st["nopx"], "nrorth": st["nrorth"]}
        chk = ctx.debug_spmv_check(read=True)
        out["stale_blocks"] = int(chk[0])
        out["stale_records"] = [[int(v) for v in chk[8 + 8 * i: 16 + 8 * i]] for i in range(min(int(chk[0]), 6))]
        ctx.close()
    except Exception as e:
        out["err"] = str(e)[:300]
    out["sec"] = round(time.time() - t0, 2)
    print(json.dumps(out), flush=True)
    pickle.dump(base, open(BASE, "wb"), protocol=4)
