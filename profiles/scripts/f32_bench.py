#!/usr/bin/env python
"""Float32-vector variant of BASELINE config C2 on one B200: eigs(Float32, Float64, A, 10; ncv=40), 2-D Laplacian
m x m (default 4000 -> n = 16M): Lanczos steps/s over `cycles` restart cycles after a warm-up, Gram-Schmidt step
kernel GB/s (algorithmic bytes, 4-byte elements) and SpMV GB/s, next to the Float64 numbers of the same run.
    python profiles/scripts/f32_bench.py [m] [cycles] [f32,f64]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
only = sys.argv[3].split(",") if len(sys.argv) > 3 else ["f32", "f64"]
pkg = entry.load_package(); syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
rowptr, col, val, shape = syn.laplacian2d(m)
n, nnz, nev, ncv = m * m, len(col), 10, 40
out = {"workload": "C2 operator, n=%d, nnz=%d, nev=10, ncv=40, 1 GPU" % (n, nnz), "hbm_peak_measured_GBps": peak}
for name, S, v in (("f32", 4, val.astype(np.float32)), ("f64", 8, val)):
    if name not in only:
        continue
    ctx = pkg.Context(pkg.ArpackSimpleOp((rowptr, col, v, shape), dtype=name), ncv)
    pad = 512 if S == 4 else 256
    n_pad = (n + pad - 1) // pad * pad
    assert ctx.symeigs_begin(nev, "LM", 0.0, 10 ** 6, None) == 0
    ctx.symeigs_iterate(3)                                   # warm-up cycles
    st0 = ctx.stats()
    ctx.profile(True)
    ctx.timer_start()
    rc = ctx.symeigs_iterate(cycles)
    ms = ctx.timer_stop()
    st = ctx.stats(); prof = ctx.profile_read()
    steps = st["nopx"] - st0["nopx"]
    spmv_ms = ctx.bench_opx(20)
    spmv_bytes = nnz * (S + 4) + (n + 1) * 4 + 2 * n * S
    gs_gbps = prof["gs_units"] * n_pad * S / max(prof["gs_ms"], 1e-9) / 1e6
    out[name] = {"rc": int(rc), "lanczos_steps": int(steps), "steps_per_s": round(steps / (ms / 1e3), 1), "ms_per_cycle": round(ms / cycles, 3),
                 "k_gs_step_GBps": round(gs_gbps, 1), "k_gs_step_frac_of_measured_peak": round(gs_gbps / peak, 3),
                 "gs_share_of_time": round(prof["gs_ms"] / ms, 3), "spmv_ms": round(spmv_ms, 4),
                 "spmv_GBps": round(spmv_bytes / spmv_ms / 1e6, 1), "nrorth": int(st["nrorth"] - st0["nrorth"]),
                 "V_GB": round(n_pad * ncv * S / 1e9, 2)}
    ctx.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "f32_bench_m%d.json" % m), "w"), indent=1)
print(json.dumps(out))
