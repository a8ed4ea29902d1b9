#!/usr/bin/env python
"""bench.py -- Lanczos steps/s of the implicitly restarted Lanczos hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

Workload (BASELINE.json configs[1], "C2"): eigs(A, 10; ncv=40), which=:LM, Float64, A = 2-D 5-point
Laplacian on a 4000 x 4000 grid (n = 16 000 000, nnz = 79 984 000), start vector = dlarnv seed
(1,3,5,7).  With N > 1 the SAME problem is row-sharded over N GPUs ("scaling": "strong").

A bench "step" is one implicit-restart cycle of the solve: dsaitr extends the factorisation from
nev to ncv columns (np Lanczos steps: SpMV + fused CGS/DGKS sweeps), the host does the ncv x ncv
tridiagonal work, dsapps compresses the basis (V <- V Q).  `value` = Lanczos steps (operator
applications, the reference's stats.nopx) per second over the K timed cycles, inputs resident in
HBM.  `e2e` = the same metric through the public host-buffer API (upload A from pinned host memory,
start vector, K cycles, download the Ritz values/bounds + residual vector + wanted basis columns).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


def pinned_copy(a):
    """numpy view of a pinned (page-locked) host copy of `a` (torch is only the allocator here)."""
    try:
        import torch

        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    except Exception:
        return np.ascontiguousarray(a), None


def dist_setup(ngpus):
    """torch.distributed (NCCL) rendezvous when launched under torchrun; returns (rank, world, local_rank, dist)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1, 0, None
    import torch
    import torch.distributed as dist

    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def run_b200(args):
    pkg = entry.load_package()
    syn = importlib.import_module(entry.PKG_NAME + ".synthetic")
    rank, world, local, dist = dist_setup(args.gpus)
    m, nev, ncv = args.m, args.nev, args.ncv
    n = m * m
    blk = (n + world - 1) // world
    r0, r1 = rank * blk, min(n, (rank + 1) * blk)
    csr = syn.laplacian2d(m, row_range=(r0, r1) if world > 1 else None)
    nnz_global = 5 * n - 4 * m
    pins = []
    rowptr, t0_ = pinned_copy(csr[0]); col, t1_ = pinned_copy(csr[1]); val, t2_ = pinned_copy(csr[2])
    pins += [t0_, t1_, t2_]
    op = pkg.ArpackSimpleOp((rowptr, col, val, (r1 - r0, n)))

    plan = None
    if world > 1:
        gdist = importlib.import_module(entry.PKG_NAME + ".dist")
        plan = gdist.halo_plan(csr[1], n, rank, world)

    def make_ctx():
        if world == 1:
            return pkg.Context(op, ncv, device=local, dots=args.dots)
        ctx = pkg.Context(op, ncv, device=local, rank=rank, world_size=world, n_global=n, row_offset=r0, plan=plan, dots=args.dots)
        gdist.setup_comm(ctx, pkg.comm_unique_id, rank, world, fused=not args.nccl_only)
        return ctx

    def barrier():
        if dist is not None:
            import torch

            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------ device-resident measurement
    ctx = make_ctx()
    rc = ctx.symeigs_begin(nev, "LM", 0.0, max(300, int(np.ceil(np.sqrt(n)))), None)
    assert rc == 0, (rc, pkg.lib().ga_last_error(ctx.h))
    for _ in range(args.warmup):
        assert ctx.symeigs_iterate(1) == 0
    ctx.profile(True)
    st0 = ctx.stats()
    nopx0 = st0["nopx"]
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ctx.timer_start()
    done_steps = 0
    for _ in range(args.steps):
        rc = ctx.symeigs_iterate(1)
        done_steps += 1
        if rc != 0:
            break
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)
    st = ctx.stats()
    prof = ctx.profile_read()
    lanczos_steps = st["nopx"] - nopx0
    value = lanczos_steps / (ms / 1e3)
    peak, peak_src = measured_peaks()
    S = 8
    n_pad = ((blk + 255) // 256) * 256
    gs_bytes = prof["gs_units"] * n_pad * S
    gs_ms = max(prof["gs_ms"], 1e-9)
    achieved = gs_bytes / (gs_ms / 1e3) / 1e9
    # algorithmic bytes of the whole timed region by SURVEY 8(d): per Lanczos step SpMV + 2nS + GS rounds
    spmv_bytes = (nnz_global * (S + 4) + (n + 1) * 8 + 2 * n * S) / world
    step_bytes_total = lanczos_steps * (spmv_bytes + 2 * blk * S) + gs_bytes
    # DRAM traffic of one captured launch of the same kernel (ncu --set full, profiles/gs_traffic.json)
    traffic, traffic_note = None, None
    tfile = os.path.join(ROOT, "profiles", "gs_traffic.json")
    if os.path.exists(tfile):
        try:
            tj = json.load(open(tfile))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = tj.get("note")
        except Exception:
            traffic = None
    ngs = max(prof.get("gs_launches", 0), 1)
    step_kernel = ngs < prof["gs_runs"]
    roofline = {
        "bound": "hbm",
        "kernel": ("k_gs_step (all CGS/DGKS sweeps of a Lanczos step in one persistent cooperative launch, TMA pipeline)"
                   if step_kernel else "k_gs (fused CGS/DGKS sweep, TMA pipeline)"),
        "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "peak_source": peak_src, "traffic": traffic, "traffic_note": traffic_note,
        "bytes_per_launch": round(gs_bytes / ngs), "launches_timed": ngs, "sweeps_timed": prof["gs_runs"],
        "avg_launch_ms": round(gs_ms / ngs, 4),
        "gs_share_of_step_time": round(gs_ms / ms, 4),
        "whole_step_achieved_GBps": round(step_bytes_total / (ms / 1e3) / 1e9 * world, 1),
        "frac_of_nominal_8TBps": round(achieved / 8000.0, 4),
    }
    launches = prof["launches"]
    timeline = None
    if args.trace:
        # in-kernel globaltimer timeline (CTA 0) of the last Gram-Schmidt step kernel of one more cycle
        ctx.debug_trace()
        ctx.symeigs_iterate(1)
        t = ctx.debug_trace(read=True).astype(np.int64)
        names = ["first_tile", "stream_end", "partials_written", "barrier_A", "cta0_reduced", "exchanged_decided", "barrier_B"]
        timeline = [{nm: round((int(t[1 + 8 * ph + i]) - int(t[0])) / 1e3, 1) for i, nm in enumerate(names)} for ph in range(3)]
    ctx.close()

    # ------------------------------------------------------------ end to end through the host-buffer API
    e2e = None
    if not args.no_e2e:
        # pinned host destinations for the results (allocated outside the timed region, like the inputs)
        e2e_bufs = [pinned_copy(np.zeros(r1 - r0))[0], pinned_copy(np.zeros((nev, r1 - r0)))[0]]
        barrier()
        t0 = time.perf_counter()
        ctx2 = make_ctx()                       # allocations + H2D of the CSR matrix (pinned host buffers)
        ta = time.perf_counter()
        rc = ctx2.symeigs_begin(nev, "LM", 0.0, args.steps, None)
        tb = time.perf_counter()
        rcit = ctx2.symeigs_iterate(args.steps)
        tc = time.perf_counter()
        resid = ctx2.download_resid(out=e2e_bufs[0])           # D2H into pinned host buffers: residual vector
        Vk = ctx2.download_v(0, nev, out=e2e_bufs[1])          # D2H: the nev wanted basis columns
        st2 = ctx2.stats()
        if dist is not None:
            import torch

            torch.cuda.synchronize()
        t1 = time.perf_counter()
        dt = max_over_ranks(t1 - t0)
        h2d = rowptr.nbytes // 2 + col.nbytes + val.nbytes  # rowptr is narrowed to int32 before the copy
        d2h = resid.nbytes + Vk.nbytes
        e2e = {"value": round(st2["nopx"] / dt, 2), "unit": "steps/s", "h2d_bytes_per_step": int(h2d // max(args.steps, 1)),
               "d2h_bytes_per_step": int(d2h // max(args.steps, 1)), "seconds": round(dt, 3), "lanczos_steps": int(st2["nopx"]),
               "breakdown_s": {"context_and_upload": round(ta - t0, 3), "start_vector_and_first_nev_steps": round(tb - ta, 3),
                               "restart_cycles": round(tc - tb, 3), "download": round(t1 - tc, 3)},
               "what": "Context(A_host)+symeigs_begin+%d restart cycles+download resid and V[:,1:nev]; per-rank bytes" % args.steps}
        ctx2.close()
        del resid, Vk

    # ------------------------------------------------------------ time to solution (optional, long)
    tts = None
    if args.tts:
        barrier()
        t0 = time.perf_counter()
        ctx3 = make_ctx()
        rc = ctx3.symeigs_begin(nev, "LM", args.tts_tol, args.tts_maxiter if args.tts_maxiter > 0 else 4 * m, None)
        rcit = 0
        while rcit == 0 and time.perf_counter() - t0 < args.tts_budget:     # chunks of 50 restart cycles, wall-clock guard
            rcit = ctx3.symeigs_iterate(50)
        info, values, bounds, vecs, iparam = ctx3.symeigs_end(nev, ritzvec=True)
        dt = max_over_ranks(time.perf_counter() - t0)
        exact = syn.laplacian2d_eigs(m, nev)
        st3 = ctx3.stats()
        tts = {"seconds": round(dt, 2), "converged": bool(rcit == 99 and info == 0 and int(iparam[4]) >= nev), "info": int(info),
               "tol": args.tts_tol, "restarts": int(iparam[2]), "nconv": int(iparam[4]),
               "lanczos_steps": int(st3["nopx"]), "nrorth": int(st3["nrorth"]),
               "max_rel_err_vs_analytic": float(np.max(np.abs(values - exact) / exact)) if len(values) == nev else None,
               "values": [float(v) for v in values]}
        ctx3.close()

    # ------------------------------------------------------------ CPU baseline (oracle port) on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(syn, m, nev, ncv, budget_s=args.cpu_budget)

    if rank == 0:
        out = {
            "metric": "lanczos_steps_per_sec", "value": round(value, 2), "unit": "steps/s", "n_gpus": world,
            "steps": done_steps, "warmup": args.warmup, "ms_per_step": round(ms / max(done_steps, 1), 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: eigs(A,%d; ncv=%d, which=:LM) Float64, 2-D 5-point Laplacian %dx%d grid (n=%d, nnz=%d), row-sharded over %d GPU(s)"
                                   % (nev, ncv, m, m, n, nnz_global, world),
                       "step": "one implicit-restart cycle (np Lanczos steps + tridiagonal host work + V*Q)",
                       "lanczos_steps_timed": int(lanczos_steps), "l2": "inputs larger than L2 (V is %.1f GB per GPU)" % (n_pad * ncv * 8 / 1e9),
                       "nrorth": int(st["nrorth"] - st0["nrorth"]), "nitref": int(st["nitref"] - st0["nitref"]),
                       "ms_per_cycle_breakdown": {k: round((st[k] - st0[k]) * 1e3 / max(done_steps, 1), 3)
                                                  for k in ("taitr", "tapps", "teigt", "tgets", "tconv")}},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if tts is not None:
            out["time_to_solution"] = tts
        if timeline is not None:
            out["gs_step_timeline_us"] = timeline
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def cpu_sample(syn, m, nev, ncv, budget_s=30.0, threads=None):
    """The CPU oracle (port of the reference algorithm, OpenMP on the V sweeps and the SpMV) timed
    on a bounded sample of the same workload: Lanczos steps at the middle columns of a restart
    cycle (so the swept basis width is the cycle's average)."""
    O = entry.load_oracle()
    A = syn.to_scipy(syn.laplacian2d(m))
    ora = O.Oracle(A, ncv)
    ora.dgetv0()
    t0 = time.perf_counter()
    ora.dsaitr(0, nev)
    per_step = (time.perf_counter() - t0) / nev
    np_ = ncv - nev
    # cost grows ~linearly with the column index; mid-cycle steps cost ~ (nev+np/2)/(nev/2) of the first ones
    est_mid = per_step * (nev + np_ / 2.0) / (nev / 2.0 + 1.0)
    S = int(max(2, min(np_, budget_s / max(est_mid, 1e-6))))
    k0 = nev + (np_ - S) // 2
    # fill columns nev..k0 so the sampled steps sweep genuine basis vectors (untimed)
    if k0 > nev:
        ora.dsaitr(nev, k0 - nev)
    ora.reset_stats()
    t0 = time.perf_counter()
    ora.dsaitr(k0, S)
    dt = time.perf_counter() - t0
    st = ora.stats()
    return {"value": round(st["nopx"] / dt, 3), "unit": "steps/s", "cores": O.num_threads(), "kind": "port",
            "sample": "%d Lanczos steps at columns %d..%d of a restart cycle (n=%d, same matrix), oracle C++ -O2 -fopenmp; "
                      "the reference itself is pure Julia and cannot run here" % (S, k0 + 1, k0 + S, m * m),
            "seconds": round(dt, 2)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (the oracle port; Julia is not installed, so
    there is no oracle/_ref) on the box's host cores, same config/metric.  Each bench step is a
    bounded sample: S Lanczos steps at the middle columns of a restart cycle."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    syn = importlib.import_module(entry.load_package().__name__ + ".synthetic")
    O = entry.load_oracle()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; only rank 0 works here, with all the host cores it may use
        try:
            O.set_num_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            O.set_num_threads(os.cpu_count() or 1)
    m, nev, ncv = args.m, args.nev, args.ncv
    n = m * m
    A = syn.to_scipy(syn.laplacian2d(m))
    ora = O.Oracle(A, ncv)
    ora.dgetv0()
    t0 = time.perf_counter()
    ora.dsaitr(0, nev)
    per_step = (time.perf_counter() - t0) / nev
    np_ = ncv - nev
    est_mid = per_step * (nev + np_ / 2.0) / (nev / 2.0 + 1.0)
    total_budget = 150.0
    S = int(max(1, min(np_, total_budget / max(1, args.steps + args.warmup) / max(est_mid, 1e-6))))
    k0 = nev + (np_ - S) // 2
    if k0 > nev:
        ora.dsaitr(nev, k0 - nev)
    for _ in range(args.warmup):
        ora.dsaitr(k0, S)
    ora.reset_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.dsaitr(k0, S)
    dt = time.perf_counter() - t0
    st = ora.stats()
    value = st["nopx"] / dt
    cores = O.num_threads()
    sample = "each step = %d Lanczos steps at columns %d..%d of a restart cycle, oracle port (C++ -O2 -fopenmp), %d threads" % (S, k0 + 1, k0 + S, cores)
    out = {
        "impl": "reference", "metric": "lanczos_steps_per_sec", "value": round(value, 3), "unit": "steps/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / max(args.steps, 1) * 1e3, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: eigs(A,%d; ncv=%d, which=:LM) Float64, 2-D 5-point Laplacian %dx%d grid (n=%d), CPU host cores" % (nev, ncv, m, m, n),
                   "step": sample},
        "cpu_baseline": {"value": round(value, 3), "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", "--grid", dest="m", type=int, default=4000, help="grid side (n = m*m); 4000 = BASELINE config C2")
    ap.add_argument("--nev", type=int, default=10)
    ap.add_argument("--ncv", type=int, default=40)
    ap.add_argument("--tts", action="store_true", help="also run the full solve (time to solution)")
    ap.add_argument("--tts-tol", type=float, default=1e-10, help="tolerance of the time-to-solution solve (north_star: eigenvalues to 1e-10)")
    ap.add_argument("--tts-maxiter", type=int, default=0, help="restart cap of the time-to-solution solve (0: 4 x grid side)")
    ap.add_argument("--tts-budget", type=float, default=420.0, help="wall-clock guard of the time-to-solution solve, seconds")
    ap.add_argument("--dots", default="exact", choices=["exact", "plain"], help="Float64 dot-product accumulators of the sweeps (library default: exact = Dot2)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--trace", action="store_true", help="add the in-kernel timeline of one Gram-Schmidt step kernel (debug)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nccl-only", action="store_true", help="multi-GPU: NCCL all-gather/send-recv instead of the fused peer-memory path")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
