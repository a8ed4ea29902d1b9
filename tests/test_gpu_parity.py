"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, against the reference's golden vectors / KATs, and -- at BASELINE.json's full size --
through size-independent properties (orthonormality, the Lanczos three-term relation, analytic
spectrum).  Tolerances follow BASELINE.json's north_star: start vector bit-exact; Float64
eigenvalues <= 1e-10 relative; Double64 <= 1e-26; restart counts within +-1 of the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD, GOLDEN_V0, kat15

pytestmark = pytest.mark.gpu

DT = {"f64": 0, "c64": 1, "dd": 2}


@pytest.fixture(autouse=True)
def _oracle_exact_arithmetic(O):
    """The library's Float64 sweeps accumulate h = V'w error-free (Dot2, GA_DOT_EXACT, the default); the oracle is run
    the same way (oracle.hpp ARITH_EXACT), so both sides sit within one rounding of the exact dot products and the
    Lanczos trajectories -- 0.717 tests, deflations, restart counts -- agree step for step."""
    with O.arith("exact"):
        yield


def _dd_float(x):
    return x[..., 0] + x[..., 1]


# ---------------------------------------------------------------- dgetv0: start vector, bit-exact
@pytest.mark.parametrize("dtype", ["f64", "c64", "dd"])
@pytest.mark.parametrize("n", [10, 129, 4097, 100003])
def test_start_vector_bit_exact(pkg, O, dtype, n):
    op = pkg.ArpackSimpleOp(sp.identity(n, format="csr"), dtype=dtype)
    ctx = pkg.Context(op, 2)
    ierr, rnorm = ctx.getv0(initv=False, j=1)
    x = np.asarray(ctx.download_resid())
    ref, seed = O.dlarnv(n, dtype=DT[dtype])
    assert ierr == 0
    assert x.tobytes() == np.asarray(ref).tobytes()          # bit-exact (integer LCG)
    assert tuple(ctx.iseed) == seed
    if n == 10 and dtype == "f64":
        assert np.array_equal(x, np.array(GOLDEN_V0))        # test/dgetv0_simple.jl:6-35
    nr = O.norm2(ref, DT[dtype])
    if dtype == "dd":
        assert abs(_dd_float(np.asarray(rnorm) - 0) - _dd_float(nr)) <= 4e-30 * _dd_float(nr) or \
            abs((rnorm[0] - nr[0]) + (rnorm[1] - nr[1])) <= 1e-29 * nr[0]
    else:
        assert abs(rnorm - nr) <= np.spacing(nr)             # within 1 ulp of the x87 norm
    ctx.close()


def test_norm2_overflow_safe(pkg, O):
    """norm2 semantics (src/arpack-blas-nrm2.jl): no overflow/underflow, ~correctly rounded."""
    import mpmath as mp

    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    n = 5000
    op = pkg.ArpackSimpleOp(sp.identity(n, format="csr"))
    ctx = pkg.Context(op, 2)
    for scale in (1.0, 1e200, 1e-200, 1e-300):
        x = rng.standard_normal(n) * scale
        ctx.upload_resid(x)
        got = ctx.norm2_resid()
        exact = mp.sqrt(sum(mp.mpf(float(v)) ** 2 for v in x))
        assert abs(mp.mpf(got) - exact) <= abs(exact) * 2.0 ** -52, scale
    x = np.zeros(n)
    x[[3, 77, 4000]] = [1e300, 1e-300, 3e10]   # mixed ranges (Blue's three accumulators)
    ctx.upload_resid(x)
    assert ctx.norm2_resid() == pytest.approx(1e300, rel=1e-15)
    ctx.upload_resid(np.zeros(n))
    assert ctx.norm2_resid() == 0.0                           # test/runtests.jl:164-166
    ctx.close()


# ---------------------------------------------------------------- OP*x
@pytest.mark.parametrize("dtype", ["f64", "c64", "dd"])
def test_opx(pkg, syn, O, dtype):
    if dtype == "c64":
        csr = syn.hermitian_stencil(37)
    else:
        csr = syn.sprand_sym(3000, 5.0, seed=11)
    A = syn.to_scipy(csr)
    n = A.shape[0]
    # add one very long row/column to exercise the long-row path
    if dtype == "f64":
        A = A.tolil()
        A[5, :] = 1.0
        A[:, 5] = 1.0
        A = A.tocsr()
    op = pkg.ArpackSimpleOp(A, dtype=dtype)
    ctx = pkg.Context(op, 2)
    x, _ = O.dlarnv(n, dtype=DT[dtype])
    y = np.asarray(ctx.opx(pkg.as_dd(x) if dtype == "dd" else x))
    if dtype == "dd":
        yref = A @ x[:, 0]
        assert np.allclose(_dd_float(y), yref, rtol=1e-14, atol=1e-14)
        yo = O.Oracle(A, 2, dtype=O.DD).opx(x)
        assert np.max(np.abs((y[:, 0] - yo[:, 0]) + (y[:, 1] - yo[:, 1]))) <= 1e-28 * np.max(np.abs(yref))
    else:
        yref = A @ x
        assert np.allclose(y, yref, rtol=1e-13, atol=1e-13)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "c64", "dd"])
def test_opx_stream_vs_first_kernel_bitwise(pkg, O, dtype, monkeypatch):
    """The TMA-streamed SpMV and the first row-segment kernel do the same arithmetic in the same
    order: bit-identical on a matrix with empty rows, ragged segment alignment, a dense band and a
    row that takes the long-row path in both (7001 non-zeros).  Rows with 2049..4096 non-zeros are
    summed sequentially by the first kernel and by the long-row tree in the streamed one: those
    agree to rounding only."""
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    n = 7001
    A = sp.random(n, n, density=3.0 / n, random_state=rng, format="lil", dtype=np.float64)
    A[17, :] = rng.standard_normal(n)            # 7001 non-zeros: long-row path
    A[4000, : 2049] = 1.5                         # just above one stage
    A[4001, : 2048] = -0.5                        # exactly one stage
    for r in range(100, 160):                     # empty rows
        A[r, :] = 0.0
    for r in range(5000, 5040):                   # dense-ish band: few rows per block
        A[r, 100:700] = rng.standard_normal(600)
    A = A.tocsr()
    A.eliminate_zeros()
    if dtype == "c64":
        B = A.copy().astype(np.complex128)
        B.data = B.data * np.exp(1j * rng.uniform(0, 6.28, B.nnz))
        A = B
    x, _ = O.dlarnv(n, dtype=DT[dtype])
    xin = pkg.as_dd(x) if dtype == "dd" else x
    outs = []
    for v1 in ("1", "0"):
        monkeypatch.setenv("GA_SPMV_V1", v1)
        ctx = pkg.Context(pkg.ArpackSimpleOp(A, dtype=dtype), 2)
        outs.append(np.array(ctx.opx(xin)))
        ctx.close()
    nnz_row = np.diff(A.indptr)
    cap = 2048 if dtype == "f64" else 1024          # SpmvStreamCfg<T>::CAP; the first kernel's is 4096 / 2048
    same_path = (nnz_row <= cap) | (nnz_row > 2 * cap)
    assert same_path.sum() >= n - 45
    assert np.array_equal(outs[0][same_path].view(np.uint64), outs[1][same_path].view(np.uint64))
    yref = A @ (x[:, 0] if dtype == "dd" else x)
    got = _dd_float(outs[1]) if dtype == "dd" else outs[1]
    assert np.allclose(got, yref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("dtype", ["f64", "c64"])
def test_opx_stream_every_ring_config_bitwise(pkg, syn, dtype):
    """Every (consumer groups, stages) pair of the streamed SpMV computes the same bits as the first
    kernel, repeatedly, on a matrix large enough for the persistent ring to wrap many times (>= 4 x 148
    row blocks, the size from which the upload-time autotune picks among these configurations).  The
    shared ring is only correct because consumers re-check which iteration a stage holds (bulk copies
    complete out of issue order on B200); the in-kernel self-check must stay silent."""
    m = 640
    csr = syn.hermitian_stencil(m) if dtype == "c64" else syn.laplacian2d(m)
    n = m * m
    rng = np.random.default_rng(11)
    x = rng.standard_normal(n) + (1j * rng.standard_normal(n) if dtype == "c64" else 0.0)
    ctx = pkg.Context(pkg.ArpackSimpleOp(csr, dtype=dtype), 2)
    ctx.tune_spmv(0, 0)
    y0 = np.array(ctx.opx(x))
    assert np.allclose(y0, syn.to_scipy(csr) @ x, rtol=1e-13, atol=1e-13)
    ctx.debug_spmv_check()
    for g, st in [(4, 4), (4, 5), (4, 6), (4, 7), (4, 8), (6, 6), (6, 7), (6, 8)]:
        ctx.tune_spmv(g, st)
        for rep in range(6):
            y = np.array(ctx.opx(x))
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), (g, st, rep)
    assert int(ctx.debug_spmv_check(read=True)[0]) == 0
    with pytest.raises(pkg.ArpackException):
        ctx.tune_spmv(6, 5)
    ctx.close()


def test_normal_opx(pkg, syn, O):
    A = syn.to_scipy(syn.tall_sparse(5000, 300, 6))
    op = pkg.ArpackNormalOp(A)
    ctx = pkg.Context(op, 4)
    x, _ = O.dlarnv(300)
    y = ctx.opx(x)
    assert np.allclose(y, A.T @ (A @ x), rtol=1e-12)
    ctx.close()
    op2 = pkg.ArpackNormalOp(A.T.tocsr())
    ctx2 = pkg.Context(op2, 4)
    assert np.allclose(ctx2.opx(x), A.T @ (A @ x), rtol=1e-12)
    ctx2.close()


@pytest.mark.parametrize("dtype", ["f64", "c64"])
def test_normal_opx_column_blocked(pkg, syn, O, dtype, monkeypatch):
    """The second product of the normal operator in column blocks (k_spmv.cu:normal_wide; automatic above 64 MiB of long
    vector, forced here with GA_NORMAL_COLBLOCK so that a small matrix takes 10 blocks): every block continues the row's
    running sum, so the result is the same bits as the unblocked product -- for a tall A (blocks of A^H), for a wide A
    (blocks of A), with rows that are empty inside a block, and through a whole svds."""
    A = syn.to_scipy(syn.tall_sparse(5000, 300, 6))
    if dtype == "c64":
        A = (A + 1j * A.multiply(A > 0.5)).tocsr()
    x, _ = O.dlarnv(300, dtype=O.C64 if dtype == "c64" else O.F64)
    x = np.asarray(x)
    ref = A.conj().T @ (A @ x)
    res = {}
    for blocked in (False, True):
        if blocked:
            monkeypatch.setenv("GA_NORMAL_COLBLOCK", "512")
        else:
            monkeypatch.delenv("GA_NORMAL_COLBLOCK", raising=False)
        for name, M in (("tall", A), ("wide", A.conj().T.tocsr())):
            ctx = pkg.Context(pkg.ArpackNormalOp(M, dtype=dtype), 4)
            res[(name, blocked)] = np.asarray(ctx.opx(x)).copy()
            ctx.close()
    for name in ("tall", "wide"):
        assert np.allclose(res[(name, False)], ref, rtol=1e-12)
        assert np.array_equal(res[(name, True)], res[(name, False)])
    if dtype == "f64":
        monkeypatch.setenv("GA_NORMAL_COLBLOCK", "512")
        U, S, V = pkg.svds(A, 3, ncv=12)
        monkeypatch.delenv("GA_NORMAL_COLBLOCK", raising=False)
        U0, S0, V0 = pkg.svds(A, 3, ncv=12)
        assert np.array_equal(S, S0) and np.array_equal(V, V0) and np.array_equal(U, U0)


# ---------------------------------------------------------------- dsaitr
def _run_both(pkg, O, A, ncv, dtype, k_np_list):
    op = pkg.ArpackSimpleOp(A, dtype=dtype)
    ctx = pkg.Context(op, ncv)
    ora = O.Oracle(A, ncv, dtype=DT[dtype])
    assert ctx.getv0()[0] == 0 and ora.dgetv0() == 0
    rn = ctx.norm2_resid()
    es = (2,) if dtype == "dd" else ()
    H = np.zeros((ncv, 2) + es)
    out = []
    for (k, np_) in k_np_list:
        info, rn = ctx.lanczos(k, np_, H, rn)
        oinfo = ora.dsaitr(k, np_)
        out.append((info, oinfo))
    return ctx, ora, H, rn, out


@pytest.mark.parametrize("dtype", ["f64", "c64", "dd"])
def test_lanczos_parity_small(pkg, syn, O, dtype):
    """ga_lanczos vs the oracle's dsaitr! on the same start vector: H, V, resid, rnorm, DGKS counts.
    Tolerance: the two differ only in summation order, so 40 steps agree to ~1e-9 (f64)."""
    if dtype == "c64":
        A = syn.to_scipy(syn.hermitian_stencil(40))
    else:
        A = syn.to_scipy(syn.laplacian2d(48))
    ncv = 24
    ctx, ora, H, rn, infos = _run_both(pkg, O, A, ncv, dtype, [(0, 6), (6, 18)])
    assert infos == [(0, 0), (0, 0)]
    V = np.asarray(ctx.download_v())
    r = np.asarray(ctx.download_resid())
    Vo = np.swapaxes(np.asarray(ora.Vt), 0, 1)
    if dtype == "dd":
        Vf, Vof, rf, rof = _dd_float(V), _dd_float(Vo), _dd_float(r), _dd_float(np.asarray(ora.resid))
        Hf, Hof = _dd_float(H), _dd_float(np.asarray(ora.H))
        rnf, rnof = rn[0] + rn[1], ora.rnorm[0] + ora.rnorm[1]
        tol = 1e-20
        # double-double agreement well beyond double precision
        assert np.max(np.abs((H[..., 0] - ora.H[..., 0]) + (H[..., 1] - ora.H[..., 1]))) <= 1e-24 * np.max(np.abs(Hof))
    else:
        Vf, Vof, rf, rof, Hf, Hof, rnf, rnof = V, Vo, r, np.asarray(ora.resid), H, np.asarray(ora.H), rn, ora.rnorm
        tol = 1e-9
    assert np.allclose(Hf, Hof, rtol=tol, atol=tol * np.max(np.abs(Hof)))
    assert np.allclose(Vf, Vof, atol=max(tol, 1e-20) * 10)
    assert np.allclose(rf, rof, atol=10 * tol * np.max(np.abs(rof)))
    assert abs(rnf - rnof) <= tol * abs(rnof)
    # invariants of test/dsaitr_simple.jl:119-137
    G = Vf.conj().T @ Vf
    assert np.allclose(G, np.eye(ncv), atol=1e-13)
    assert np.linalg.norm(Vf.conj().T @ rf) <= 1e-13 * np.linalg.norm(rf)
    T = np.diag(Hf[:, 1]) + np.diag(Hf[1:, 0], 1) + np.diag(Hf[1:, 0], -1)
    R = A @ Vf - Vf @ T
    assert np.allclose(R[:, :-1], 0, atol=1e-12) and np.allclose(R[:, -1], rf, atol=1e-12)
    st, so = ctx.stats(), ora.stats()
    assert st["nopx"] == so["nopx"] == ncv
    assert st["nrorth"] == so["nrorth"] and st["nitref"] == so["nitref"] and st["nrstrt"] == 0
    ctx.close()


@pytest.mark.parametrize("dtype,ncv", [("f64", 44), ("c64", 44), ("dd", 44), ("f64", 60), ("c64", 60), ("dd", 60)])
def test_lanczos_wide_basis_tensor_map_path(pkg, syn, O, dtype, ncv):
    """Bases wider than 32 columns leave room for fewer than three 2 KiB-segment stages; the step kernel then works on
    tiles of half the rows, each moved by ONE tensor-map copy (cp.async.bulk.tensor, k_gs_step.cuh).  Same arithmetic:
    Float64 H / V bitwise equal to the oracle (exact dots), ComplexF64 / Double64 at their lock-step tolerances,
    V'V = I, and the three-term relation."""
    A = syn.to_scipy(syn.hermitian_stencil(40) if dtype == "c64" else syn.laplacian2d(48))
    ctx, ora, H, rn, infos = _run_both(pkg, O, A, ncv, dtype, [(0, 10), (10, ncv - 10)])
    assert infos == [(0, 0), (0, 0)]
    V = np.asarray(ctx.download_v())
    Vo = np.swapaxes(np.asarray(ora.Vt), 0, 1)
    if dtype == "f64":
        assert np.array_equal(H, np.asarray(ora.H)) and np.array_equal(V, Vo)
        Vf, Hf, rf = V, H, np.asarray(ctx.download_resid())
    elif dtype == "dd":
        Vf, Hf, rf = _dd_float(V), _dd_float(H), _dd_float(np.asarray(ctx.download_resid()))
        assert np.max(np.abs((H[..., 0] - ora.H[..., 0]) + (H[..., 1] - ora.H[..., 1]))) <= 1e-22 * np.max(np.abs(Hf))
    else:
        Vf, Hf, rf = V, H, np.asarray(ctx.download_resid())
        k = 24                                                  # plain complex accumulators: rounding differences grow with j
        assert np.allclose(Hf[:k], np.asarray(ora.H)[:k], rtol=1e-8, atol=1e-8 * np.max(np.abs(Hf)))
    assert np.allclose(Vf.conj().T @ Vf, np.eye(ncv), atol=1e-12)
    T = np.diag(Hf[:, 1]) + np.diag(Hf[1:, 0], 1) + np.diag(Hf[1:, 0], -1)
    R = A @ Vf - Vf @ T
    assert np.allclose(R[:, :-1], 0, atol=1e-11) and np.allclose(R[:, -1], rf, atol=1e-11)
    assert ctx.stats()["nopx"] == ora.stats()["nopx"] == ncv
    ctx.close()


def test_lanczos_identity_restart(pkg, O):
    """test/dsaitr_simple.jl:140-178: identity operator -> invariant subspace after one step ->
    dgetv0 restart inside the call; resid stays ~0; counters agree with the oracle.  The start
    vector is ones(16): v = 1/4 exactly and sum(v^2) = 1 exactly in ANY summation order, so the
    residual is exactly zero on both sides and the restart is taken deterministically."""
    n = 16
    A = sp.identity(n, format="csr")
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), 3)
    ora = O.Oracle(A, 3)
    x = np.ones(n)
    ctx.upload_resid(x)
    ora.resid[:] = x
    ora.rnorm = np.linalg.norm(x)
    H = np.zeros((3, 2))
    info, rn = ctx.lanczos(0, 2, H, np.linalg.norm(x))
    oinfo = ora.dsaitr(0, 2)
    assert info == oinfo == 0
    assert ctx.stats()["nrstrt"] == ora.stats()["nrstrt"] > 0
    assert ctx.stats()["nitref"] == ora.stats()["nitref"] > 0      # both DGKS refinement rounds are taken (:1424-1446)
    assert np.linalg.norm(ctx.download_resid()) <= n * np.finfo(float).eps
    V = ctx.download_v(0, 2)
    assert np.allclose(V.T @ V, np.eye(2), atol=1e-14)
    assert tuple(ctx.iseed) == tuple(ora.iseed)              # same dlarnv stream consumed
    assert np.allclose(V, ora.Vt[:2].T, atol=1e-14)
    assert H[1, 0] == 0.0                                    # H[j,1] = 0 after a restart (:1259-1263)
    ctx.close()


def test_lanczos_second_refinement_lockstep(pkg, O):
    """A nearly invariant subspace (diag(1, ..., 1, 1 + 1e-8)): the first DGKS correction does not pass the 0.717 test, so the
    second refinement round runs (src/dsaitr.jl:1424-1446) -- in the step kernel that is the on-demand sweep that forms
    the second round's dot products (k_gs_step.cuh phases 3 and 4).  H, V and the counters against the oracle."""
    n = 50
    d = np.ones(n)
    d[-1] = 1 + 1e-8
    A = sp.diags(d).tocsr()
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), 4)
    ora = O.Oracle(A, 4)
    assert ctx.getv0()[0] == 0 and ora.dgetv0() == 0
    H = np.zeros((4, 2))
    info, rn = ctx.lanczos(0, 4, H, ctx.norm2_resid())
    assert info == ora.dsaitr(0, 4) == 0
    st, so = ctx.stats(), ora.stats()
    assert st["nitref"] == so["nitref"] > 0 and st["nrorth"] == so["nrorth"] and st["nrstrt"] == so["nrstrt"]
    assert np.array_equal(H, np.asarray(ora.H))
    V = np.asarray(ctx.download_v())
    assert np.array_equal(V, np.swapaxes(np.asarray(ora.Vt), 0, 1))
    assert rn == ora.rnorm
    ctx.close()


def test_lanczos_zero_initial_resid(pkg, O):
    """test/dsaitr_simple.jl:180-212."""
    n = 10
    A = sp.diags(np.arange(1.0, n + 1)).tocsr()
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), 3)
    ctx.upload_resid(np.random.default_rng(0).standard_normal(n))
    H = np.zeros((3, 2))
    info, rn = ctx.lanczos(0, 3, H, 0.0)
    assert info == 0
    V = ctx.download_v()
    st = ctx.stats()
    assert st["nrstrt"] > 0 and st["nopx"] == 3
    assert np.allclose(V.T @ V, np.eye(3), atol=1e-14)
    assert np.linalg.norm(V.T @ ctx.download_resid()) <= 1e-14
    ctx.close()


def test_getv0_restart_failure(pkg):
    """test/dgetv0_simple.jl:149-196."""
    n = 5
    A = sp.diags(np.arange(1.0, n + 1)).tocsr()
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), 6 if False else 5)
    ctx.upload_v(np.eye(n)[:, :4], 0)
    x = np.zeros(n)
    x[:4] = [0.3, -0.2, 0.9, 0.5]
    ctx.upload_resid(x)                                      # lies entirely in span(V[:,1:4])
    ierr, rn = ctx.getv0(initv=True, j=5)
    assert ierr == -1 and rn == 0.0 and np.all(ctx.download_resid() == 0)
    ctx2 = pkg.Context(pkg.ArpackSimpleOp(A), 3)
    V = np.zeros((n, 1))
    V[:, 0] = 1 / np.sqrt(n)
    ctx2.upload_v(V, 0)
    ctx2.upload_resid(2 * np.random.default_rng(1).random(n) - 1)
    ierr, rn = ctx2.getv0(initv=True, j=2)
    r = ctx2.download_resid()
    assert ierr == 0 and abs(np.sum(r)) <= n * np.finfo(float).eps and rn > 0
    ctx.close()
    ctx2.close()


# ---------------------------------------------------------------- dsapps tail / dseupd tail
@pytest.mark.parametrize("dtype", ["f64", "c64", "dd"])
def test_apply_shifts_parity(pkg, syn, O, dtype):
    """device tail of dsapps! vs the oracle's dsapps! fed the same factorisation."""
    A = syn.to_scipy(syn.hermitian_stencil(30) if dtype == "c64" else syn.laplacian2d(30))
    ncv, kev = 20, 7
    np_ = ncv - kev
    ctx, ora, H, rn, infos = _run_both(pkg, O, A, ncv, dtype, [(0, ncv)])
    # make the device state identical to the oracle's before the shifts
    Vo = np.swapaxes(np.asarray(ora.Vt), 0, 1).copy()
    ctx.upload_v(pkg.as_dd(Vo) if dtype == "dd" else Vo, 0)
    ctx.upload_resid(pkg.as_dd(np.asarray(ora.resid).copy()) if dtype == "dd" else np.asarray(ora.resid).copy())
    Hd = np.asarray(ora.H)
    Hf = _dd_float(Hd) if dtype == "dd" else Hd
    T = np.diag(Hf[:, 1]) + np.diag(Hf[1:, 0], 1) + np.diag(Hf[1:, 0], -1)
    shifts = np.sort(np.linalg.eigvalsh(T))[:np_]            # unwanted = smallest (which = LM-like)
    sh = np.zeros((np_, 2)) if dtype == "dd" else shifts
    if dtype == "dd":
        sh[:, 0] = shifts
    ora.dsapps(kev, np_, sh)
    Q = np.asarray(ora.Q).copy()
    beta = np.asarray(ora.H)[kev, 0]
    rnorm = ctx.apply_shifts(kev, np_, Q, beta)
    V = np.asarray(ctx.download_v(0, kev + 1))
    r = np.asarray(ctx.download_resid())
    Vn = np.swapaxes(np.asarray(ora.Vt), 0, 1)[:, : kev + 1]
    rno = O.norm2(np.asarray(ora.resid), DT[dtype])
    if dtype == "dd":
        assert np.max(np.abs((V[..., 0] - Vn[..., 0]) + (V[..., 1] - Vn[..., 1]))) <= 1e-28
        assert np.max(np.abs((r[..., 0] - ora.resid[..., 0]) + (r[..., 1] - ora.resid[..., 1]))) <= 1e-28 * (1 + abs(rno[0]))
        assert abs((rnorm[0] - rno[0]) + (rnorm[1] - rno[1])) <= 1e-28 * rno[0]
    else:
        assert np.allclose(V, Vn, atol=1e-14)
        assert np.allclose(r, np.asarray(ora.resid), atol=1e-14 * (1 + abs(rno)))
        assert abs(rnorm - rno) <= 4 * np.spacing(rno)
    ctx.close()


def test_ritz_vectors_parity(pkg, syn):
    A = syn.to_scipy(syn.laplacian2d(20))
    n, ncv, nconv = A.shape[0], 12, 5
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), ncv)
    rng = np.random.default_rng(2)
    V0 = rng.standard_normal((n, ncv))
    Qh = np.linalg.qr(rng.standard_normal((ncv, ncv)))[0]
    ctx.upload_v(V0, 0)
    Z = ctx.ritz_vectors(nconv, Qh)
    assert np.allclose(Z, (V0 @ Qh)[:, :nconv], atol=1e-13)
    assert np.allclose(ctx.download_v(), V0 @ Qh, atol=1e-13)   # V <- V*Q like the reference (:1431)
    ctx.close()


# ---------------------------------------------------------------- whole solves (KATs)
def _top(w, k, which="LM"):
    if which == "LM":
        return np.sort(w[np.argsort(-np.abs(w))[:k]])
    if which == "LA":
        return np.sort(np.sort(w)[-k:])
    if which == "SA":
        return np.sort(np.sort(w)[:k])
    raise ValueError(which)


def _check_against_oracle(pkg, O, A, nev, ncv, which="LM", dtype="f64", v0=None, rtol=1e-10):
    res = pkg.eigs(pkg.ArpackSimpleOp(A, dtype=dtype), nev, ncv=ncv, which=which, v0=v0)
    ora = O.Oracle(A, ncv, dtype=DT[dtype]).symeigs(nev, which=which, v0=v0)
    vals, ovals = np.asarray(res.values), np.asarray(ora["values"])
    if dtype == "dd":
        d = np.abs((vals[:, 0] - ovals[:, 0]) + (vals[:, 1] - ovals[:, 1]))
        assert np.max(d / np.abs(ovals[:, 0])) <= 1e-26
    else:
        assert np.max(np.abs(vals - ovals) / np.abs(ovals)) <= rtol
    assert res.info == ora["info"] == 0
    assert abs(int(res.iparam[2]) - ora["niter"]) <= 1          # restart count within +-1
    assert int(res.iparam[4]) == ora["nconv"]
    return res, ora


def test_eigs_kat_15x15(pkg, O):
    """test/interface_simple.jl:7-11 and test/dsaupd_arpackjll.jl:232-241."""
    A = kat15()
    w = np.linalg.eigvalsh(A.toarray())
    for ncv in (15, 6):
        res, _ = _check_against_oracle(pkg, O, A, 3, ncv)
        assert np.allclose(res.values, _top(w, 3), rtol=1e-12)
        Z = res.vectors
        assert np.allclose(A @ Z, Z * res.values, atol=1e-11)
        assert np.allclose(Z.T @ Z, np.eye(3), atol=1e-12)
    res = pkg.eigs(A, 3, which="LA", ritzvec=False)
    assert np.allclose(res.values, _top(w, 3, "LA"), rtol=1e-12) and res.vectors is None
    res = pkg.eigs(A, 3, which="SA")
    assert np.allclose(res.values, _top(w, 3, "SA"), rtol=1e-12)
    res = pkg.eigs(A, 4, which="BE", ncv=10)
    assert np.allclose(res.values, np.sort(np.concatenate([np.sort(w)[:2], np.sort(w)[-2:]])), rtol=1e-12)


@pytest.mark.parametrize("n,nev,ncv", [(10, 3, 6), (30, 4, 10), (1000, 6, 12)])
def test_eigs_kat_diagonal(pkg, O, n, nev, ncv):
    A = sp.diags(np.arange(1.0, n + 1)).tocsr()
    v0 = np.ones(n) / np.sqrt(n)
    res, ora = _check_against_oracle(pkg, O, A, nev, ncv, v0=v0)
    assert np.allclose(res.values, np.arange(n - nev + 1, n + 1), rtol=1e-12)


def test_eigs_kat_issue5(pkg, O):
    """test/interface_simple.jl:15-24: a rank-4 matrix with nev = 5, ncv = 8.  The Krylov space of this matrix has
    dimension 2, so what happens after step 2 is decided by rounding noise alone: with the plain (BLAS-like) dot
    products the noise keeps the factorisation going and the reference's assertions hold; with error-free dot products
    the residual and the Ritz estimates vanish exactly, dsaup2 finds no shift to apply (src/dsaup2.jl:1177-1183,1293-1295)
    and dsaupd reports info = 3 -- on the device and in the oracle alike."""
    A = sp.lil_matrix((10, 10))
    for i in (0, 1, 4, 5):
        A[i, i] = 1.0
    A = A.tocsr()
    res = pkg.eigs(A, 5, ncv=8, dots="plain")
    Z, lam = res.vectors, res.values
    assert np.allclose(A @ Z, Z * lam, atol=1e-12)
    assert np.allclose(Z.T @ Z, np.eye(5), atol=1e-12)
    assert np.sum(lam) == pytest.approx(4.0, rel=1e-12)
    ora = O.Oracle(A, 8, dtype=O.F64).symeigs(5)          # exact arithmetic (module fixture)
    assert ora["info"] == 3
    with pytest.raises(pkg.ArpackException, match="info=3"):
        pkg.eigs(A, 5, ncv=8)


@pytest.mark.parametrize("m,ncv", [(48, 24), (300, 40)])
def test_exact_dots_lockstep_bitwise(pkg, syn, O, m, ncv):
    """dgetv0 + ncv Lanczos steps on the 2-D Laplacian, device (Dot2 sweeps, double-double norms) vs the oracle in its
    exact-arithmetic mode: H, every basis column, the residual and rnorm are the same bits, although the device splits
    the rows over 148 CTAs x 16 warps and the oracle sums them in order."""
    A = syn.to_scipy(syn.laplacian2d(m))
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), ncv)
    assert ctx.get_dots() == "exact"
    ora = O.Oracle(A, ncv)
    ctx.getv0(); ora.dgetv0()
    rn = ctx.norm2_resid()
    H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, ncv, H, rn)
    ora.dsaitr(0, ncv)
    V = np.asarray(ctx.download_v()); Vo = np.swapaxes(np.asarray(ora.Vt), 0, 1)
    assert info == 0
    assert np.array_equal(H, np.asarray(ora.H))
    assert all(np.array_equal(V[:, c], Vo[:, c]) for c in range(ncv))
    assert np.array_equal(np.asarray(ctx.download_resid()), np.asarray(ora.resid))
    assert rn == ora.rnorm
    assert ctx.stats()["nrorth"] == ora.stats()["nrorth"] and ctx.stats()["nitref"] == ora.stats()["nitref"]
    ctx.close()


@pytest.mark.parametrize("m,nev,ncv", [(64, 4, 20), (100, 4, 20), (100, 10, 40)])
def test_exact_dots_restart_counts_equal(pkg, syn, O, m, nev, ncv):
    """north_star: restart counts within +-1 of the reference on the same matrix and seed.  On the 2-D Laplacian the
    wanted eigenvalues are exactly double, which makes the count sensitive to the last bits of h (round 1: 55 vs 53,
    108 vs 105 with plain accumulators); with error-free dots on both sides the counts, the operator applications and
    the eigenvalues (bit for bit) are equal."""
    csr = syn.laplacian2d(m)
    res = pkg.eigs(pkg.ArpackSimpleOp(csr), nev, ncv=ncv)
    ora = O.Oracle(syn.to_scipy(csr), ncv).symeigs(nev)
    assert abs(int(res.iparam[2]) - ora["niter"]) <= 1
    assert int(res.iparam[2]) == ora["niter"] and res.stats["nopx"] == ora["stats"]["nopx"]
    assert np.array_equal(np.asarray(res.values), np.asarray(ora["values"]))
    exact = syn.laplacian2d_eigs(m, nev)
    assert np.max(np.abs(res.values - exact) / exact) <= 1e-10


def test_plain_dots_option(pkg, syn, O):
    """GA_OPT_DOT_MODE = GA_DOT_PLAIN: one FMA per product (BLAS-like).  Same eigenvalues to 1e-10, H within a few ulps
    of the exact-dots H over one pass, and the option is per context."""
    m, ncv = 64, 20
    A = syn.to_scipy(syn.laplacian2d(m))
    Hs = {}
    for mode in ("exact", "plain"):
        ctx = pkg.Context(pkg.ArpackSimpleOp(A), ncv, dots=mode)
        assert ctx.get_dots() == mode
        ctx.getv0()
        rn = ctx.norm2_resid()
        H = np.zeros((ncv, 2))
        info, rn = ctx.lanczos(0, ncv, H, rn)
        assert info == 0
        Hs[mode] = H.copy()
        ctx.close()
    assert np.allclose(Hs["exact"], Hs["plain"], rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        pkg.Context(pkg.ArpackSimpleOp(A), ncv, dots="fast")
    res = pkg.eigs(pkg.ArpackSimpleOp(A), 4, ncv=ncv, dots="plain")
    exact = syn.laplacian2d_eigs(m, 4)
    assert np.max(np.abs(res.values - exact) / exact) <= 1e-10


def test_normalize_tiny_rnorm_dlascl_branch(pkg, syn, O):
    """src/dsaitr.jl:1097-1105: v_j = resid / rnorm is a multiplication by 1/rnorm only while rnorm >= floatmin; below
    that the reference scales like dlascl (_scale_from_to(rnorm, 1, v), src/arpack-blas.jl:25-76) because 1/rnorm would
    overflow.  A start vector of subnormal entries takes that branch on the device (k_vec.cu) and in the oracle.  The
    oracle runs with the reference's x86 norm here (80-bit accumulators, src/arpack-blas-nrm2.jl:272-299): its
    double-double norm (:157-215, the non-x86 path) squares in binary64 and returns 0 for such a vector, so only an
    x86 run of the reference can reach this branch at all; the device norm (scaled three-range accumulation) agrees
    with the x87 value."""
    m, ncv = 16, 8
    A = syn.to_scipy(syn.laplacian2d(m))
    n = m * m
    x = np.asarray(O.dlarnv(n)[0]) * 1e-310                          # subnormal entries
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), ncv)
    ctx.upload_resid(x)
    rn = ctx.norm2_resid()
    assert 0.0 < rn < np.finfo(float).tiny
    with O.arith("plain"):
        ora = O.Oracle(A, ncv)
        ora.resid[:] = x
        assert rn == O.norm2(x)
        ora.rnorm = rn
        assert ora.dsaitr(0, 1) == 0
        vo = np.asarray(ora.Vt)[0].copy()
        Ho = np.asarray(ora.H)[:1].copy()
        rno = ora.rnorm
    H = np.zeros((ncv, 2))
    info, rn1 = ctx.lanczos(0, 1, H, rn)
    assert info == 0
    v = np.asarray(ctx.download_v(0, 1))[:, 0]
    assert np.array_equal(v, vo)                                     # element-wise scaling: the same bits
    assert abs(np.linalg.norm(v) - 1.0) <= 1e-12                     # relative precision of subnormal inputs ~1e-13
    assert np.allclose(H[:1], Ho, rtol=1e-13) and abs(rn1 - rno) <= 1e-13 * rno
    ctx.close()


def test_eigs_laplacian_analytic(pkg, syn, O):
    """C2 at reduced size: eigs(A, 10; ncv=40) on the 2-D Laplacian vs the analytic spectrum, the
    oracle, Ritz estimates (bounds <= tol*|lambda|) and true residuals."""
    m = 120
    csr = syn.laplacian2d(m)
    A = syn.to_scipy(csr)
    res, ora = _check_against_oracle(pkg, O, A, 10, 40)
    exact = syn.laplacian2d_eigs(m, 10)
    assert np.max(np.abs(res.values - exact) / exact) <= 1e-10
    tol = np.finfo(float).eps / 2
    assert np.all(res.bounds <= tol * np.maximum(np.finfo(float).eps ** (2 / 3), np.abs(res.values)) * (1 + 1e-12))
    Z = res.vectors
    R = A @ Z - Z * res.values
    assert np.max(np.linalg.norm(R, axis=0) / np.abs(res.values)) <= 1e-10
    assert np.allclose(Z.T @ Z, np.eye(10), atol=1e-12)
    assert abs(res.stats["nopx"] - ora["stats"]["nopx"]) <= 40


def test_eigs_sprand_c1(pkg, syn, O):
    """C1 at BASELINE size (README example shape): sprand-symmetric n = 100 000, nev = 2, ncv = 12, which = :LM.
    Oracle parity: eigenvalues to 1e-10, identical restart and operator-application counts."""
    A = syn.to_scipy(syn.sprand_sym(100000, 5.0, seed=1234))
    res, ora = _check_against_oracle(pkg, O, A, 2, 12)
    assert int(res.iparam[2]) == ora["niter"] and res.stats["nopx"] == ora["stats"]["nopx"]
    Z = res.vectors
    assert np.max(np.linalg.norm(A @ Z - Z * res.values, axis=0) / np.abs(res.values)) <= 1e-10
    assert np.allclose(Z.T @ Z, np.eye(2), atol=1e-12)


def test_eigs_sprand_c3_double64(pkg, syn, O):
    """C3 at BASELINE size: the C1 matrix in Double64 (double-double kernels), nev = 2, ncv = 12: eigenvalues to 1e-26
    of the oracle's Double64 path, restart counts within +-1, Ritz residuals at the Double64 level."""
    csr = syn.sprand_sym(100000, 5.0, seed=1234)
    A = syn.to_scipy(csr)
    res, ora = _check_against_oracle(pkg, O, A, 2, 12, dtype="dd")
    assert abs(res.stats["nopx"] - ora["stats"]["nopx"]) <= 12
    lam = np.asarray(res.values)
    Z = np.asarray(res.vectors)                                    # (n, 2, (hi, lo))
    for i in range(2):
        z = Z[:, i, 0] + Z[:, i, 1]
        r = A @ Z[:, i, 0] - lam[i, 0] * Z[:, i, 0]                # binary64 evaluation of the residual: bounded by eps64 |lambda|
        assert np.linalg.norm(r) <= 64 * np.finfo(float).eps * abs(lam[i, 0])
        assert abs(np.dot(z, z) - 1.0) <= 1e-14


def test_eigs_complex_hermitian(pkg, syn, O):
    """C4 at reduced size + test/dsaupd_complex.jl:78-80."""
    n = 20
    H = (sp.diags([1j * np.ones(n - 1), 2 * np.ones(n) + 0j, -1j * np.ones(n - 1)], [-1, 0, 1]) * (n + 1)).tocsr()
    w = np.linalg.eigvalsh(H.toarray())
    res, _ = _check_against_oracle(pkg, O, H, 4, 12, dtype="c64")
    assert np.allclose(res.values, _top(w, 4), rtol=1e-12)
    A = syn.to_scipy(syn.hermitian_stencil(60))
    res, _ = _check_against_oracle(pkg, O, A, 6, 30, dtype="c64")
    Z = res.vectors
    assert np.allclose(Z.conj().T @ Z, np.eye(6), atol=1e-12)
    assert np.max(np.linalg.norm(A @ Z - Z * res.values, axis=0) / np.abs(res.values)) <= 1e-10


def test_eigs_double64(pkg, syn, O):
    """C3 + test/interface_high.jl:60-66: Double64 eigenvalues to 1e-26 (oracle and exact)."""
    import mpmath as mp

    mp.mp.prec = 250
    n = 100
    A = (sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * (n + 1)).tocsr()
    res, ora = _check_against_oracle(pkg, O, A, 4, 20, dtype="dd")
    exact = [(n + 1) * (2 - 2 * mp.cos(k * mp.pi / (n + 1))) for k in range(n - 3, n + 1)]
    got = [mp.mpf(float(v[0])) + mp.mpf(float(v[1])) for v in np.asarray(res.values)]
    assert max(abs(g - e) / e for g, e in zip(got, exact)) < mp.mpf(10) ** -26
    A2 = syn.to_scipy(syn.sprand_sym(3000, 5.0, seed=1234))
    res2, ora2 = _check_against_oracle(pkg, O, A2, 2, 12, dtype="dd")
    # the Double64 run refines the Float64 one
    f = pkg.eigs(A2, 2, ncv=12)
    assert np.allclose(_dd_float(np.asarray(res2.values)), f.values, rtol=1e-12)


def _mytestmat(m, n, minval=None, maxdiagval=2.0):
    """test/utility.jl:259-265"""
    import scipy.sparse as sp

    minval = 1.0 / m if minval is None else minval
    v = np.linspace(1.0, minval, m).astype(np.result_type(minval, np.float64))
    B = sp.coo_matrix((np.linspace(1.0, maxdiagval, n - 1), (np.arange(n - 1), np.arange(n - 1))), shape=(m, n - 1))
    return sp.hstack([sp.csr_matrix(v.reshape(-1, 1)), B]).tocsr()


def _check_svd(A, U, s, V, tol):
    """check_svd (test/utility.jl:244-253)"""
    k = len(s)
    assert np.allclose(U.conj().T @ U, np.eye(k), atol=1e-12)
    assert np.allclose(V.conj().T @ V, np.eye(k), atol=1e-12)
    r = [np.linalg.norm(A @ V[:, i] - U[:, i] * s[i]) for i in range(k)]
    assert max(r) <= np.finfo(np.float64).eps * tol, r


def test_svds_normal(pkg, syn):
    """C5 at reduced size: svds via the normal operator (src/interface.jl:304-337), U and V from the
    device (ga_svd_vectors) against the CPU oracle's Householder orth! and a dense SVD."""
    from oracle import oracle_svd as OS

    A = syn.to_scipy(syn.tall_sparse(4000, 200, 8))
    U, S, V = pkg.svds(A, 5, ncv=20)
    s = np.linalg.svd(A.toarray(), compute_uv=False)
    assert np.allclose(S, s[:5], rtol=1e-9)
    _check_svd(A, U, S, V, tol=200)
    Uo, So, Vo, _ = OS.svds(A, 5, ncv=20)
    assert np.allclose(S, So, rtol=1e-10)
    # same unique factors up to the sign of each eigenvector pair (Lanczos start identical, so signs match too)
    for i in range(5):
        sg = np.sign(np.vdot(Vo[:, i], V[:, i]))
        assert np.allclose(V[:, i] * sg, Vo[:, i], atol=1e-8)
        assert np.allclose(U[:, i] * sg, Uo[:, i], atol=1e-8)


def test_svds_golden_values(pkg):
    """the reference's own svds known answers (test/interface_simple.jl:56-63,100-105; interface_complex.jl:30-43)"""
    import scipy.sparse as sp

    A = _mytestmat(10, 8)
    U, s, V = pkg.svds(A, 2, which="BE")
    assert np.allclose(s, GOLD["svds_mytestmat_10_8_BE"]["values"], rtol=1e-12)
    _check_svd(A, U, s, V, tol=75)
    U, s, V = pkg.svds(A.T.tocsr(), 2, which="BE")             # m <= n: the AA' side
    assert np.allclose(s, GOLD["svds_mytestmat_10_8_BE"]["values"], rtol=1e-12)
    _check_svd(A.T.tocsr(), U, s, V, tol=75)
    A1 = sp.csr_matrix(np.ones((5, 5)))
    U, s, V = pkg.svds(A1, 1)
    assert np.allclose(s, [5.0])
    _check_svd(A1, U, s, V, tol=10)
    Ac = _mytestmat(10, 8, -3.0j)
    U, s, V = pkg.svds(Ac, 2, which="BE")
    assert np.allclose(s, GOLD["complex_svds_BE"]["values"], rtol=1e-12)
    _check_svd(Ac, U, s, V, tol=60)
    U, s, V = pkg.svds(Ac, 2, which="LM")
    assert np.allclose(s, GOLD["complex_svds_LM"]["values"], rtol=1e-12)
    _check_svd(Ac, U, s, V, tol=60)
    U, s, V = pkg.svds(A, 2, which="SA", ritzvec=False)
    assert U.shape[1] == 0 and V.shape[1] == 0 and len(s) == 2


def test_svds_singular_case(pkg):
    """test/interface_simple.jl:83-98 on the device: a rank-deficient tall matrix, smallest singular values.
    svds(A, 1; which=:SA, ncv=2) must raise (nev < ncv leaves no room: the reference's ArpackException), the ncv = 20
    solve either meets sigma ~ 0 while forming U (ArpackException, src/idonow_ops.jl:300-303) or returns residuals within
    n sqrt(eps), and svds(A, 2; which=:SA, ritzvec=false) returns [0, 0] with empty U, V."""
    import scipy.sparse as sp

    n = 80
    A = sp.vstack([sp.diags(np.arange(1.0, n + 1)), sp.csr_matrix((20, n))]).tolil()
    A[n - 5:n, n - 5:n] = 0
    A = A.tocsr()
    with pytest.raises((pkg.ArpackException, ValueError)):
        pkg.svds(A, 1, which="SA", ncv=2)
    try:
        U, s, V = pkg.svds(A, 1, which="SA", ncv=20)
    except pkg.ArpackException as e:
        assert "sigma" in str(e)
    else:
        assert np.max(pkg.svd_residuals(A, U, s, V)) <= n * np.sqrt(np.finfo(float).eps)
    U, s, V = pkg.svds(A, 2, which="SA", ritzvec=False)
    assert np.allclose(s, [0.0, 0.0], atol=2 * n * np.sqrt(np.finfo(float).eps))
    assert U.shape[1] == 0 and V.shape[1] == 0


def test_svds_arpack_example(pkg):
    """ARPACK's dsvd example operator as a matrix (test/interface_simple.jl:160-222)"""
    import scipy.sparse as sp

    m, n = 500, 100
    h, k = 1.0 / (m + 1), 1.0 / (n + 1)
    A = np.zeros((m, n))
    for j in range(n):
        t = (j + 1) * k
        sv = (np.arange(m) + 1) * h
        A[: j + 1, j] = k * sv[: j + 1] * (t - 1)
        A[j + 1:, j] = k * t * (sv[j + 1:] - 1)
    U, s, V = pkg.svds(sp.csr_matrix(A), 4, ncv=10)
    gold = sorted(GOLD["svds_arpack_example_500x100"]["values"], reverse=True)
    assert np.allclose(s, gold, rtol=1e-10)
    _check_svd(A, U, s, V, tol=4)


def test_maxiter_exception(pkg, syn):
    """src/interface.jl:269-270: info == 1 with failonmaxiter raises ArpackException."""
    A = syn.to_scipy(syn.laplacian2d(60))
    with pytest.raises(pkg.ArpackException):
        pkg.eigs(A, 6, ncv=14, maxiter=2)
    res = pkg.eigs(A, 6, ncv=14, maxiter=2, failonmaxiter=False)
    assert res.info == 1 and int(res.iparam[2]) == 3 and len(res.values) == int(res.iparam[4]) < 6


# ---------------------------------------------------------------- full size (BASELINE config C2)
def test_full_size_properties(pkg, syn):
    """n = 16M 2-D Laplacian, ncv = 40: two restart cycles on the device, then size-independent
    checks on downloaded columns: unit norms, orthogonality of sampled pairs, and the Lanczos
    three-term relation A v_j = b_{j-1} v_{j-1} + a_j v_j + b_j v_{j+1} for sampled j."""
    m = 4000
    csr = syn.laplacian2d(m)
    A = syn.to_scipy(csr)
    ncv, nev = 40, 10
    ctx = pkg.Context(pkg.ArpackSimpleOp(csr), ncv)
    assert ctx.getv0()[0] == 0
    rn = ctx.norm2_resid()
    H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, ncv, H, rn)
    assert info == 0 and rn > 0
    for j in (1, 17, 38):
        Vj = ctx.download_v(j - 1, 3)
        nrm = np.linalg.norm(Vj, axis=0)
        assert np.allclose(nrm, 1.0, atol=1e-13)
        G = Vj.T @ Vj
        assert np.allclose(G, np.eye(3), atol=1e-12)
        lhs = A @ Vj[:, 1]
        rhs = H[j, 0] * Vj[:, 0] + H[j, 1] * Vj[:, 1] + H[j + 1, 0] * Vj[:, 2]
        assert np.linalg.norm(lhs - rhs) <= 1e-12 * np.linalg.norm(lhs)
    v_first, v_last = ctx.download_v(0, 1)[:, 0], ctx.download_v(ncv - 1, 1)[:, 0]
    assert abs(v_first @ v_last) <= 1e-12
    r = ctx.download_resid()
    assert abs(np.linalg.norm(r) - rn) <= 1e-13 * rn
    assert abs(v_first @ r) <= 1e-12 * rn and abs(v_last @ r) <= 1e-12 * rn
    st = ctx.stats()
    assert st["nopx"] == ncv
    # Ritz values of the 40-step factorisation lie inside the analytic spectrum (0, 8)
    T = np.diag(H[:, 1]) + np.diag(H[1:, 0], 1) + np.diag(H[1:, 0], -1)
    w = np.linalg.eigvalsh(T)
    assert w.min() > 0 and w.max() < 8
    ctx.close()


def test_full_size_complex_c4(pkg, syn):
    """BASELINE config C4 at full size (complex Hermitian stencil, n = 4M, nev = 6, ncv = 30): the whole
    solve, checked through properties only -- true Ritz residuals ||A z - lambda z|| <= 1e-10 |lambda|
    against SciPy's SpMV, orthonormal Ritz vectors, and a solve with the first SpMV kernel giving the
    same restart count and eigenvalues (the streamed kernel's configurations are bit-identical)."""
    m = 2000
    csr = syn.hermitian_stencil(m)
    A = syn.to_scipy(csr)
    res = []
    for first_kernel in (False, True):
        ctx = pkg.Context(pkg.ArpackSimpleOp(csr, dtype="c64"), 30)
        if first_kernel:
            ctx.tune_spmv(0, 0)
        assert ctx.symeigs_begin(6, "LM", 0.0, 300, None) == 0
        ctx.symeigs_iterate(-1)
        rc, vals, bounds, Z, iparam = ctx.symeigs_end(6, True)
        assert rc == 0 and len(vals) == 6
        res.append((vals.copy(), int(iparam[2]), ctx.stats()["nopx"]))
        if not first_kernel:
            for i in range(6):
                r = A @ Z[:, i] - vals[i] * Z[:, i]
                assert np.linalg.norm(r) <= 1e-10 * abs(vals[i])
            G = Z.conj().T @ Z
            assert np.allclose(G, np.eye(6), atol=1e-12)
        ctx.close()
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]
    assert np.array_equal(res[0][0], res[1][0])


def test_large_ncv_fallback(pkg, syn, O):
    """ncv up to GA_MAX_NCV = 64 (config C5 uses ncv = 60): the sweep kernel falls back from 2 KiB
    to 1 KiB TMA segments when fewer than two pipeline stages would fit."""
    A = syn.to_scipy(syn.laplacian2d(70))
    res, ora = _check_against_oracle(pkg, O, A, 20, 64)
    exact = syn.laplacian2d_eigs(70, 20)
    assert np.max(np.abs(res.values - exact) / exact) <= 1e-10
    Z = res.vectors
    assert np.allclose(Z.T @ Z, np.eye(20), atol=1e-12)
    with pytest.raises(Exception):
        pkg.eigs(A, 20, ncv=65)          # beyond GA_MAX_NCV: refused, never silently degraded


def test_dsaupd_error_codes(pkg):
    """test/dsaupd_errors.jl: the info codes of dsaupd!'s argument checks (src/dsaupd.jl:963-1002) and the zero
    start vector (-9, src/dsaup2.jl:1071-1076), through the C ABI; the interface level wraps them in
    ArpackException (src/interface.jl:269-273)."""
    n = 10
    A = sp.diags(np.arange(1.0, n + 1)).tocsr()
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), n - 1)
    assert ctx.symeigs_begin(6, v0=np.zeros(n)) == 0
    assert ctx.symeigs_iterate(-1) == 99
    assert ctx.symeigs_end(6)[0] == -9                      # "initial vector is zero"
    assert ctx.symeigs_begin(0) == -2                       # "nev must be positive"
    assert ctx.symeigs_begin(n - 1) == -3                   # ncv must be greater than nev
    assert ctx.symeigs_begin(2, maxiter=-1) == -4
    assert ctx.symeigs_begin(2, maxiter=0) == 0             # mxiter = 0 is allowed (one iteration)
    assert ctx.symeigs_begin(2, which="SS") == -5
    assert ctx.symeigs_begin(1, which="BE") == -13
    assert ctx.symeigs_begin(0, which="SS") == -5           # a later check overrides an earlier one
    ctx.close()
    with pytest.raises(pkg.ArpackException):
        pkg.eigs(A, 2, which="SS")
    with pytest.raises(pkg.ArpackException):
        pkg.eigs(A, 6, ncv=9, v0=np.zeros(n))
    with pytest.raises(ValueError):                         # src/interface.jl:231-232
        pkg.eigs(A, 4, ncv=11)
    with pytest.raises(ValueError):
        pkg.eigs(A, 4, ncv=4)
