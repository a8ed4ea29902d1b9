"""GPU parity tests for Float32 vectors with a Float64 factorisation -- the reference's mixed precision
symeigs(Float32, Float64, op, nev) / eigs(Float32, A, k) (src/interface.jl:26-34, README.md:28-35).

Bars: the dlarnv start vector is bit-exact (binary32 conversion incl. the rv == 1 reseed branch,
src/arpack-blas.jl:520-525) against the oracle's sequential restatement; everything floating point is
compared at Float32 tolerances, which is what the reference's own Float32 tests do
(`vals ≈ Float32.(...)`, test/interface_simple.jl:42-49; svds(Float32, op, 4) golden to 6 digits, :221-222)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD, kat15

pytestmark = pytest.mark.gpu

LCG_A, M48 = 0x1EE1429CC9F5, 1 << 48


def seed_hitting_one(g, t=777):
    """a dlarnv seed whose g-th draw converts to exactly 1.0f in binary32 (X_g = 2^48 - 1 - 2t)"""
    X = M48 - 1 - 2 * t
    s = (X * pow(pow(LCG_A, g, M48), -1, M48)) % M48
    return ((s >> 36) & 4095, (s >> 24) & 4095, (s >> 12) & 4095, s & 4095)


@pytest.mark.parametrize("n", [10, 129, 4097, 100003])
def test_f32_start_vector_bit_exact(pkg, O, n):
    ctx = pkg.Context(pkg.ArpackSimpleOp(sp.identity(n, format="csr", dtype=np.float32)), 2)
    assert ctx.dtype == pkg.GA_F32
    ierr, rnorm = ctx.getv0(initv=False, j=1)
    x = np.asarray(ctx.download_resid())
    ref, seed, nres = O.dlarnv_f32(n)
    assert ierr == 0 and x.dtype == np.float32
    assert x.tobytes() == ref.tobytes()
    assert tuple(int(s) for s in ctx.iseed) == seed
    exact = np.linalg.norm(ref.astype(np.float64))
    assert abs(rnorm - exact) <= 1e-14 * exact              # norm accumulated in binary64 (norm2(Float64, ::Vector{Float32}))
    ctx.close()


@pytest.mark.parametrize("g,n", [(1, 300), (5, 300), (128, 1000), (129, 1000), (4000, 10000), (70001, 100003)])
def test_f32_start_vector_reseed_branch(pkg, O, g, n):
    """seeds crafted so that draw g rounds to 1.0f: the reference bumps every limb of the block's base seed by 2 and
    redraws; the rest of the stream and the seed left behind change (src/arpack-blas.jl:520-525)."""
    sd = seed_hitting_one(g)
    ref, seed, nres = O.dlarnv_f32(n, sd)
    assert nres >= 1 and np.max(ref) < 1.0
    ctx = pkg.Context(pkg.ArpackSimpleOp(sp.identity(n, format="csr", dtype=np.float32)), 2)
    ctx.iseed[:] = sd
    assert ctx.getv0()[0] == 0
    x = np.asarray(ctx.download_resid())
    assert x.tobytes() == ref.tobytes()
    assert tuple(int(s) for s in ctx.iseed) == seed
    ctx.close()


def test_f32_opx(pkg, syn):
    rng = np.random.default_rng(1)
    for csr in (syn.laplacian2d(70), syn.sprand_sym(20000, 5.0, seed=7)):
        A = syn.to_scipy(csr)
        Af = A.astype(np.float32)
        ctx = pkg.Context(pkg.ArpackSimpleOp(Af), 4)
        x = rng.standard_normal(A.shape[0]).astype(np.float32)
        y = np.asarray(ctx.opx(x))
        ref = Af.astype(np.float64) @ x.astype(np.float64)
        assert y.dtype == np.float32
        assert np.linalg.norm(y - ref) <= 1e-6 * np.linalg.norm(ref)
        ctx.close()


@pytest.mark.parametrize("ncv", [12, 40, 56])
def test_f32_lanczos_properties(pkg, syn, O, ncv):
    """ncv Lanczos steps in Float32 on the 2-D Laplacian (integer entries: A is exact in Float32): H against the
    Float64 oracle from the same (bit-identical, widened) start vector, orthonormality of V, the three-term
    relation and the residual norm, at Float32 tolerances.  ncv = 56 takes the 1 KiB-segment sweep path."""
    m = 90
    A = syn.to_scipy(syn.laplacian2d(m))
    n = m * m
    ctx = pkg.Context(pkg.ArpackSimpleOp(A.astype(np.float32)), ncv)
    rc, rn = ctx.getv0()
    v0 = np.asarray(ctx.download_resid()).astype(np.float64)
    H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, ncv, H, rn)
    assert info == 0 and rn > 0
    o = O.Oracle(A, ncv)
    o.resid[:] = v0
    o.dgetv0(1, True, 1)
    o.dsaitr(0, ncv)
    Ho = np.asarray(o.H)
    k = min(ncv, 12)                                        # Lanczos amplifies rounding differences with j
    assert np.max(np.abs(H[:k] - Ho[:k])) <= 2e-5 * np.max(np.abs(Ho))
    V = np.asarray(ctx.download_v()).astype(np.float64)
    assert np.max(np.abs(V.T @ V - np.eye(ncv))) <= 2e-5
    r = np.asarray(ctx.download_resid()).astype(np.float64)
    assert abs(np.linalg.norm(r) - rn) <= 1e-6 * rn
    assert np.max(np.abs(V.T @ r)) <= 2e-5 * rn
    T = np.diag(H[:, 1]) + np.diag(H[1:, 0], 1) + np.diag(H[1:, 0], -1)
    E = A @ V - V @ T
    E[:, -1] -= r
    assert np.max(np.linalg.norm(E, axis=0)) <= 2e-5 * 8.0   # A V = V T + r e_k'
    st = ctx.stats()
    assert st["nopx"] == ncv
    ctx.close()


def test_f32_lockstep_against_f32_oracle(pkg, syn, O):
    """The Float32 device path in lock step with the oracle's own Float32 instantiation (oracle.F32: Float32 vectors,
    binary32 dlarnv incl. reseed, Float64 factorisation): bit-identical start vector and seed, the same DGKS counters,
    H to Float32 accuracy over the first steps (the device accumulates the dots in binary64, the reference's sgemv in
    binary32, so later steps drift at the 1e-7-per-step level), and the same eigenvalues from a whole solve."""
    m, ncv = 60, 20
    A = syn.to_scipy(syn.laplacian2d(m)).astype(np.float32)
    ctx = pkg.Context(pkg.ArpackSimpleOp(A), ncv)
    ora = O.Oracle(A, ncv, dtype=O.F32)
    assert ctx.getv0()[0] == 0 and ora.dgetv0() == 0
    assert np.array_equal(np.asarray(ctx.download_resid()), np.asarray(ora.resid))
    assert tuple(ctx.iseed) == tuple(int(v) for v in ora.iseed)
    rn = ctx.norm2_resid()
    assert abs(rn - ora.rnorm) <= 1e-14 * rn
    H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, ncv, H, rn)
    assert info == 0 and ora.dsaitr(0, ncv) == 0
    Ho = np.asarray(ora.H)
    assert np.max(np.abs(H[:10] - Ho[:10])) <= 2e-5 * np.max(np.abs(Ho))
    assert ctx.stats()["nopx"] == ora.stats()["nopx"] == ncv
    assert abs(ctx.stats()["nrorth"] - ora.stats()["nrorth"]) <= 2
    ctx.close()
    res = pkg.eigs(np.float32, np.float64, A, 4, ncv=ncv)
    o = O.Oracle(A, ncv, dtype=O.F32).symeigs(4, tol=float(np.finfo(np.float32).eps) / 2)
    assert o["info"] == 0 and np.allclose(res.values, np.asarray(o["values"]), rtol=1e-5)


def test_f32_eigs_kats(pkg):
    """test/interface_simple.jl:36-49 (15 x 15 KAT, which = :LA, Float32 and Float32/Float64) and the docstring
    example symeigs(Float32, ArpackSimpleOp(Diagonal(1.0:n)), 5; which=:SA) (src/interface.jl:105-110)."""
    A = kat15()
    w = np.linalg.eigvalsh(A.toarray())
    for lead in ((np.float32,), (np.float32, np.float64), ("f32",)):
        res = pkg.eigs(*lead, A, 3, which="LA", ritzvec=False)
        assert res.vectors is None
        assert res.values.dtype == (np.float64 if len(lead) == 2 else np.float32)     # eltype(vals), test/interface_simple.jl:43
        assert np.allclose(res.values, np.sort(w)[-3:], rtol=3e-6)
    res = pkg.eigs(np.float32, A, 3)
    assert res.vectors.dtype == np.float32
    Z = res.vectors.astype(np.float64)
    assert np.allclose(res.values, np.sort(w[np.argsort(-np.abs(w))[:3]]), rtol=3e-6)
    assert np.max(np.abs(A @ Z - Z * res.values)) <= 2e-5
    assert np.max(np.abs(Z.T @ Z - np.eye(3))) <= 1e-5
    n = 100
    D = sp.diags(np.arange(1.0, n + 1)).tocsr()
    res = pkg.symeigs(np.float32, pkg.ArpackSimpleOp(D, dtype="f32"), 5, which="SA")
    assert np.allclose(res.values, np.arange(1.0, 6.0), rtol=1e-5)
    with pytest.raises(ValueError):
        pkg.eigs(np.float64, np.float32, A, 3)              # TF narrower than TV is not offered
    with pytest.raises(Exception):
        pkg.eigs(np.float32, A, sp.identity(15, format="csr"), 3)   # generalised problem: Float64/ComplexF64/Double64 only


def test_f32_eigs_laplacian(pkg, syn):
    """C2's operator at reduced size in Float32: default tolerance eps(Float32)/2 (the lowest precision sets it,
    src/interface.jl:33-34), eigenvalues against the analytic spectrum, true residuals at Float32 level."""
    m = 120
    A = syn.to_scipy(syn.laplacian2d(m))
    st = {}
    res = pkg.eigs(np.float32, np.float64, A, 10, ncv=40, stats=st)
    exact = syn.laplacian2d_eigs(m, 10)
    assert res.info == 0 and len(res.values) == 10
    assert np.max(np.abs(res.values - exact) / exact) <= 1e-5
    Z = res.vectors.astype(np.float64)
    R = A @ Z - Z * res.values
    assert np.max(np.linalg.norm(R, axis=0) / np.abs(res.values)) <= 2e-5
    assert np.max(np.abs(Z.T @ Z - np.eye(10))) <= 2e-5
    assert 0 < int(res.iparam[2]) < 300


def test_f32_svds(pkg):
    """svds(Float32, op, 4; ncv=10) on ARPACK's dsvd example: the reference's 6-digit golden values
    (test/interface_simple.jl:221-222)."""
    m, n = 500, 100
    h, k = 1.0 / (m + 1), 1.0 / (n + 1)
    A = np.zeros((m, n))
    for j in range(n):
        t = (j + 1) * k
        sv = (np.arange(m) + 1) * h
        A[: j + 1, j] = k * sv[: j + 1] * (t - 1)
        A[j + 1:, j] = k * t * (sv[j + 1:] - 1)
    U, s, V = pkg.svds(sp.csr_matrix(A.astype(np.float32)), 4, ncv=10)
    assert U.dtype == np.float32 and V.dtype == np.float32
    assert np.allclose(s, sorted(GOLD["svds_arpack_example_500x100_float32"]["values"], reverse=True), rtol=2e-4)
    V = V.astype(np.float64)
    U = U.astype(np.float64)
    assert np.max(np.abs(U.T @ U - np.eye(4))) <= 1e-4 and np.max(np.abs(V.T @ V - np.eye(4))) <= 1e-4
    assert np.max(np.abs(A @ V - U * s)) <= 1e-4


def test_f32_refusals(pkg):
    """what Float32 contexts do not offer fails loudly (no silent widening, no fallback)."""
    A = sp.identity(40, format="csr", dtype=np.float32)
    with pytest.raises(Exception):
        pkg.Context(pkg.ArpackSimpleOp(A), 4, rank=0, world_size=2, n_global=80, row_offset=0)
