"""CPU tests that PIN the oracle (oracle/) against the reference's own golden vectors, known
answer tests and invariants.  Citations: /root/reference/test/*.jl."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import DSDRV3_GOLDEN, GOLD, GOLDEN_V0, dsdrv3 as _dsdrv3, kat15


# ---------------------------------------------------------------- LCG / start vector (bit-exact)
def test_golden_start_vector(O):
    """test/dgetv0_simple.jl:6-35: resid == v0 for seed (1,3,5,7), n = 10; seed must change."""
    x, seed = O.dlarnv(10)
    assert np.array_equal(x, np.array(GOLDEN_V0))
    assert seed != (1, 3, 5, 7)
    o = O.Oracle(sp.diags(np.arange(1.0, 11.0)).tocsr(), 2)
    assert o.dgetv0(itry=1, initv=False, j=1) == 0
    assert np.array_equal(o.resid, np.array(GOLDEN_V0))
    assert tuple(o.iseed) != (1, 3, 5, 7)
    assert o.rnorm == pytest.approx(np.linalg.norm(GOLDEN_V0), rel=1e-15)


def test_golden_dlarnv_stream_fixture(O):
    """tests/golden/reference_kats.json:dlarnv_stream (closed form pinned by the golden start vector): last three
    values (bit patterns) and the seed left behind for n = 1, 10, 127, 128, 129, 4097."""
    for case in GOLD["dlarnv_stream"]["cases"]:
        x, seed = O.dlarnv(case["n"])
        assert [float.hex(float(v)) for v in x[-3:]] == case["last3_hex"], case["n"]
        assert list(seed) == case["seed_after"], case["n"]


def test_dlarnv_float32_stream_fixture(O):
    """_dlarnv_idist_2! with RT = Float32 incl. the rv == 1 reseed branch (src/arpack-blas.jl:520-525): the C++ oracle
    against the independent numpy-float32 restatement committed in tests/golden (seeds crafted to hit the branch at
    draws 5, 128, 129 and 513)."""
    for case in GOLD["dlarnv_stream_float32"]["cases"]:
        x, seed, k = O.dlarnv_f32(case["n"], tuple(case["seed"]))
        assert k == case["reseeds"], case
        assert [float.hex(float(v)) for v in x[:3]] == case["first3_hex"]
        assert [float.hex(float(v)) for v in x[-3:]] == case["last3_hex"]
        assert float.hex(float(sum(float(v) for v in x))) == case["sum_hex"]
        assert list(seed) == case["seed_after"]
        assert x.dtype == np.float32 and np.all(x < 1.0) and np.all(x >= -1.0)
    # away from the reseed branch the Float32 stream is the rounded Float64 stream up to one binary32 ulp
    xf, _, k = O.dlarnv_f32(5000)
    xd, _ = O.dlarnv(5000)
    assert k == 0 and np.max(np.abs(xf.astype(np.float64) - xd)) <= 2.0 ** -23


def test_lcg_table_rows(O):
    """rows 1, 2, 3, 64, 128 of _dlaruv_mm (src/arpack-blas.jl:281-408) given literally."""
    t = O.lcg_table()
    assert tuple(t[0]) == (494, 322, 2508, 2549)
    assert tuple(t[1]) == (2637, 789, 3754, 1145)
    assert tuple(t[2]) == (255, 1440, 1766, 2253)
    assert tuple(t[63]) == (2770, 1238, 1956, 2817)
    assert tuple(t[127]) == (545, 2366, 3801, 1537)


@pytest.mark.parametrize("n", [0, 1, 2, 127, 128, 129, 255, 256, 257, 1000, 4097])
def test_lcg_closed_form(O, n):
    """X_g = seed * a^g mod 2^48 (what test/dlarnv_arpackjll.jl:26-57 checks against LAPACK,
    including the seed left behind)."""
    a = (494 << 36) | (322 << 24) | (2508 << 12) | 2549
    X0 = (1 << 36) | (3 << 24) | (5 << 12) | 7
    x, seed = O.dlarnv(n)
    g = np.arange(1, n + 1)
    exp = np.array([2 * (((X0 * pow(a, int(k), 1 << 48)) % (1 << 48)) / 2.0 ** 48) - 1 for k in g])
    assert np.array_equal(x, exp)
    if n > 0:
        Xn = (X0 * pow(a, n, 1 << 48)) % (1 << 48)
        assert seed == ((Xn >> 36) & 4095, (Xn >> 24) & 4095, (Xn >> 12) & 4095, Xn & 4095)
    else:
        assert seed == (1, 3, 5, 7)


def test_lcg_complex_and_double64(O):
    """complex: element i = (value 2i+1, value 2i+2) (src/arpack-blas.jl:531-536); Double64: lo = 0."""
    xr, _ = O.dlarnv(20)
    xc, seedc = O.dlarnv(10, dtype=O.C64)
    assert np.array_equal(xc.real, xr[0::2]) and np.array_equal(xc.imag, xr[1::2])
    xd, seedd = O.dlarnv(20, dtype=O.DD)
    assert np.array_equal(xd[:, 0], xr) and np.all(xd[:, 1] == 0)
    assert seedc == seedd


# ---------------------------------------------------------------- norm2
def test_norm2(O):
    """test/runtests.jl:164-166 (zeros) and test/dnrm2_openblas.jl (<= 1 ulp of the exact norm)."""
    assert O.norm2(np.zeros(5)) == 0.0
    rng = np.random.default_rng(0)
    import mpmath as mp

    mp.mp.prec = 200
    for n in [1, 2, 7, 8, 9, 100, 1001]:
        x = rng.standard_normal(n)
        exact = mp.sqrt(sum(mp.mpf(float(v)) ** 2 for v in x))
        got = O.norm2(x)
        assert abs(mp.mpf(got) - exact) <= abs(exact) * 2.0 ** -52
    big = np.array([1e200, 1e200, 3e-200])  # x87 range: no overflow
    assert O.norm2(big) == pytest.approx(np.sqrt(2) * 1e200, rel=1e-15)
    z = rng.standard_normal(9) + 1j * rng.standard_normal(9)
    assert O.norm2(z) == pytest.approx(np.linalg.norm(z), rel=1e-15)


# ---------------------------------------------------------------- host pieces
def test_dsconv_dsortr_dsgets(O):
    """test/runtests.jl:113-140, test/dsgets_simple.jl."""
    assert O.dsconv(np.ones(10), np.zeros(10), 1e-4) == 10
    assert O.dsconv(np.ones(10), np.ones(10), 1e-4) == 0
    x1, x2 = O.dsortr("LA", True, [3.0, 1.0, 2.0], [30.0, 10.0, 20.0])
    assert list(x1) == [1.0, 2.0, 3.0] and list(x2) == [10.0, 20.0, 30.0]
    x1, _ = O.dsortr("SA", False, [3.0, 1.0, 2.0], [0.0, 0.0, 0.0])
    assert list(x1) == [3.0, 2.0, 1.0]
    x1, _ = O.dsortr("LM", False, [-3.0, 1.0, 2.0], [0.0, 0.0, 0.0])
    assert list(np.abs(x1)) == [1.0, 2.0, 3.0]
    ritz, bounds, shifts = O.dsgets(1, "LM", 2, 3, [1.0, -5.0, 3.0, 2.0, -4.0], [0.1, 0.5, 0.3, 0.2, 0.4])
    assert sorted(np.abs(ritz[3:])) == [4.0, 5.0]          # wanted in the last kev slots
    assert list(bounds[:3]) == sorted(bounds[:3], reverse=True)  # shifts: largest bounds first
    assert np.array_equal(shifts, ritz[:3])


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11, 13, 16, 17, 99, 127, 128, 129, 500, 513, 1000, 1025])
def test_dstqrb_laplacian(O, n):
    """test/dstqrb.jl:104-112."""
    d, z = O.dstqrb(2 * np.ones(n), -np.ones(max(n - 1, 0)))
    k = np.arange(1, n + 1)
    lams = np.sort(-2 * np.cos(np.pi * k / (n + 1)) + 2.0)
    assert np.max(np.abs(lams - d)) <= 2 * np.sqrt(n) * np.finfo(float).eps
    zexp = np.abs(np.sqrt(2 / (n + 1)) * np.sin(n * k * np.pi / (n + 1)))
    assert np.linalg.norm(np.abs(z) - zexp) <= 2 * n * np.finfo(float).eps


@pytest.mark.parametrize("n", [2, 8, 16, 128, 500, 1024, 3, 9, 127, 129, 1001, 2001])
def test_dstqrb_ones(O, n):
    """test/dstqrb.jl:87-103."""
    d, z = O.dstqrb(np.zeros(n), np.ones(n - 1))
    k = np.arange(1, n + 1)
    if n % 2 == 0:
        lams = np.sort(2 * np.cos(np.pi * k / (n + 1)))
        assert np.max(np.abs(d - lams) / np.spacing(np.abs(lams))) <= 2 * n
    else:
        assert abs(d[n // 2]) <= 10 * np.finfo(float).eps
    zexp = np.abs(np.sqrt(2 / (n + 1)) * np.sin(n * k * np.pi / (n + 1)))
    assert np.linalg.norm(np.abs(z) - zexp) <= 2 * n * np.finfo(float).eps


def test_dstqrb_zeros_and_dsteqr(O):
    """test/dstqrb.jl:80-86 and test/dsteqr-simple.jl (eigenvector matrix is orthogonal, T Z = Z D)."""
    for n in range(1, 8):
        d, z = O.dstqrb(np.zeros(n), np.zeros(max(n - 1, 0)))
        assert np.all(d == 0) and z[-1] == 1.0 and np.all(z[:-1] == 0)
    rng = np.random.default_rng(3)
    n = 12
    dd, ee = rng.standard_normal(n), rng.standard_normal(n - 1)
    ee[5] = 0.0  # split case (test/dsteqr-compare.jl)
    T = np.diag(dd) + np.diag(ee, 1) + np.diag(ee, -1)
    lam, Z = O.dsteqr(dd, ee)
    assert np.allclose(Z.T @ Z, np.eye(n), atol=1e-14)
    assert np.allclose(T @ Z, Z * lam, atol=1e-13)
    assert np.allclose(lam, np.linalg.eigvalsh(T), atol=1e-13)


def test_dseigt(O):
    """test/runtests.jl:141-162: tridiagonal Laplacian through dseigt!, bounds = rnorm |last row|."""
    n = 30
    H = np.zeros((n, 2))
    H[:, 1] = 2.0
    H[1:, 0] = -1.0
    eig, bounds = O.dseigt(0.5, H)
    k = np.arange(1, n + 1)
    assert np.allclose(eig, np.sort(2 - 2 * np.cos(np.pi * k / (n + 1))), atol=1e-14)
    assert np.allclose(np.sort(bounds), np.sort(0.5 * np.abs(np.sqrt(2 / (n + 1)) * np.sin(n * k * np.pi / (n + 1)))), atol=1e-14)


# ---------------------------------------------------------------- dsaitr / dgetv0 invariants
def test_dsaitr_orthogonality(O):
    """test/dsaitr_simple.jl:95-137: V'V = I, ||V'r|| ~ 0 after 3 and then 6 more steps."""
    n = 10
    o = O.Oracle(sp.diags(np.arange(1.0, n + 1)).tocsr(), 9)
    rng = np.random.default_rng(1)
    r0 = rng.standard_normal(n)
    o.resid[:] = r0
    o.rnorm = np.linalg.norm(r0)
    assert o.dsaitr(0, 3) == 0
    V = o.Vt[:3].T
    assert np.allclose(V.T @ V, np.eye(3), atol=1e-14)
    assert np.linalg.norm(V.T @ o.resid) <= 5e-15
    assert o.dsaitr(3, 6) == 0
    V = o.Vt[:9].T
    assert np.allclose(V.T @ V, np.eye(9), atol=1e-13)
    assert np.linalg.norm(V.T @ o.resid) <= 1e-13
    st = o.stats()
    assert st["nrstrt"] == 0 and st["nopx"] == 9
    # Lanczos relation A V = V T + r e_k'
    H = o.H
    T = np.diag(H[:9, 1]) + np.diag(H[1:9, 0], 1) + np.diag(H[1:9, 0], -1)
    A = np.diag(np.arange(1.0, n + 1))
    R = A @ V - V @ T
    assert np.allclose(R[:, :8], 0, atol=1e-12) and np.allclose(R[:, 8], o.resid, atol=1e-12)


def test_dsaitr_identity_restart(O):
    """test/dsaitr_simple.jl:140-178: identity operator -> invariant subspace -> restart, resid ~ 0."""
    n = 10
    o = O.Oracle(sp.identity(n).tocsr(), 3)
    o.resid[:] = np.arange(1.0, n + 1)
    o.rnorm = np.linalg.norm(o.resid)
    o.dsaitr(0, 1)
    assert o.stats()["nrstrt"] == 0
    assert np.linalg.norm(o.resid) <= n * np.finfo(float).eps
    o.resid[:] = np.arange(1.0, n + 1)
    o.rnorm = np.linalg.norm(o.resid)
    o.dsaitr(0, 2)
    assert o.stats()["nrstrt"] > 0
    assert np.linalg.norm(o.resid) <= n * np.finfo(float).eps


def test_dsaitr_zero_initial_resid(O):
    """test/dsaitr_simple.jl:180-212: rnorm = 0 on entry -> dgetv0 restart, still orthonormal."""
    n = 10
    o = O.Oracle(sp.diags(np.arange(1.0, n + 1)).tocsr(), 3)
    o.resid[:] = np.random.default_rng(0).standard_normal(n)
    o.rnorm = 0.0
    assert o.dsaitr(0, 3) == 0
    V = o.Vt[:3].T
    st = o.stats()
    assert st["nrstrt"] > 0 and st["nopx"] > 0
    assert np.allclose(V.T @ V, np.eye(3), atol=1e-14)
    assert np.linalg.norm(V.T @ o.resid) <= 4 * np.finfo(float).eps * max(1.0, np.linalg.norm(o.resid))


def test_dgetv0_restart_failure(O):
    """test/dgetv0_simple.jl:149-196: V = I spans everything -> ierr = -1, resid = 0; then a case
    that works (orthogonalise against the constant vector)."""
    n = 5
    o = O.Oracle(sp.diags(np.arange(1.0, n + 1)).tocsr(), 6)
    o.Vt[:5] = np.eye(n)
    o.resid[:] = 2 * np.random.default_rng(0).random(n) - 1
    ierr = o.dgetv0(itry=1, initv=True, j=6)
    assert ierr == -1 and np.all(o.resid == 0) and o.rnorm == 0.0
    o = O.Oracle(sp.diags(np.arange(1.0, n + 1)).tocsr(), 3)
    o.Vt[0] = 1 / np.sqrt(n)
    o.resid[:] = 2 * np.random.default_rng(1).random(n) - 1
    ierr = o.dgetv0(itry=1, initv=True, j=2)
    assert ierr == 0 and abs(np.sum(o.resid)) <= n * np.finfo(float).eps and o.rnorm > 0


def test_dgetv0_initv_semantics(O):
    """test/dgetv0_simple.jl:40-69 (bmat = :I, initv = true: resid and the seed are left alone, rnorm = norm2(resid))
    and :91-147 (bmat = :G, initv = true: resid <- OP resid0, rnorm = sqrt|resid' B resid|, one OP*x, one B*x)."""
    n = 10
    M = sp.diags(np.arange(1.0, n + 1)).tocsr()
    r0 = 2 * np.random.default_rng(0).random(n) - 1
    o = O.Oracle(M, 2)
    o.resid[:] = r0
    assert o.dgetv0(itry=1, initv=True, j=1) == 0
    assert np.array_equal(o.resid, r0) and tuple(o.iseed) == (1, 3, 5, 7)
    assert o.rnorm == pytest.approx(np.linalg.norm(r0), rel=1e-15)
    B = sp.diags(np.arange(1, n + 1) / 10.0).tocsr()
    o = O.Oracle((B @ M).tocsr(), 2, B=B)                  # OP = inv(B) (B M) = M
    o.resid[:] = r0
    assert o.dgetv0(itry=1, initv=True, j=1) == 0
    assert np.allclose(o.resid, M @ r0, rtol=1e-14, atol=0) and tuple(o.iseed) == (1, 3, 5, 7)
    assert o.rnorm == pytest.approx(np.sqrt(abs((M @ r0) @ (B @ (M @ r0)))), rel=1e-14)
    st = o.stats()
    assert st["nopx"] == 1 and st["nbx"] == 1


def test_dgetv0_generalized_restart_failure(O):
    """test/dgetv0_simple.jl:198-247: OP resid0 lies in the span of the (B-orthonormal) basis, so all six
    orthogonalisation passes collapse: ierr = -1, one OP*x, 7 B*x (the initial one + 6).  The reference uses a singular
    B = diag(1,1,0,...); the same situation with a positive definite B: A = diag(a1, a2, 0, ...) so that
    OP resid0 = inv(B) A resid0 has two non-zero entries, and V[:, i] = e_i / sqrt(B_ii)."""
    n, j = 8, 3
    b = np.arange(1, n + 1) / 4.0
    a = np.zeros(n); a[:2] = [3.0, -2.0]
    o = O.Oracle(sp.diags(a).tocsr(), j + 1, B=sp.diags(b).tocsr())
    for i in range(j):
        o.Vt[i] = 0.0
        o.Vt[i, i] = 1 / np.sqrt(b[i])
    o.resid[:] = 2 * np.random.default_rng(0).random(n) - 1
    assert o.dgetv0(itry=1, initv=True, j=j) == -1
    assert np.all(o.resid == 0) and o.rnorm == 0.0
    st = o.stats()
    assert st["nopx"] == 1 and st["nbx"] == 7


# ---------------------------------------------------------------- eigenvalue KATs
def _top(w, k, which="LM"):
    if which == "LM":
        return np.sort(w[np.argsort(-np.abs(w))[:k]])
    return np.sort(np.sort(w)[-k:])


def test_kat_15x15(O):
    """test/interface_simple.jl:7-11 (default ncv) and test/dsaupd_arpackjll.jl:232-241 (ncv=6, nev=3:
    'needs early shift and deflation')."""
    A = kat15()
    w = np.linalg.eigvalsh(A.toarray())
    for ncv in (15, 6):
        r = O.Oracle(A, ncv).symeigs(3)
        assert r["info"] == 0 and r["nconv"] == 3
        assert np.allclose(r["values"], _top(w, 3), rtol=1e-12)
        Z = r["vectors"]
        assert np.allclose(A @ Z, Z * r["values"], atol=1e-11)
        assert np.allclose(Z.T @ Z, np.eye(3), atol=1e-12)
    r = O.Oracle(A, 15).symeigs(3, which="LA", ritzvec=False)
    assert np.allclose(r["values"], _top(w, 3, "LA"), rtol=1e-12)


@pytest.mark.parametrize("n,nev,ncv", [(10, 3, 6), (30, 4, 10), (1000, 6, 12)])
def test_kat_diagonal(O, n, nev, ncv):
    """test/dsaupd_simple.jl:61-64, test/dsaupd_idonow.jl:3-15, test/dsaupd_arpackjll.jl:134-149."""
    A = sp.diags(np.arange(1.0, n + 1)).tocsr()
    r = O.Oracle(A, ncv).symeigs(nev, v0=np.ones(n) / np.sqrt(n))
    assert r["info"] == 0
    assert np.allclose(r["values"], np.arange(n - nev + 1, n + 1), rtol=1e-12)


def test_kat_issue5(O):
    """test/interface_simple.jl:15-24: rank-deficient 10 x 10, nev = 5, ncv = 8."""
    A = sp.lil_matrix((10, 10))
    for i in (0, 1, 4, 5):
        A[i, i] = 1.0
    A = A.tocsr()
    r = O.Oracle(A, 8).symeigs(5)
    Z, lam = r["vectors"], r["values"]
    assert np.allclose(A @ Z, Z * lam, atol=1e-12)
    assert np.allclose(Z.T @ Z, np.eye(5), atol=1e-12)
    assert np.sum(lam) == pytest.approx(4.0, rel=1e-12)


def test_kat_complex(O):
    """test/dsaupd_complex.jl:78-80 / docs example: Hermitian tridiagonal, Ritz values vs dense eigvals."""
    n = 20
    H = (sp.diags([1j * np.ones(n - 1), 2 * np.ones(n) + 0j, -1j * np.ones(n - 1)], [-1, 0, 1]) * (n + 1)).tocsr()
    w = np.linalg.eigvalsh(H.toarray())
    r = O.Oracle(H, 12).symeigs(4)
    assert r["info"] == 0
    assert np.allclose(r["values"], _top(w, 4), rtol=1e-12)
    Z = r["vectors"]
    assert np.allclose(Z.conj().T @ Z, np.eye(4), atol=1e-12)


def test_kat_double64(O):
    """test/interface_high.jl:60-66: tridiagonal Laplacian (dsdrv3's A) in Double64 to ~1e-28."""
    import mpmath as mp

    mp.mp.prec = 250
    n = 100
    A = (sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * (n + 1)).tocsr()
    r = O.Oracle(A, 20, dtype=O.DD).symeigs(4)
    assert r["info"] == 0
    exact = [(n + 1) * (2 - 2 * mp.cos(k * mp.pi / (n + 1))) for k in range(n - 3, n + 1)]
    got = [mp.mpf(float(v[0])) + mp.mpf(float(v[1])) for v in r["values"]]
    assert max(abs(g - e) / e for g, e in zip(got, exact)) < mp.mpf(10) ** -28


def test_kat_float32_vectors(O, syn):
    """Float32 vectors with a Float64 factorisation, symeigs(Float32, Float64, op, nev) (src/interface.jl:26-34):
    the reference's own Float32 bars -- `vals ≈ Float32.(...)` on the 15 x 15 KAT with which = :LA
    (test/interface_simple.jl:36-49) and the docstring example symeigs(Float32, ArpackSimpleOp(Diagonal(1.0:n)), 5;
    which=:SA) (src/interface.jl:105-106) -- plus the start vector (binary32 dlarnv), a lock step against the Float64
    instantiation and the analytic 2-D Laplacian."""
    A = kat15()
    w = np.linalg.eigvalsh(A.toarray())
    r = O.Oracle(A, 15, dtype=O.F32).symeigs(3, which="LA", ritzvec=False)
    assert r["info"] == 0 and np.allclose(r["values"], np.sort(w)[-3:], rtol=1e-5)
    r = O.Oracle(A, 15, dtype=O.F32).symeigs(3, tol=2.0 ** -24)
    assert r["vectors"].dtype == np.float32
    Z = r["vectors"].astype(np.float64)
    assert np.allclose(r["values"], _top(w, 3), rtol=1e-5)
    assert np.max(np.abs(A @ Z - Z * r["values"])) <= 2e-5 and np.max(np.abs(Z.T @ Z - np.eye(3))) <= 1e-5
    D = sp.diags(np.arange(1.0, 101.0)).tocsr()
    r = O.Oracle(D, 20, dtype=O.F32).symeigs(5, which="SA", tol=2.0 ** -24)
    assert r["info"] == 0 and np.allclose(r["values"], np.arange(1.0, 6.0), rtol=1e-5)
    m = 60
    L = syn.to_scipy(syn.laplacian2d(m))
    r = O.Oracle(L, 40, dtype=O.F32).symeigs(10, tol=2.0 ** -24)
    exact = syn.laplacian2d_eigs(m, 10)
    assert r["info"] == 0 and np.max(np.abs(r["values"] - exact) / exact) <= 1e-6
    o32 = O.Oracle(L, 12, dtype=O.F32)
    assert o32.dgetv0(1, False, 1) == 0
    ref, seed, _ = O.dlarnv_f32(m * m)
    assert np.asarray(o32.resid).tobytes() == ref.tobytes() and tuple(o32.iseed) == seed
    v0 = np.asarray(o32.resid).astype(np.float64).copy()
    assert o32.dsaitr(0, 12) == 0
    o64 = O.Oracle(L, 12)
    o64.resid[:] = v0
    o64.dgetv0(1, True, 1)
    o64.dsaitr(0, 12)
    assert np.max(np.abs(np.asarray(o32.H) - np.asarray(o64.H))) <= 1e-5
    V = np.asarray(o32.Vt, dtype=np.float64)
    assert np.max(np.abs(V @ V.T - np.eye(12))) <= 1e-5
    assert o32.stats()["nrorth"] == o64.stats()["nrorth"]


def test_svds_normal_op(O):
    """test/interface_simple.jl:100-105 style: ArpackNormalOp eigenvalues = squared singular values."""
    rng = np.random.default_rng(5)
    A = sp.random(60, 25, density=0.2, random_state=np.random.RandomState(5), format="csr")
    o = O.Oracle(A, 12, normal=True)
    assert o.n == 25
    r = o.symeigs(3)
    s = np.linalg.svd(A.toarray(), compute_uv=False)
    assert np.allclose(np.sqrt(r["values"]), np.sort(s[:3]), rtol=1e-10)
    o2 = O.Oracle(A.T.tocsr(), 12, normal=True)
    assert o2.n == 25
    r2 = o2.symeigs(3)
    assert np.allclose(r2["values"], r["values"], rtol=1e-10)


def test_vs_scipy_arpack(O, syn):
    """Independent implementation of the same algorithm (SciPy's ARPACK): same eigenvalues and a
    comparable amount of work on the 2-D Laplacian."""
    import scipy.sparse.linalg as sla

    m = 60
    A = syn.to_scipy(syn.laplacian2d(m))
    cnt = [0]

    def mv(x):
        cnt[0] += 1
        return A @ x

    v0, _ = O.dlarnv(m * m)
    w = sla.eigsh(sla.LinearOperator(A.shape, matvec=mv, dtype=float), k=10, ncv=40, which="LM", v0=v0, tol=0,
                  return_eigenvectors=False)
    r = O.Oracle(A, 40).symeigs(10)
    assert np.allclose(np.sort(w), r["values"], rtol=1e-12)
    assert np.allclose(r["values"], syn.laplacian2d_eigs(m, 10), rtol=1e-12)
    assert abs(r["stats"]["nopx"] - cnt[0]) <= 0.15 * cnt[0]


# ---------------------------------------------------------------- generalised problem: mode 2, bmat = :G
def test_generalized_dsdrv3_golden(O):
    """test/interface_high.jl:45,72,102: eigs(Symmetric(A), Symmetric(B), 4) on dsdrv3, n = 100, gives
    [121003.49732902342, 121616.60247324049, 122057.49457079473, 122323.22366457577]."""
    A, B = _dsdrv3(100)
    o = O.Oracle(A, 20, B=B)
    r = o.symeigs(4)
    assert r["info"] == 0 and r["ierr"] == 0 and r["nconv"] == 4
    assert np.allclose(r["values"], DSDRV3_GOLDEN, rtol=1e-13, atol=0)
    assert int(o.iparam[6]) == 2                                     # mode
    st = r["stats"]
    assert st["nbx"] > st["nopx"] > 0                                # every B-norm costs one B*x
    Z = r["vectors"]
    assert np.allclose(Z.T @ (B @ Z), np.eye(4), atol=1e-12)         # B-orthonormal Ritz vectors
    for i in range(4):
        assert np.linalg.norm(A @ Z[:, i] - r["values"][i] * (B @ Z[:, i])) <= 1e-9 * np.linalg.norm(A @ Z[:, i])


@pytest.mark.parametrize("which,maxiter", [("LM", 300), ("SA", 400)])
def test_generalized_vs_dense(O, which, maxiter):
    """test/dsaupd_idonow.jl:17-110: n = 100, ncv = 10, nev = 4, mode 2 / bmat :G against eigvals(B \\ A)."""
    import scipy.linalg as sl

    A, B = _dsdrv3(100)
    w = sl.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    r = O.Oracle(A, 10, B=B).symeigs(4, which=which, maxiter=maxiter)
    ref = np.sort(w)[-4:] if which == "LM" else np.sort(w)[:4]
    assert r["info"] == 0
    assert np.allclose(np.sort(r["values"]), ref, rtol=1e-10, atol=0)


def test_generalized_double64(O):
    """dsdrv3 in Double64 (test/interface_high.jl:69-73) against the analytic pencil eigenvalues
    6 (n+1)^2 (2 - 2 cos t) / (4 + 2 cos t), t = k pi / (n+1), to 1e-26 relative."""
    import mpmath as mp

    mp.mp.prec = 250
    n = 100
    A, B = _dsdrv3(n)
    # B's entries 1/(6(n+1)) are not doubles: build the Double64 matrix from exact (hi, lo) pairs
    def dd(x):
        hi = float(x)
        return hi, float(x - mp.mpf(hi))
    h = mp.mpf(1) / (6 * (n + 1))
    Bd = B.copy().astype(np.float64)
    vals = np.zeros((B.nnz, 2))
    for k in range(B.nnz):
        vals[k] = dd(4 * h) if abs(B.data[k]) > 2.5 * float(h) else dd(h)
    o = O.Oracle(A, 20, dtype=O.DD, B=Bd, b_val=vals)
    r = o.symeigs(4)
    assert r["info"] == 0 and r["ierr"] == 0
    d = r["values"]
    exact = [6 * (n + 1) ** 2 * (2 - 2 * mp.cos(k * mp.pi / (n + 1))) / (4 + 2 * mp.cos(k * mp.pi / (n + 1)))
             for k in range(n - 3, n + 1)]
    got = [mp.mpf(float(v[0])) + mp.mpf(float(v[1])) for v in d]
    assert max(abs(g - e) / e for g, e in zip(got, exact)) < mp.mpf(10) ** -26


def test_generalized_complex_hermitian(O):
    """Hermitian A, Hermitian positive definite B (hermeigs(A, B, k), src/interface.jl:158-163) against
    scipy.linalg.eigh(A, B)."""
    import scipy.linalg as sl

    rng = np.random.default_rng(3)
    n = 60
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, n - 1))
    A = sp.diags([np.conj(ph), rng.uniform(1, 3, n), ph], [-1, 0, 1]).tocsr()
    pb = 0.3 * np.exp(1j * rng.uniform(0, 2 * np.pi, n - 1))
    B = sp.diags([np.conj(pb), rng.uniform(1.5, 2.5, n), pb], [-1, 0, 1]).tocsr()
    w = sl.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    r = O.Oracle(A, 16, B=B).symeigs(4)
    assert r["info"] == 0
    ref = w[np.argsort(-np.abs(w))[:4]]
    assert np.allclose(np.sort(r["values"]), np.sort(ref), rtol=1e-10, atol=0)
    Z = r["vectors"]
    assert np.allclose(Z.conj().T @ (B @ Z), np.eye(4), atol=1e-11)


def test_generalized_rejects_indefinite_b(O):
    A, B = _dsdrv3(20)
    Bbad = B.tolil(); Bbad[3, 3] = -1.0
    with pytest.raises(RuntimeError):
        O.Oracle(A, 8, B=Bbad.tocsr())


def test_generalized_vs_scipy_arpack(O):
    """SciPy's ARPACK in its own mode 2 (eigsh(A, M=B) with the exact solve for B) on dsdrv3 and on a random SPD pencil:
    an independent implementation of the generalised Lanczos iteration gives the same eigenvalues, and a comparable
    number of OP applications from the same start vector."""
    import scipy.sparse.linalg as sla

    A, B = _dsdrv3(100)
    lu = sla.splu(B.tocsc())
    cnt = [0]

    def minv(x):
        cnt[0] += 1
        return lu.solve(x)

    v0, _ = O.dlarnv(100)
    w = sla.eigsh(A, k=4, M=B, Minv=sla.LinearOperator(A.shape, matvec=minv, dtype=float), ncv=20, which="LM", v0=v0, tol=0,
                  return_eigenvectors=False)
    r = O.Oracle(A, 20, B=B).symeigs(4)
    assert np.allclose(np.sort(w), r["values"], rtol=1e-12)
    assert abs(r["stats"]["nopx"] - cnt[0]) <= 0.25 * cnt[0] + 20
    rng = np.random.default_rng(11)
    n = 60
    M1 = sp.random(n, n, density=0.08, random_state=np.random.RandomState(1), format="csr")
    A2 = (M1 + M1.T + sp.diags(rng.standard_normal(n))).tocsr()
    M2 = sp.random(n, n, density=0.05, random_state=np.random.RandomState(2), format="csr")
    B2 = (M2 @ M2.T + 2.0 * sp.identity(n)).tocsr()
    w2 = sla.eigsh(A2, k=5, M=B2, ncv=24, which="LA", tol=0, return_eigenvectors=False)
    r2 = O.Oracle(A2, 24, B=B2).symeigs(5, which="LA")
    assert r2["info"] == 0 and np.allclose(np.sort(w2), r2["values"], rtol=1e-10)

