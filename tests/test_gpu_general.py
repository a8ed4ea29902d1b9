"""GPU parity tests for the generalised problem A x = lambda B x (mode 2, bmat = :G;
ArpackSymmetricGeneralizedOp, src/idonow_ops.jl:462-492) through the C ABI: the device operator
(OP = inv(B) A by conjugate gradients, B x), the B-norm branches of dgetv0!/dsaitr!
(src/dsaitr.jl:1195-1204,1298-1300,1405-1407) in lock step with the oracle, whole solves against the
reference's dsdrv3 golden eigenvalues (test/interface_high.jl:45,72,102), eigvals(B \\ A)
(test/dsaupd_idonow.jl:17-110), Double64 (1e-26) and the Hermitian pencil."""
import numpy as np
import pytest
import scipy.linalg as sl
import scipy.sparse as sp

from conftest import DSDRV3_GOLDEN, dsdrv3

pytestmark = pytest.mark.gpu


def test_generalized_operator(pkg):
    """opx = inv(B) (A x) to CG tolerance, bx = B x exactly the CSR product."""
    A, B = dsdrv3(100)
    ctx = pkg.Context(pkg.ArpackSymmetricGeneralizedOp(A, B), 20)
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(100)
        y = np.asarray(ctx.opx(x))
        ref = sl.solve(B.toarray(), A @ x)
        assert np.linalg.norm(y - ref) <= 1e-13 * np.linalg.norm(ref)
        bx = np.asarray(ctx.bx(x))
        assert np.linalg.norm(bx - B @ x) <= 1e-15 * np.linalg.norm(B @ x)
    nsolve, niter, relres = ctx.solve_stats()
    assert nsolve == 3 and niter > 0 and relres <= 1e-14
    ctx.close()


def test_generalized_lockstep_with_oracle(pkg, O):
    """dgetv0! (bmat = :G: OP applied to the start vector, B-norm) then 12 dsaitr! steps: H, V, rnorm
    and every counter (nopx, nbx, nrorth, nitref) against the oracle; V is B-orthonormal."""
    A, B = dsdrv3(100)
    ncv = 12
    ctx = pkg.Context(pkg.ArpackSymmetricGeneralizedOp(A, B), ncv)
    o = O.Oracle(A, ncv, B=B)
    rc, rn = ctx.getv0()
    o.dgetv0(1, False, 1)
    assert rc == 0
    assert abs(rn - o.rnorm) <= 1e-12 * o.rnorm
    assert abs(ctx.bnorm_resid() - rn) <= 1e-14 * rn          # sqrt|r' B r| recomputed (src/dsaup2.jl:1378-1411)
    r_gpu = np.asarray(ctx.download_resid())
    assert np.max(np.abs(r_gpu - o.resid)) <= 1e-12 * np.max(np.abs(o.resid))
    H = np.zeros((ncv, 2))
    info, rn = ctx.lanczos(0, ncv, H, rn)
    o.dsaitr(0, ncv)
    assert info == 0
    Ho = np.asarray(o.H)
    assert np.max(np.abs(H - Ho)) <= 1e-12 * np.max(np.abs(Ho))
    V = np.asarray(ctx.download_v(0, ncv))
    assert np.max(np.abs(V - np.swapaxes(o.Vt, 0, 1))) <= 1e-10
    assert abs(rn - o.rnorm) <= 1e-10 * max(1.0, np.max(np.abs(Ho)))
    assert np.max(np.abs(V.T @ (B @ V) - np.eye(ncv))) <= 1e-13
    st, so = ctx.stats(), o.stats()
    for k in ("nopx", "nbx", "nrorth", "nitref", "nrstrt"):
        assert st[k] == so[k] + (1 if k == "nbx" else 0), k     # + the B*x of the bnorm_resid() call above
    ctx.close()


@pytest.mark.parametrize("which,ncv,maxiter", [("LM", 20, 300), ("LM", 10, 300), ("SA", 10, 400)])
def test_generalized_eigs_dsdrv3(pkg, O, which, ncv, maxiter):
    A, B = dsdrv3(100)
    st = {}
    res = pkg.eigs(A, B, 4, ncv=ncv, which=which, maxiter=maxiter, stats=st)
    w = sl.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    ref = np.sort(w)[-4:] if which == "LM" else np.sort(w)[:4]
    assert res.info == 0 and int(res.iparam[6]) == 2
    assert np.allclose(np.sort(res.values), ref, rtol=1e-10, atol=0)
    if which == "LM":
        assert np.allclose(res.values, DSDRV3_GOLDEN, rtol=1e-12, atol=0)   # test/interface_high.jl:45
    ora = O.Oracle(A, ncv, B=B).symeigs(4, which=which, maxiter=maxiter)
    assert np.allclose(res.values, ora["values"], rtol=1e-10, atol=0)
    assert abs(int(res.iparam[2]) - ora["niter"]) <= 1
    assert abs(st["nopx"] - ora["stats"]["nopx"]) <= ncv and abs(st["nbx"] - ora["stats"]["nbx"]) <= 3 * ncv
    Z = res.vectors
    assert np.max(np.abs(Z.T @ (B @ Z) - np.eye(4))) <= 1e-12
    for i in range(4):
        assert np.linalg.norm(A @ Z[:, i] - res.values[i] * (B @ Z[:, i])) <= 1e-9 * np.linalg.norm(A @ Z[:, i])


def test_generalized_double64(pkg, O):
    """dsdrv3 in Double64 (test/interface_high.jl:69-73): golden values in the high words, 1e-26 vs the oracle."""
    A, B = dsdrv3(100)
    res = pkg.eigs(A, B, 4, ncv=20, dtype="dd")
    ora = O.Oracle(A, 20, dtype=O.DD, B=B).symeigs(4)
    v, ov = np.asarray(res.values), np.asarray(ora["values"])
    d = (v[:, 0] - ov[:, 0]) + (v[:, 1] - ov[:, 1])
    assert np.max(np.abs(d) / np.abs(ov[:, 0])) <= 1e-26
    assert np.allclose(v[:, 0], DSDRV3_GOLDEN, rtol=1e-13, atol=0)
    assert abs(int(res.iparam[2]) - ora["niter"]) <= 1


def test_generalized_complex_hermitian(pkg):
    """hermeigs(A, B, k) (src/interface.jl:158-163) against scipy.linalg.eigh(A, B)."""
    rng = np.random.default_rng(3)
    n = 60
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, n - 1))
    A = sp.diags([np.conj(ph), rng.uniform(1, 3, n), ph], [-1, 0, 1]).tocsr()
    pb = 0.3 * np.exp(1j * rng.uniform(0, 2 * np.pi, n - 1))
    B = sp.diags([np.conj(pb), rng.uniform(1.5, 2.5, n), pb], [-1, 0, 1]).tocsr()
    w = sl.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    res = pkg.hermeigs(A, B, 4, ncv=16)
    ref = np.sort(w[np.argsort(-np.abs(w))[:4]])
    assert np.allclose(np.sort(res.values), ref, rtol=1e-10, atol=0)
    Z = res.vectors
    assert np.max(np.abs(Z.conj().T @ (B @ Z) - np.eye(4))) <= 1e-11


def test_generalized_large_property(pkg, syn):
    """Size-independent properties at n = 10 000: A = 2-D Laplacian, B = 2 I + A / 4 share eigenvectors, so
    lambda = mu / (2 + mu / 4) is analytic; true residuals and B-orthonormality of the Ritz vectors."""
    m = 100
    A = syn.to_scipy(syn.laplacian2d(m))
    n = m * m
    B = (sp.identity(n) * 2.0 + 0.25 * A).tocsr()
    res = pkg.eigs(A, B, 6, ncv=24, maxiter=1000)
    k = np.arange(1, m + 1)
    mu1 = 2 - 2 * np.cos(k * np.pi / (m + 1))
    mu = np.sort((mu1[:, None] + mu1[None, :]).ravel())[-6:]
    exact = mu / (2 + mu / 4)
    assert np.allclose(np.sort(res.values), exact, rtol=1e-10, atol=0)
    Z = res.vectors
    assert np.max(np.abs(Z.T @ (B @ Z) - np.eye(6))) <= 1e-10
    for i in range(6):
        assert np.linalg.norm(A @ Z[:, i] - res.values[i] * (B @ Z[:, i])) <= 1e-9 * np.linalg.norm(A @ Z[:, i])


def test_getv0_initv_semantics(pkg):
    """test/dgetv0_simple.jl:40-69 (bmat = :I, initv = true: resid and the seed are left alone, rnorm = norm2(resid))
    and :91-147 (bmat = :G, initv = true: resid <- OP resid0, rnorm = sqrt|resid' B resid|, one OP*x, one B*x)."""
    n = 10
    M = sp.diags(np.arange(1.0, n + 1)).tocsr()
    r0 = 2 * np.random.default_rng(0).random(n) - 1
    ctx = pkg.Context(pkg.ArpackSimpleOp(M), 2)
    ctx.upload_resid(r0)
    ierr, rn = ctx.getv0(initv=True, j=1)
    assert ierr == 0 and np.array_equal(np.asarray(ctx.download_resid()), r0) and tuple(ctx.iseed) == (1, 3, 5, 7)
    assert abs(rn - np.linalg.norm(r0)) <= 2 * np.spacing(np.linalg.norm(r0))
    ctx.close()
    B = sp.diags(np.arange(1, n + 1) / 10.0).tocsr()
    ctx = pkg.Context(pkg.ArpackSymmetricGeneralizedOp((B @ M).tocsr(), B), 2)      # OP = inv(B) (B M) = M
    ctx.upload_resid(r0)
    ierr, rn = ctx.getv0(initv=True, j=1)
    r = np.asarray(ctx.download_resid())
    assert ierr == 0 and tuple(ctx.iseed) == (1, 3, 5, 7)
    assert np.allclose(r, M @ r0, rtol=1e-13, atol=0)
    assert abs(rn - np.sqrt(abs((M @ r0) @ (B @ (M @ r0))))) <= 1e-13 * rn
    st = ctx.stats()
    assert st["nopx"] == 1 and st["nbx"] == 1
    ctx.close()


def test_getv0_generalized_restart_failure(pkg):
    """test/dgetv0_simple.jl:198-247: OP resid0 lies in the span of the B-orthonormal basis, all six orthogonalisation
    passes collapse: ierr = -1, resid = 0, one OP*x and 7 B*x (the reference's counts).  Positive definite variant of the
    reference's singular B, see tests/test_oracle.py::test_dgetv0_generalized_restart_failure."""
    n, j = 8, 3
    b = np.arange(1, n + 1) / 4.0
    a = np.zeros(n); a[:2] = [3.0, -2.0]
    A = sp.csr_matrix((a, (np.arange(n), np.arange(n))), shape=(n, n))            # explicit zeros kept on the diagonal
    ctx = pkg.Context(pkg.ArpackSymmetricGeneralizedOp(A, sp.diags(b).tocsr()), j + 1)
    V = np.zeros((n, j))
    for i in range(j):
        V[i, i] = 1 / np.sqrt(b[i])
    ctx.upload_v(V, 0)
    ctx.upload_resid(2 * np.random.default_rng(0).random(n) - 1)
    ierr, rn = ctx.getv0(initv=True, j=j)
    assert ierr == -1 and rn == 0.0 and np.all(np.asarray(ctx.download_resid()) == 0)
    st = ctx.stats()
    assert st["nopx"] == 1 and st["nbx"] == 7
    ctx.close()
