"""CPU: pins the svds post-processing oracle (oracle/oracle_svd.py) to the reference's golden
singular values and check_svd properties (test/interface_simple.jl:56-105, 160-222; test/utility.jl:244-265)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD
from oracle import oracle_svd as OS

EPS = np.finfo(np.float64).eps


def mytestmat(m, n, minval=None, maxdiagval=2.0):
    """test/utility.jl:259-265: [v, B], v = range(1, minval, m), B = diag(range(1, maxdiagval, n-1)) padded to m rows."""
    minval = 1.0 / m if minval is None else minval
    v = np.linspace(1.0, minval, m)
    B = sp.coo_matrix((np.linspace(1.0, maxdiagval, n - 1), (np.arange(n - 1), np.arange(n - 1))), shape=(m, n - 1))
    return sp.hstack([sp.csr_matrix(v.reshape(-1, 1)), B]).tocsr()


def arpack_svd_example(m=500, n=100):
    """the dense operator of ARPACK's dsvd example (test/interface_simple.jl:160-215) as a matrix"""
    h, k = 1.0 / (m + 1), 1.0 / (n + 1)
    A = np.zeros((m, n))
    for j in range(n):
        t = (j + 1) * k
        s = (np.arange(m) + 1) * h
        A[: j + 1, j] = k * s[: j + 1] * (t - 1)
        A[j + 1:, j] = k * t * (s[j + 1:] - 1)
    return A


def check_svd(A, U, s, V, tol):
    k = len(s)
    assert np.allclose(U.T @ U, np.eye(k), atol=1e-12)
    assert np.allclose(V.T @ V, np.eye(k), atol=1e-12)
    assert np.max(OS.svd_residuals(A, U, s, V)) <= EPS * tol


def test_orth_is_positive_diagonal_qr():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((200, 7))
    Q = OS.orth(X)
    Qn, Rn = np.linalg.qr(X)
    Qn = Qn * np.sign(np.diag(Rn))[None, :]
    assert np.allclose(Q, Qn, atol=1e-13)
    assert np.all(np.diag(Q.T @ X) > 0)


def test_golden_mytestmat_be():
    A = mytestmat(10, 8)
    U, s, V, _ = OS.svds(A, 2, which="BE")
    assert np.allclose(s, GOLD["svds_mytestmat_10_8_BE"]["values"], rtol=1e-12)     # test/interface_simple.jl:104
    check_svd(A, U, s, V, tol=75)


def test_golden_all_ones():
    A = sp.csr_matrix(np.ones((5, 5)))
    U, s, V, _ = OS.svds(A, 1)
    assert np.allclose(s, [5.0])                                                      # :59-62
    check_svd(A, U, s, V, tol=10)


def test_golden_arpack_svd_example():
    A = arpack_svd_example()
    U, s, V, _ = OS.svds(sp.csr_matrix(A), 4, ncv=10)
    gold = sorted(GOLD["svds_arpack_example_500x100"]["values"], reverse=True)
    assert np.allclose(s, gold, rtol=1e-10)                                           # :218-220
    check_svd(A, U, s, V, tol=3)


def test_golden_arpack_svd_example_float32():
    """svds(Float32, op, 4; ncv=10) ≈ [4.10123E-02, 6.04880E-02, 1.17844E-01, 5.57234E-01] (test/interface_simple.jl:221-222)"""
    from oracle import oracle as O

    A = arpack_svd_example()
    U, s, V, res = OS.svds(sp.csr_matrix(A), 4, ncv=10, dtype=O.F32)
    gold = sorted(GOLD["svds_arpack_example_500x100_float32"]["values"], reverse=True)
    assert res["vectors"].dtype == np.float32
    assert np.allclose(s, gold, rtol=2e-4)            # Julia's ≈ on Float32: rtol = sqrt(eps(Float32)) = 3.5e-4
    assert np.max(np.abs(A @ V - U * s)) <= 1e-4


def test_wide_matrix_right_side():
    A = mytestmat(10, 8).T.tocsr()          # m <= n: OP = A A', eigenvectors are U
    U, s, V, _ = OS.svds(A, 2, which="BE")
    assert np.allclose(s, GOLD["svds_mytestmat_10_8_BE"]["values"], rtol=1e-12)
    check_svd(A, U, s, V, tol=75)


def test_singular_case():
    n = 80
    A = sp.vstack([sp.diags(np.arange(1.0, n + 1)), sp.csr_matrix((20, n))]).tolil()
    A[n - 5:n, n - 5:n] = 0
    A = A.tocsr()
    try:                                                                              # :89-91
        U, s, V, _ = OS.svds(A, 1, which="SA", ncv=20)
    except ArithmeticError:
        return  # sigma ~ 0 met when forming U: the reference raises ArpackException in that case (:300-303)
    assert np.max(OS.svd_residuals(A, U, s, V)) <= n * np.sqrt(EPS)
