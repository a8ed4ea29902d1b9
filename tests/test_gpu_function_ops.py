"""GPU tests for the function operators of the reference's ArpackOp API: ArpackSimpleFunctionOp(F, n)
(src/idonow_ops.jl:360-368) and ArpackNormalFunctionOp(av, atv, m, n) (:406-431) with the caller's OP*x running on
DEVICE vectors inside the enqueued Lanczos steps (C ABI: ga_set_op_callback / ga_set_normal_op_callbacks).  The
callbacks here are a few torch ops on tensors that alias the library's vectors; torch is only the caller's tool."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD

pytestmark = pytest.mark.gpu


def _lap1d(y, x):
    """matrix-free 1-D Laplacian tridiag(-1, 2, -1), in place on device tensors"""
    y.copy_(x).mul_(2.0)
    y[1:] -= x[:-1]
    y[:-1] -= x[1:]


def test_simple_function_op_matches_matrix_op(pkg, O):
    n = 300
    A = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsr()
    calls = [0]

    def F(y, x):
        calls[0] += 1
        _lap1d(y, x)

    op = pkg.ArpackSimpleFunctionOp(F, n)
    ctx = pkg.Context(op, 8)
    x = np.random.default_rng(0).standard_normal(n)
    assert np.allclose(np.asarray(ctx.opx(x)), A @ x, rtol=0, atol=1e-14)
    ctx.close()
    st = {}
    res = pkg.symeigs(op, 4, ncv=24, stats=st)
    ref = pkg.eigs(A, 4, ncv=24)
    exact = 2 - 2 * np.cos(np.arange(n - 3, n + 1) * np.pi / (n + 1))
    assert np.allclose(res.values, exact, rtol=1e-12, atol=0)
    assert np.allclose(res.values, ref.values, rtol=1e-13, atol=0)
    assert abs(int(res.iparam[2]) - int(ref.iparam[2])) <= 1          # same restarts as the CSR operator
    assert calls[0] == st["nopx"] + 1                                   # one callback per OP*x (+ the ctx.opx above)
    Z = res.vectors
    assert np.max(np.abs(A @ Z - Z * res.values)) <= 1e-11
    assert np.max(np.abs(Z.T @ Z - np.eye(4))) <= 1e-12


def test_simple_function_op_float32_and_errors(pkg):
    n = 200
    res = pkg.symeigs(pkg.ArpackSimpleFunctionOp(_lap1d, n, dtype="f32"), 3, ncv=20)
    exact = 2 - 2 * np.cos(np.arange(n - 2, n + 1) * np.pi / (n + 1))
    assert res.vectors.dtype == np.float32 and np.allclose(res.values, exact, rtol=1e-5)

    def bad(y, x):
        raise RuntimeError("user operator failed")

    with pytest.raises(pkg.ArpackException):
        pkg.symeigs(pkg.ArpackSimpleFunctionOp(bad, n), 3, ncv=20)


def test_normal_function_op_arpack_svd_example(pkg):
    """test/interface_simple.jl:170-220: svds(ArpackNormalFunctionOp(av!, atv!, 500, 100), 4; ncv=10) gives ARPACK's dsvd
    values; here av/atv are dense device products of the same operator."""
    import torch

    m, n = 500, 100
    h, k = 1.0 / (m + 1), 1.0 / (n + 1)
    A = np.zeros((m, n))
    for j in range(n):
        t = (j + 1) * k
        sv = (np.arange(m) + 1) * h
        A[: j + 1, j] = k * sv[: j + 1] * (t - 1)
        A[j + 1:, j] = k * t * (sv[j + 1:] - 1)
    Ad = torch.as_tensor(A, device="cuda")
    op = pkg.ArpackNormalFunctionOp(lambda y, x: torch.mv(Ad, x, out=y), lambda y, x: torch.mv(Ad.T, x, out=y), m, n)
    U, s, V = pkg.svds(op, 4, ncv=10)
    gold = sorted(GOLD["svds_arpack_example_500x100"]["values"], reverse=True)
    assert np.allclose(s, gold, rtol=1e-10)
    assert np.max(np.abs(A @ V - U * s)) <= 1e-12
    assert np.max(np.abs(U.T @ U - np.eye(4))) <= 1e-12 and np.max(np.abs(V.T @ V - np.eye(4))) <= 1e-12
    opt = pkg.ArpackNormalFunctionOp(lambda y, x: torch.mv(Ad.T, x, out=y), lambda y, x: torch.mv(Ad, x, out=y), n, m)
    U2, s2, V2 = pkg.svds(opt, 4, ncv=10)                               # m <= n: the A A' side
    assert np.allclose(s2, gold, rtol=1e-10)
