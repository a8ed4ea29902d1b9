"""CPU oracle (test infrastructure only) for the svds post-processing of GenericArpack.jl:
numpy restatement of `svds` (src/interface.jl:312-337), `_adjust_eigenvalues_to_singular_values!`
(src/idonow_ops.jl:255-260), `eigenvecs_to_singvecs!` (:295-333) and `orth!` (:279-293) with its
Householder pieces `_dlarfg!`, `_dgeqr2!`, `_dlarf_left!`, `dorg2r!`
(src/arpack-blas-qr.jl:48-74, 137-185, 264-291, 496-520).  Real element types (Float64).

The eigen-solve underneath is the C++ oracle (oracle.py: symeigs on ArpackNormalOp).
Pinned by the reference's golden singular values (test/interface_simple.jl:59-63, 100-105, 218-220)
and by its check_svd properties (test/utility.jl:244-253): tests/test_oracle_svd.py.
"""
import numpy as np

from . import oracle as O


def _dlapy2(x, y):  # src/arpack-blas.jl:159-169
    xa, ya = abs(x), abs(y)
    w, z = max(xa, ya), min(xa, ya)
    if z == 0.0:
        return w
    return w * np.sqrt(1.0 + (z / w) ** 2)


def dlarfg(x):
    """_dlarfg!(x) (src/arpack-blas-qr.jl:137-185): x[0] <- beta, x[1:] <- v, returns tau."""
    alpha = x[0]
    if len(x) <= 1:
        return 0.0
    xnorm = O.norm2(x[1:])
    tau = 0.0
    beta = alpha
    if xnorm != 0.0:
        beta = -np.copysign(_dlapy2(alpha, xnorm), alpha)
        safmin = np.finfo(np.float64).tiny / (np.finfo(np.float64).eps / 2)
        knt = 0
        if abs(beta) < safmin:
            rsafmn = 1.0 / safmin
            while knt < 20:
                knt += 1
                x[1:] *= rsafmn
                beta *= rsafmn
                alpha *= rsafmn
                if abs(beta) >= safmin:
                    break
            xnorm = O.norm2(x[1:])
            beta = -np.copysign(_dlapy2(alpha, xnorm), alpha)
        tau = (beta - alpha) / beta
        x[1:] *= 1.0 / (alpha - beta)
        for _ in range(knt):
            beta *= safmin
    x[0] = beta
    return tau


def dlarf_left(v, tau, C):
    """_dlarf_left! (src/arpack-blas-qr.jl:264-291): C <- (I - tau v v') C."""
    if tau == 0.0 or C.shape[1] == 0:
        return
    nz = np.nonzero(v)[0]
    lastv = int(nz[-1]) + 1 if len(nz) else 0
    if lastv == 0:
        return
    w = C[:lastv, :].T @ v[:lastv]
    C[:lastv, :] -= tau * np.outer(v[:lastv], w)


def dgeqr2(A):
    """_dgeqr2! (src/arpack-blas-qr.jl:48-74), in place; returns tau."""
    m, n = A.shape
    k = min(m, n)
    tau = np.zeros(k)
    for i in range(k):
        tau[i] = dlarfg(A[i:, i])
        if i < n - 1:
            aii = A[i, i]
            A[i, i] = 1.0
            dlarf_left(A[i:, i].copy(), tau[i], A[i:, i + 1:])
            A[i, i] = aii
    return tau


def dorg2r(A, tau):
    """dorg2r! (src/arpack-blas-qr.jl:496-520), in place (k = n here)."""
    m, n = A.shape
    k = len(tau)
    for i in range(k - 1, -1, -1):
        if i < n - 1:
            A[i, i] = 1.0
            dlarf_left(A[i:, i].copy(), tau[i], A[i:, i + 1:])
        if i < m - 1:
            A[i + 1:, i] *= -tau[i]
        A[i, i] = 1.0 - tau[i]
        A[:i, i] = 0.0
    return A


def orth(X):
    """orth! (src/idonow_ops.jl:279-293): Householder QR, Q formed explicitly, columns re-signed so
    that R's diagonal is positive (the unique thin-QR Q factor)."""
    X = np.array(X, dtype=np.float64, order="F")
    tau = dgeqr2(X)
    rsign = np.array([1.0 if X[i, i] == 0 else np.sign(X[i, i]) for i in range(X.shape[1])])
    dorg2r(X, tau)
    X *= rsign[None, :]
    return X


def eigenvecs_to_singvecs(A, Z):
    """eigenvecs_to_singvecs! (src/idonow_ops.jl:295-333).  A: scipy sparse m x n; Z: eigenvectors of
    A'A (m > n) or AA' (m <= n).  Returns (U, V, norms)."""
    m, n = A.shape
    eps = np.finfo(np.float64).eps
    warnthresh, errorthresh = eps ** (2.0 / 3.0), eps / 2
    k = Z.shape[1]
    norms = np.zeros(k)
    if m > n:   # At_side == :left
        V = np.array(Z, dtype=np.float64)
        U = np.zeros((m, k))
        for i in range(k):
            U[:, i] = A @ V[:, i]
            nrm = float(O.norm2(U[:, i]))
            norms[i] = nrm
            if nrm <= errorthresh:
                raise ArithmeticError("encountered sigma ~ %g when computing U subspace" % nrm)
            U[:, i] *= 1.0 / nrm      # _scale_from_to(nrm, 1, .) in the safe range
        U = orth(U)
    else:
        U = np.array(Z, dtype=np.float64)
        V = np.zeros((n, k))
        for i in range(k):
            V[:, i] = A.T @ U[:, i]
            nrm = float(O.norm2(V[:, i]))
            norms[i] = nrm
            if nrm <= errorthresh:
                raise ArithmeticError("encountered sigma ~ %g when computing V subspace" % nrm)
            V[:, i] *= 1.0 / nrm
        V = orth(V)
    return U, V, norms, warnthresh


def svds(A, k, which="LM", ncv=None, dtype=None, **kw):
    """svds(A, k; which) (src/interface.jl:312-337) -> (U, S, V).  dtype = oracle.F32: svds(Float32, op, k) -- the
    eigen-solve runs on Float32 vectors with tol = eps(Float32)/2 (what TF = Float32 gives in the reference); the
    vector post-processing below is carried out in Float64 on the widened eigenvectors."""
    import scipy.sparse as sp

    A = sp.csr_matrix(A)
    m, n = A.shape
    size = m if m <= n else n
    arwhich = {"LM": "LM", "LA": "LM", "SM": "SA", "SA": "SA", "BE": "BE"}[which]   # src/idonow_ops.jl:240-253
    if ncv is None:
        ncv = min(size, max(2 * k, 20))
    if dtype == O.F32:
        kw.setdefault("tol", 2.0 ** -24)
    ora = O.Oracle(A, ncv, normal=True, dtype=dtype)
    res = ora.symeigs(k, which=arwhich, **kw)
    if res["info"] not in (0, 1) or res["values"] is None:
        raise ArithmeticError("symmetric aupd gave error code info=%d" % res["info"])
    vals = np.sqrt(np.maximum(res["values"], 0.0))                        # src/idonow_ops.jl:255-260
    Z = res["vectors"]
    # dsortr(:SM / :LM, ...) on magnitudes = descending / ascending order (values are >= 0)
    if which in ("LM", "LA"):
        order = np.argsort(-vals, kind="stable")
    elif which in ("SA", "SM"):
        order = np.argsort(vals, kind="stable")
    else:
        order = np.arange(len(vals))
    S = vals[order]
    if Z is None or Z.shape[1] == 0:
        return np.zeros((m, 0)), S, np.zeros((n, 0)), res
    U, V, norms, _ = eigenvecs_to_singvecs(A, Z[:, order])
    return U, S, V, res


def svd_residuals(A, U, s, V):
    """svd_residuals (src/interface.jl:388-407): ||A V[:,i] - U[:,i] s[i]||."""
    return np.array([float(np.linalg.norm(A @ V[:, i] - U[:, i] * s[i])) for i in range(len(s))])
