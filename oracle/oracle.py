"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

The oracle is a CPU restatement of GenericArpack.jl's implicitly restarted Lanczos path
(see oracle/oracle.hpp for the file:line map).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product package
genericarpack.jl_b200 never does.

dtype codes: 0 = Float64, 1 = ComplexF64, 2 = Double64 (numpy float64 arrays with a trailing
axis of 2 = (hi, lo), the memory layout of DoubleFloats.Double64).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

F64, C64, DD, F32 = 0, 1, 2, 3   # F32: Float32 vectors with a Float64 factorisation (symeigs(Float32, Float64, ...))
WHICH = {"LM": 0, "SM": 1, "LA": 2, "SA": 3, "BE": 4}


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (g++ -fopenmp)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "oracle.hpp", "dd.hpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [i32, i64, i32, vp, vp, vp, i64, i64]
        L.oracle_create_general.restype = vp
        L.oracle_create_general.argtypes = [i32, i64, i32, vp, vp, vp, vp, vp, vp]
        L.oracle_bx.argtypes = [vp, vp, vp]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_size.restype = i64
        L.oracle_size.argtypes = [vp]
        L.oracle_ptr.restype = vp
        L.oracle_ptr.argtypes = [vp, i32]
        L.oracle_iseed.restype = ctypes.POINTER(i64)
        L.oracle_iseed.argtypes = [vp]
        L.oracle_iparam.restype = ctypes.POINTER(i64)
        L.oracle_iparam.argtypes = [vp]
        L.oracle_get_rnorm.argtypes = [vp, vp]
        L.oracle_set_rnorm.argtypes = [vp, vp]
        L.oracle_get_stats.argtypes = [vp, vp, vp]
        L.oracle_reset_stats.argtypes = [vp]
        L.oracle_opx.argtypes = [vp, vp, vp]
        L.oracle_dgetv0.argtypes = [vp, i32, i32, i32]
        L.oracle_dsaitr.argtypes = [vp, i32, i32]
        L.oracle_dsapps.argtypes = [vp, i32, i32, vp]
        L.oracle_dsaupd.argtypes = [vp, i32, i32, vp, i32, i32]
        L.oracle_dseupd.argtypes = [vp, i32, vp, vp, i32, i32, vp]
        L.oracle_dlarnv.argtypes = [i32, vp, i64, vp]
        L.oracle_dlarnv_f32.restype = i64
        L.oracle_dlarnv_f32.argtypes = [vp, i64, vp]
        L.oracle_lcg_table.argtypes = [vp]
        L.oracle_norm2.argtypes = [i32, vp, i64, vp]
        L.oracle_dstqrb.argtypes = [i32, i32, vp, vp, vp]
        L.oracle_dsteqr.argtypes = [i32, i32, vp, vp, vp]
        L.oracle_plane_rotation.argtypes = [dbl, dbl, vp]
        L.oracle_dsconv.argtypes = [i32, vp, vp, dbl]
        L.oracle_dsortr.argtypes = [i32, i32, i32, vp, vp]
        L.oracle_dsgets.argtypes = [i32, i32, i32, i32, vp, vp, vp]
        L.oracle_dseigt.argtypes = [dbl, i32, vp, i32, vp, vp, vp]
        L.oracle_dd_op.argtypes = [i32, vp, vp, vp]
        L.oracle_num_threads.restype = i32
        L.oracle_set_num_threads.argtypes = [i32]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def np_dtype(dtype):
    return np.complex128 if dtype == C64 else (np.float32 if dtype == F32 else np.float64)


def elem_shape(dtype):
    return (2,) if dtype == DD else ()


def to_dtype(x, dtype):
    """Convert a numpy VECTOR into the oracle's element layout.  Double64: a (n, 2) float64 array
    is taken as (hi, lo) pairs, a (n,) array is widened to (x, 0)."""
    x = np.asarray(x)
    if dtype == DD:
        if x.ndim == 2 and x.shape[-1] == 2:
            return np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(x.shape + (2,), dtype=np.float64)
        out[..., 0] = x
        return out
    return np.ascontiguousarray(x, dtype=np_dtype(dtype))


# ------------------------------------------------------------------ stand-alone pieces
def dlarnv(n, iseed=(1, 3, 5, 7), dtype=F64):
    """_dlarnv_idist_2! : returns (x, new_iseed)."""
    seed = np.array(iseed, dtype=np.int64)
    x = np.zeros((n,) + elem_shape(dtype), dtype=np_dtype(dtype))
    lib().oracle_dlarnv(dtype, _p(seed), n, _p(x))
    return x, tuple(int(s) for s in seed)


def dlarnv_f32(n, iseed=(1, 3, 5, 7)):
    """_dlarnv_idist_2! with RT = Float32 (binary32 conversion, incl. the rv == 1 reseed branch,
    src/arpack-blas.jl:520-525): returns (x float32, new_iseed, number of reseeds taken)."""
    seed = np.array(iseed, dtype=np.int64)
    x = np.zeros(n, dtype=np.float32)
    k = lib().oracle_dlarnv_f32(_p(seed), n, _p(x))
    return x, tuple(int(s) for s in seed), int(k)


def lcg_table():
    t = np.zeros((128, 4), dtype=np.int32)
    lib().oracle_lcg_table(_p(t))
    return t


def norm2(x, dtype=None):
    x = np.ascontiguousarray(x)
    if dtype is None:
        dtype = C64 if np.iscomplexobj(x) else F64
    n = x.shape[0]
    out = np.zeros(2)
    lib().oracle_norm2(dtype, _p(x), n, _p(out))
    return out if dtype == DD else float(out[0])


def dstqrb(d, e, dtype=F64):
    """dstqrb!: returns (eigenvalues ascending, last row of eigenvector matrix)."""
    d = np.array(d, dtype=np.float64, copy=True)
    e = np.array(e, dtype=np.float64, copy=True)
    n = d.shape[0]
    if e.shape[0] < max(n - 1, 1):
        e = np.concatenate([e, np.zeros((1,) + e.shape[1:])])
    z = np.zeros_like(d)
    rc = lib().oracle_dstqrb(dtype, n, _p(d), _p(e), _p(z))
    if rc:
        raise RuntimeError("dstqrb failed to converge")
    return d, z


def dsteqr(d, e, dtype=F64):
    """simple_dsteqr!: returns (eigenvalues ascending, eigenvector matrix n x n)."""
    d = np.array(d, dtype=np.float64, copy=True)
    e = np.array(e, dtype=np.float64, copy=True)
    n = d.shape[0]
    if e.shape[0] < max(n - 1, 1):
        e = np.concatenate([e, np.zeros((1,) + e.shape[1:])])
    Z = np.zeros((n, n) + elem_shape(dtype), dtype=np.float64)  # column-major: Z[c, r]
    rc = lib().oracle_dsteqr(dtype, n, _p(d), _p(e), _p(Z))
    if rc:
        raise RuntimeError("dsteqr failed to converge")
    return d, np.swapaxes(Z, 0, 1)


def plane_rotation(f, g):
    out = np.zeros(3)
    lib().oracle_plane_rotation(f, g, _p(out))
    return tuple(out)


def dsconv(ritz, bounds, tol):
    ritz = np.ascontiguousarray(ritz, dtype=np.float64)
    bounds = np.ascontiguousarray(bounds, dtype=np.float64)
    return lib().oracle_dsconv(len(ritz), _p(ritz), _p(bounds), tol)


def dsortr(which, apply, x1, x2):
    x1 = np.array(x1, dtype=np.float64)
    x2 = np.array(x2, dtype=np.float64)
    lib().oracle_dsortr(WHICH[which], int(apply), len(x1), _p(x1), _p(x2))
    return x1, x2


def dsgets(ishift, which, kev, np_, ritz, bounds):
    ritz = np.array(ritz, dtype=np.float64)
    bounds = np.array(bounds, dtype=np.float64)
    shifts = np.zeros(max(np_, 1))
    lib().oracle_dsgets(ishift, WHICH[which], kev, np_, _p(ritz), _p(bounds), _p(shifts))
    return ritz, bounds, shifts[:np_]


def dseigt(rnorm, H):
    """H: n x 2 (col 0 sub-diagonal from row 2, col 1 diagonal). returns (eig, bounds)."""
    H = np.asfortranarray(H, dtype=np.float64)
    n = H.shape[0]
    eig = np.zeros(n)
    bounds = np.zeros(n)
    workl = np.zeros(3 * n)
    lib().oracle_dseigt(rnorm, n, _p(H), n, _p(eig), _p(bounds), _p(workl))
    return eig, bounds


def dd_op(op, a, b=(0.0, 0.0)):
    a = np.array(a, dtype=np.float64)
    b = np.array(b, dtype=np.float64)
    out = np.zeros(2)
    lib().oracle_dd_op({"add": 0, "sub": 1, "mul": 2, "div": 3, "sqrt": 4}[op], _p(a), _p(b), _p(out))
    return out


# ------------------------------------------------------------------ handle-based solver
class Oracle:
    """One problem instance: operator + the work arrays symeigs would allocate
    (src/interface.jl:245-251).  A is a scipy.sparse matrix (any format); it is stored as
    full CSR with int64 rowptr / int32 col.  normal=True builds ArpackNormalOp; B != None builds
    ArpackSymmetricGeneralizedOp(A, factorize(B), B) (mode 2, bmat = :G; B positive definite); b_val
    optionally gives B's stored entries in the element layout (Double64 (hi, lo) pairs that are not doubles)."""

    def __init__(self, A, ncv, dtype=None, normal=False, B=None, b_val=None):
        import scipy.sparse as sp

        A = sp.csr_matrix(A)
        A.sort_indices()
        if dtype is None:
            dtype = C64 if (np.iscomplexobj(A.data) or (B is not None and np.iscomplexobj(sp.csr_matrix(B).data))) else F64
        self.dtype = dtype
        self.ncv = ncv
        self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.col = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.val = to_dtype(A.data, dtype)
        m, n = A.shape
        L = lib()
        if B is not None:
            assert m == n and not normal
            B = sp.csr_matrix(B)
            B.sort_indices()
            assert B.shape == (n, n)
            self.b_rowptr = np.ascontiguousarray(B.indptr, dtype=np.int64)
            self.b_col = np.ascontiguousarray(B.indices, dtype=np.int32)
            self.b_val = to_dtype(B.data if b_val is None else b_val, dtype)
            self.h = L.oracle_create_general(dtype, n, ncv, _p(self.rowptr), _p(self.col), _p(self.val),
                                             _p(self.b_rowptr), _p(self.b_col), _p(self.b_val))
        elif normal:
            self.h = L.oracle_create(dtype, 0, ncv, _p(self.rowptr), _p(self.col), _p(self.val), m, n)
        else:
            assert m == n
            self.h = L.oracle_create(dtype, n, ncv, _p(self.rowptr), _p(self.col), _p(self.val), 0, 0)
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.n = int(L.oracle_size(self.h))
        es = elem_shape(dtype)
        nd = np_dtype(dtype)
        rdt = np.float64
        n_, res = self.n, es

        def view(what, shape, dt):
            ptr = L.oracle_ptr(self.h, what)
            cnt = int(np.prod(shape))
            buf = (ctypes.c_char * (cnt * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt).reshape(shape)

        # V is column-major n x ncv -> exposed as V[c, r(, 2)]
        self.Vt = view(0, (ncv, n_) + res, nd)
        self.resid = view(1, (n_,) + res, nd)
        self.workd = view(2, (3 * n_,) + res, nd)
        self.workl = view(3, (ncv * ncv + 8 * ncv,) + res, rdt)
        self.iseed = np.ctypeslib.as_array(L.oracle_iseed(self.h), shape=(4,))
        self.iparam = np.ctypeslib.as_array(L.oracle_iparam(self.h), shape=(11,))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # --- scalars
    @property
    def rnorm(self):
        out = np.zeros(2)
        lib().oracle_get_rnorm(self.h, _p(out))
        return out.copy() if self.dtype == DD else float(out[0])

    @rnorm.setter
    def rnorm(self, v):
        a = np.zeros(2)
        a[: np.size(v)] = np.ravel(v)
        lib().oracle_set_rnorm(self.h, _p(a))

    def stats(self):
        c = np.zeros(5, dtype=np.int64)
        t = np.zeros(11)
        lib().oracle_get_stats(self.h, _p(c), _p(t))
        names_c = ["nopx", "nbx", "nrorth", "nitref", "nrstrt"]
        names_t = ["taupd", "taup2", "taitr", "teigt", "tgets", "tapps", "tconv", "tmvopx", "tgetv0", "titref", "trvec"]
        d = {k: int(v) for k, v in zip(names_c, c)}
        d.update({k: float(v) for k, v in zip(names_t, t)})
        return d

    def reset_stats(self):
        lib().oracle_reset_stats(self.h)

    # --- H and Q views inside workl (src/dsaupd.jl:1016-1023)
    @property
    def H(self):
        """ncv x 2 (col 0 sub-diagonal, col 1 diagonal); a view."""
        ncv = self.ncv
        return np.swapaxes(self.workl[: 2 * ncv].reshape((2, ncv) + elem_shape(self.dtype)), 0, 1)

    @property
    def Q(self):
        ncv = self.ncv
        return np.swapaxes(self.workl[4 * ncv : 4 * ncv + ncv * ncv].reshape((ncv, ncv) + elem_shape(self.dtype)), 0, 1)

    # --- routines
    def opx(self, x):
        x = to_dtype(x, self.dtype)
        y = np.zeros_like(x)
        lib().oracle_opx(self.h, _p(y), _p(x))
        return y

    def bx(self, x):
        x = to_dtype(x, self.dtype)
        y = np.zeros_like(x)
        lib().oracle_bx(self.h, _p(y), _p(x))
        return y

    def dgetv0(self, itry=1, initv=False, j=1):
        return lib().oracle_dgetv0(self.h, itry, int(initv), j)

    def dsaitr(self, k, np_):
        return lib().oracle_dsaitr(self.h, k, np_)

    def dsapps(self, kev, np_, shifts):
        s = np.ascontiguousarray(shifts, dtype=np.float64)
        lib().oracle_dsapps(self.h, kev, np_, _p(s))

    def _tol(self, tol):
        a = np.zeros(2)
        a[: np.size(tol)] = np.ravel(tol)
        return a

    def dsaupd(self, nev, which="LM", tol=0.0, maxiter=300, v0=None):
        initv = 0
        if v0 is not None:
            self.resid[...] = to_dtype(v0, self.dtype)
            initv = 1
        t = self._tol(tol)
        return lib().oracle_dsaupd(self.h, WHICH[which], nev, _p(t), maxiter, initv)

    def dseupd(self, nev, which="LM", tol=0.0, rvec=True):
        nconv = int(self.iparam[4])
        es = elem_shape(self.dtype)
        d = np.zeros((max(nconv, 1),) + es, dtype=np.float64)
        Zt = np.zeros((max(nconv, 1), self.n) + es, dtype=np_dtype(self.dtype))
        t = self._tol(tol)
        ierr = lib().oracle_dseupd(self.h, int(rvec), _p(d), _p(Zt), WHICH[which], nev, _p(t))
        return ierr, d[:nconv], np.swapaxes(Zt[:nconv], 0, 1) if rvec else None

    def symeigs(self, nev, which="LM", tol=0.0, maxiter=None, v0=None, ritzvec=True):
        """src/interface.jl:215-290.  returns dict(values, vectors, info, nconv, niter, bounds, stats)."""
        if maxiter is None:
            maxiter = max(300, int(np.ceil(np.sqrt(self.n))))
        info = self.dsaupd(nev, which, tol, maxiter, v0)
        out = {"info": info, "niter": int(self.iparam[2]), "nconv": int(self.iparam[4])}
        if info not in (0, 1):
            out.update(values=None, vectors=None, ierr=None, stats=self.stats())
            return out
        ierr, d, Z = self.dseupd(nev, which, tol, ritzvec)
        ncv = self.ncv
        out.update(values=d, vectors=Z, ierr=ierr, bounds=self.workl[5 * ncv : 5 * ncv + out["nconv"]].copy(),
                   stats=self.stats())
        return out


def set_arith(mode):
    """"plain" (default: plain binary64 sweeps, x87 norm) or "exact" (error-free dots, the reference's double-double
    norm variant, fixed update order: the trajectory then does not depend on blocking or thread count)."""
    lib().oracle_set_arith(1 if mode in (1, "exact") else 0)


def get_arith():
    return "exact" if lib().oracle_get_arith() == 1 else "plain"


class arith:
    """Context manager: `with oracle.arith("exact"): ...`"""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get_arith()
        set_arith(self.mode)

    def __exit__(self, *a):
        set_arith(self.prev)


def set_num_threads(n):
    """OpenMP threads of the oracle's sweeps / SpMV (torchrun exports OMP_NUM_THREADS=1 to every rank; the
    reference arm of bench.py runs on rank 0 alone and may use all host cores)."""
    lib().oracle_set_num_threads(int(n))


def num_threads():
    return int(lib().oracle_num_threads())
