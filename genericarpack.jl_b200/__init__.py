"""genericarpack.jl_b200 -- B200-native implicitly restarted Lanczos for GenericArpack.jl.

Python host-side mirror of the reference's user interface for the accelerated path
(/root/reference/src/interface.jl, src/idonow_ops.jl): ``eigs`` / ``symeigs`` / ``hermeigs`` /
``svds`` with the same keyword names, defaults and error behaviour, driving the hand-written
CUDA library ``libgarpack_b200.so`` through its C ABI (include/garpack_b200.h) with ctypes.
Julia is not installed in the build image; INTEGRATION.md shows the ``ccall`` shim a Julia
maintainer adds instead of this file.

There is NO CPU fallback: importing works without a GPU (so the symbol table can be checked),
but every compute entry point raises if the library or a CUDA device is missing.

Element types (``dtype=``): ``"f64"`` (Float64), ``"c64"`` (ComplexF64), ``"dd"`` (Double64,
numpy float64 arrays with a trailing axis of 2 = (hi, lo), the memory layout of Julia's
DoubleFloats.Double64).

The directory name contains a dot, so load it with ``__graft_entry__.load_package()`` (or
importlib) under the module name ``genericarpack_jl_b200``.
"""
import ctypes
import math
import os

import numpy as np

__all__ = [
    "ArpackException", "ArpackSimpleOp", "ArpackSimpleFunctionOp", "ArpackNormalFunctionOp", "ArpackNormalOp", "ArpackSymmetricGeneralizedOp", "B200ArpackOp", "ArpackEigen", "ArpackStats",
    "Context", "comm_unique_id", "DDArr", "as_dd", "dd_to_float", "to_elem", "eigs", "symeigs", "hermeigs", "svds", "lib", "build", "GA_F64", "GA_C64", "GA_DD", "GA_F32",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("GA_B200_LIB") or os.path.join(_HERE, "libgarpack_b200.so")   # GA_B200_LIB: development A/B builds only
_LIB = None

GA_F64, GA_C64, GA_DD, GA_F32 = 0, 1, 2, 3
GA_ERR_CUDA, GA_ERR_ARG, GA_ERR_NOOP, GA_ERR_COMM, GA_ERR_NUMERIC, GA_ERR_SOLVE = -100, -101, -102, -103, -104, -105
_DTYPES = {"f64": GA_F64, "float64": GA_F64, "c64": GA_C64, "complex128": GA_C64, "dd": GA_DD, "double64": GA_DD,
           "f32": GA_F32, "float32": GA_F32,   # Float32 vectors, Float64 factorisation: symeigs(Float32, Float64, ...)
           GA_F64: GA_F64, GA_C64: GA_C64, GA_DD: GA_DD, GA_F32: GA_F32}
_WHICH = {"LM": 0, "SM": 1, "LA": 2, "SA": 3, "BE": 4}
GA_OPT_DOT_MODE, GA_DOT_EXACT, GA_DOT_PLAIN = 1, 0, 1   # include/garpack_b200.h

# every symbol include/garpack_b200.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "ga_create", "ga_destroy", "ga_trim_cache", "ga_set_option", "ga_get_option", "ga_last_error", "ga_device_sm_count", "ga_comm_unique_id", "ga_comm_init", "ga_comm_ipc_export", "ga_comm_ipc_import",
    "ga_set_csr", "ga_set_csr_dist", "ga_set_normal_csr", "ga_set_normal_csr_dist", "ga_opx", "ga_upload_resid", "ga_download_resid", "ga_upload_v",
    "ga_download_v", "ga_getv0", "ga_lanczos", "ga_apply_shifts", "ga_ritz_vectors", "ga_svd_vectors", "ga_norm2_resid", "ga_bnorm_resid",
    "ga_get_stats", "ga_reset_stats", "ga_timer_start", "ga_timer_stop", "ga_profile", "ga_profile_read", "ga_profile_gs_launches", "ga_profile_gs_times", "ga_bench_opx", "ga_debug_trace", "ga_debug_spmv_check", "ga_tune_spmv", "ga_set_op_callback", "ga_set_normal_op_callbacks", "ga_set_generalized_csr", "ga_set_generalized_csr_dist", "ga_set_solve_callback", "ga_bx", "ga_get_solve_stats", "ga_host_tridiag_eig", "ga_host_ritz_and_bounds", "ga_host_sort_pair", "ga_host_select_shifts", "ga_host_count_converged", "ga_host_chase_shifts", "ga_symeigs", "ga_symeigs_begin", "ga_symeigs_iterate", "ga_symeigs_end",
]


class ArpackException(Exception):  # src/interface.jl:125-127
    pass


class ga_stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("nopx", "nbx", "nrorth", "nitref", "nrstrt")] + [
        (n, ctypes.c_double)
        for n in ("taupd", "taup2", "taitr", "teigt", "tgets", "tapps", "tconv", "tmvopx", "tmvbx", "tgetv0",
                  "titref", "trvec")
    ]


class ArpackStats(dict):
    """Counters/timers of the reference's ArpackStats (src/macros.jl:173-198)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def build(force=False):
    """Compile libgarpack_b200.so in-tree (nvcc, sm_100a)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_ga_b200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force)


def lib():
    """The loaded CUDA library.  Raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise RuntimeError(
                "libgarpack_b200.so is missing: run `python genericarpack.jl_b200/build.py` "
                "(there is no CPU fallback for the Lanczos hot path)")
        L = ctypes.CDLL(_SO)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.ga_create.argtypes = [ctypes.POINTER(vp), i32, i64, i64, i64, i32, i32, i32, i32]
        L.ga_destroy.argtypes = [vp]
        L.ga_destroy.restype = None
        L.ga_last_error.argtypes = [vp]
        L.ga_last_error.restype = ctypes.c_char_p
        L.ga_device_sm_count.argtypes = [vp]
        L.ga_comm_unique_id.argtypes = [vp]
        L.ga_comm_init.argtypes = [vp, vp]
        L.ga_set_csr.argtypes = [vp, vp, vp, vp]
        L.ga_set_option.argtypes = [vp, i32, i64]
        L.ga_get_option.argtypes = [vp, i32, vp]
        L.ga_comm_ipc_export.argtypes = [vp, vp]
        L.ga_comm_ipc_import.argtypes = [vp, vp]
        L.ga_set_csr_dist.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
        L.ga_set_normal_csr.argtypes = [vp, i64, i64, vp, vp, vp]
        L.ga_set_normal_csr_dist.argtypes = [vp, i64, vp, vp, vp, i32, vp, vp, vp, vp]
        L.ga_opx.argtypes = [vp, vp, vp]
        L.ga_set_generalized_csr.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32]
        L.ga_set_generalized_csr_dist.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
        L.ga_set_solve_callback.argtypes = [vp, vp, vp]
        L.ga_set_op_callback.argtypes = [vp, vp, vp]
        L.ga_set_normal_op_callbacks.argtypes = [vp, i64, i64, vp, vp, vp]
        L.ga_bx.argtypes = [vp, vp, vp]
        L.ga_get_solve_stats.argtypes = [vp, vp, vp, vp]
        L.ga_upload_resid.argtypes = [vp, vp]
        L.ga_download_resid.argtypes = [vp, vp]
        L.ga_upload_v.argtypes = [vp, i32, i32, vp, i64]
        L.ga_download_v.argtypes = [vp, i32, i32, vp, i64]
        L.ga_getv0.argtypes = [vp, vp, i32, i32, vp]
        L.ga_lanczos.argtypes = [vp, i32, i32, vp, i32, vp, vp]
        L.ga_apply_shifts.argtypes = [vp, i32, i32, vp, i32, vp, vp]
        L.ga_ritz_vectors.argtypes = [vp, i32, vp, i32, vp, i64]
        L.ga_norm2_resid.argtypes = [vp, vp]
        L.ga_bnorm_resid.argtypes = [vp, vp]
        L.ga_svd_vectors.argtypes = [vp, i32, vp, vp, i64, vp]
        L.ga_get_stats.argtypes = [vp, ctypes.POINTER(ga_stats)]
        L.ga_reset_stats.argtypes = [vp]
        L.ga_timer_start.argtypes = [vp]
        L.ga_timer_stop.argtypes = [vp, vp]
        L.ga_profile.argtypes = [vp, i32]
        L.ga_profile_read.argtypes = [vp, vp, vp, vp, vp]
        L.ga_profile_gs_launches.argtypes = [vp, vp]
        L.ga_profile_gs_times.argtypes = [vp, vp, i64]
        L.ga_bench_opx.argtypes = [vp, i32, vp]
        L.ga_debug_trace.argtypes = [vp, vp]
        L.ga_debug_spmv_check.argtypes = [vp, vp]
        L.ga_tune_spmv.argtypes = [vp, i32, i32]
        L.ga_symeigs.argtypes = [vp, i32, i32, vp, i32, vp, i32, vp, vp, vp, i64, vp]
        L.ga_host_tridiag_eig.argtypes = [i32, i32, vp, vp, vp, i32, i32]
        L.ga_host_ritz_and_bounds.argtypes = [i32, vp, i32, vp, i32, vp, vp, vp]
        L.ga_host_sort_pair.argtypes = [i32, i32, i32, vp, vp]
        L.ga_host_select_shifts.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp]
        L.ga_host_count_converged.argtypes = [i32, i32, vp, vp, vp]
        L.ga_host_chase_shifts.argtypes = [i32, i32, i32, vp, vp, i32, vp, i32]
        L.ga_symeigs_begin.argtypes = [vp, i32, i32, vp, i32, vp]
        L.ga_symeigs_iterate.argtypes = [vp, i32]
        L.ga_symeigs_end.argtypes = [vp, i32, vp, vp, vp, i64, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _np_dtype(dt):
    return np.complex128 if dt == GA_C64 else (np.float32 if dt == GA_F32 else np.float64)


def _eshape(dt):
    return (2,) if dt == GA_DD else ()


class DDArr(np.ndarray):
    """float64 ndarray whose LAST axis (length 2) is the (hi, lo) pair of a Double64."""


def as_dd(x):
    """Mark an array that already has the trailing (hi, lo) axis as Double64 data."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.shape[-1] == 2
    return x.view(DDArr)


def to_elem(x, dt):
    """numpy array -> element layout of dtype dt.  Double64: arrays marked with as_dd() are
    taken as (hi, lo) pairs, plain arrays are widened to (x, 0)."""
    if dt == GA_DD:
        if isinstance(x, DDArr):
            return np.ascontiguousarray(x).view(DDArr)
        x = np.asarray(x, dtype=np.float64)
        out = np.zeros(x.shape + (2,), dtype=np.float64)
        out[..., 0] = x
        return out.view(DDArr)
    return np.ascontiguousarray(np.asarray(x), dtype=_np_dtype(dt))


def dd_to_float(x):
    """Double64 array -> nearest float64."""
    x = np.asarray(x)
    return x[..., 0] + x[..., 1]


# ------------------------------------------------------------------ operators (src/idonow_ops.jl)
class B200ArpackOp:
    """ArpackOp whose OP*x runs on the B200: a full (both triangles) CSR matrix.

    Mirrors the required methods of ``abstract type ArpackOp`` (src/idonow_ops.jl:50-56, 360-368):
    ``size(op)`` (problem dimension), ``arpack_mode(op)``, ``bmat(op)``, ``opx!``.
    """

    arpack_mode = 1
    bmat = "I"
    normal = False

    def __init__(self, A, dtype=None):
        import scipy.sparse as sp

        if isinstance(A, tuple):  # (rowptr, col, val, shape)
            rowptr, col, val, shape = A
            self.shape = tuple(shape)
            self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
            self.col = np.ascontiguousarray(col, dtype=np.int32)
            data = np.asarray(val)
        else:
            A = sp.csr_matrix(A)
            A.sort_indices()
            self.shape = A.shape
            self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
            self.col = np.ascontiguousarray(A.indices, dtype=np.int32)
            data = A.data
        if dtype is None:
            dtype = GA_C64 if np.iscomplexobj(data) else (GA_F32 if np.asarray(data).dtype == np.float32 else GA_F64)
        self.dtype = _DTYPES[dtype]
        self.val = to_elem(data, self.dtype)

    def size(self):
        return self.shape[0]

    def is_arpack_mode_valid_for_op(self, mode):
        return mode == 1


class ArpackSimpleOp(B200ArpackOp):  # src/idonow_ops.jl:115-123
    pass


class ArpackNormalOp(B200ArpackOp):  # src/idonow_ops.jl:216-277
    normal = True

    def size(self):
        m, n = self.shape
        return m if m <= n else n

    @property
    def At_side(self):
        m, n = self.shape
        return "right" if m <= n else "left"


_OP_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                          ctypes.c_void_p)    # ga_op_fn


class _DevVec:
    """a device vector handed to an operator callback, exported through __cuda_array_interface__ so that
    torch.as_tensor(v, device="cuda") / cupy.asarray(v) wrap it without a copy"""

    def __init__(self, ptr, n, dt):
        shape, typestr = {GA_F64: ((n,), "<f8"), GA_C64: ((n,), "<c16"), GA_DD: ((n, 2), "<f8"), GA_F32: ((n,), "<f4")}[dt]
        self.ptr, self.n = int(ptr), int(n)
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3,     # torch refuses the read-only flag
                                         "strides": None}


def _wrap_op_fn(F, dt):
    """ga_op_fn around a Python callable F(y, x): x and y arrive as torch CUDA tensors that alias the library's
    device vectors, and the call runs under the library's stream so that everything F launches is enqueued there."""

    def thunk(user, x_ptr, y_ptr, nx, ny, stream):
        try:
            import torch

            with torch.cuda.stream(torch.cuda.ExternalStream(int(stream))):
                x = torch.as_tensor(_DevVec(x_ptr, nx, dt), device="cuda")   # treat as read-only
                y = torch.as_tensor(_DevVec(y_ptr, ny, dt), device="cuda")
                F(y, x)
            return 0
        except Exception:  # an exception must not unwind through the C frames
            import traceback

            traceback.print_exc()
            return 1

    return _OP_FN(thunk)


class ArpackSimpleFunctionOp(B200ArpackOp):
    """ArpackSimpleFunctionOp(F, n) (src/idonow_ops.jl:360-368): opx!(y, op, x) = op.F(y, x) with a matrix-free F.
    Here F(y, x) receives torch CUDA tensors aliasing the library's device vectors (zero copy) and must write y
    in place with work enqueued on the current torch stream (= the library's stream); the Lanczos loop stays on
    the device, F is called once per OP*x while the steps are being enqueued."""

    def __init__(self, F, n, dtype="f64"):
        self.F, self.shape, self.dtype = F, (int(n), int(n)), _DTYPES[dtype]
        self._cb = _wrap_op_fn(F, self.dtype)          # keeps the ctypes thunk alive as long as the op

    def size(self):
        return self.shape[0]


class ArpackNormalFunctionOp(B200ArpackOp):
    """ArpackNormalFunctionOp(av, atv, m, n) (src/idonow_ops.jl:406-431): OP = A'A (m > n) or AA' through the two
    callbacks av(y, x): y = A x and atv(y, x): y = A' x on device tensors (see ArpackSimpleFunctionOp)."""

    normal = True

    def __init__(self, av, atv, m, n, dtype="f64"):
        self.av, self.atv, self.shape, self.dtype = av, atv, (int(m), int(n)), _DTYPES[dtype]
        self._cb = (_wrap_op_fn(av, self.dtype), _wrap_op_fn(atv, self.dtype))

    def size(self):
        m, n = self.shape
        return m if m <= n else n

    @property
    def At_side(self):
        m, n = self.shape
        return "right" if m <= n else "left"


class ArpackSymmetricGeneralizedOp(B200ArpackOp):
    """ArpackSymmetricGeneralizedOp(A, S, B) (src/idonow_ops.jl:462-492): A x = lambda B x in mode 2 with
    bmat = :G.  A and B are full (both triangles) sparse matrices, B Hermitian positive definite.  The
    reference's factorisation object S is replaced by a preconditioned conjugate-gradient solve with B on
    the device (relative residual ``cg_tol``, default 8 eps; at most ``cg_maxit`` iterations)."""

    arpack_mode = 2
    bmat = "G"

    def __init__(self, A, B, dtype=None, cg_tol=None, cg_maxit=0, sharded=False):
        """sharded=True: A and B are this rank's rows (n_local x n, global columns) for a multi-GPU Context, which also
        needs the common halo plan of dist.shard_generalized."""
        import scipy.sparse as sp

        if dtype is None:
            cplx = any(np.iscomplexobj(M[2] if isinstance(M, tuple) else sp.csr_matrix(M).data) for M in (A, B))
            dtype = GA_C64 if cplx else GA_F64
        super().__init__(A, dtype=dtype)
        self.Bop = B200ArpackOp(B, dtype=dtype)
        if self.Bop.shape != self.shape or (self.shape[0] != self.shape[1] and not sharded):
            raise ValueError("A and B must be square matrices of the same size")
        self.cg_tol, self.cg_maxit = cg_tol, int(cg_maxit)

    def is_arpack_mode_valid_for_op(self, mode):
        return mode == 2


# ------------------------------------------------------------------ low-level context
class Context:
    """One device problem instance = the arrays symeigs allocates (src/interface.jl:245-251) on
    the GPU, plus the operator.  Thin, 1:1 over the C ABI; used by the tests and by eigs."""

    def __init__(self, op, ncv, device=0, rank=0, world_size=1, n_global=None, row_offset=0, plan=None, dots=None):
        """world_size > 1: `op` holds this rank's rows (global columns) and `plan` is the halo plan
        of dist.halo_plan(...) for them.  dots: "exact" (default; Dot2 accumulators, results independent of the row
        partition) or "plain" (one FMA per product) for the Float64 sweeps -- ga_set_option(GA_OPT_DOT_MODE)."""
        L = lib()
        self.op = op
        self.dtype = op.dtype
        self.ncv = int(ncv)
        if world_size == 1:
            self.n = int(op.size())
        elif op.normal:       # op holds this rank's rows of the tall orientation T; the plan knows the Lanczos block
            self.n = int(plan["row_range"][1] - plan["row_range"][0])
        else:
            self.n = int(len(op.rowptr) - 1)
        self.n_global = int(n_global) if n_global is not None else self.n
        if world_size > 1 and plan is not None:
            from . import dist as _dist
            want = _dist.padded_block(self.n_global, world_size, self.dtype)
            if int(plan["n_pad"]) != want:
                raise ValueError("the halo plan was built for a local block padded to %d rows, this context pads to %d: "
                                 "pass dtype= to dist.halo_plan / dist.shard_normal (Float32 blocks are padded to 512 rows)"
                                 % (int(plan["n_pad"]), want))
        h = ctypes.c_void_p()
        rc = L.ga_create(ctypes.byref(h), self.dtype, self.n_global, self.n, int(row_offset), self.ncv, device, rank,
                         world_size)
        if rc != 0 or not h:
            raise RuntimeError("ga_create failed with code %d (is a CUDA device present? there is no CPU fallback)" % rc)
        self.h = h
        if isinstance(op, (ArpackSimpleFunctionOp, ArpackNormalFunctionOp)):
            if world_size != 1:
                raise ValueError("function operators run on one GPU")
            if isinstance(op, ArpackSimpleFunctionOp):
                self._ck(L.ga_set_op_callback(self.h, ctypes.cast(op._cb, ctypes.c_void_p), None))
            else:
                m, n = op.shape
                self._ck(L.ga_set_normal_op_callbacks(self.h, m, n, ctypes.cast(op._cb[0], ctypes.c_void_p),
                                                      ctypes.cast(op._cb[1], ctypes.c_void_p), None))
        elif isinstance(op, ArpackSymmetricGeneralizedOp):
            tol = None if op.cg_tol is None else np.array([float(op.cg_tol), 0.0])
            if world_size > 1:   # this rank's rows of A and B, one halo plan for both (dist.shard_generalized)
                if plan is None or "col_local_b" not in plan:
                    raise ValueError("a multi-GPU generalised problem needs the common halo plan of A and B (dist.shard_generalized)")
                self.plan = plan
                self._ck(L.ga_set_generalized_csr_dist(self.h, _p(op.rowptr), _p(plan["col_local"]), _p(op.val), _p(op.Bop.rowptr),
                                                       _p(plan["col_local_b"]), _p(op.Bop.val), _p(tol), op.cg_maxit,
                                                       int(plan["nhalo"]), _p(plan["recv_counts"]), _p(plan["send_counts"]),
                                                       _p(plan["send_idx"]), _p(plan["send_dest_off"])))
            else:
                self._ck(L.ga_set_generalized_csr(self.h, _p(op.rowptr), _p(op.col), _p(op.val), _p(op.Bop.rowptr),
                                                  _p(op.Bop.col), _p(op.Bop.val), _p(tol), op.cg_maxit))
        elif op.normal and world_size > 1:
            if plan is None:
                raise ValueError("a multi-GPU context needs the halo plan (dist.halo_plan / dist.shard_normal)")
            self.plan = plan
            self.mt_local = int(len(op.rowptr) - 1)
            self._ck(L.ga_set_normal_csr_dist(self.h, self.mt_local, _p(op.rowptr), _p(plan["col_local"]), _p(op.val),
                                              int(plan["nhalo"]), _p(plan["recv_counts"]), _p(plan["send_counts"]),
                                              _p(plan["send_idx"]), _p(plan["send_dest_off"])))
        elif op.normal:
            m, n = op.shape
            self._ck(L.ga_set_normal_csr(self.h, m, n, _p(op.rowptr), _p(op.col), _p(op.val)))
        elif world_size > 1:
            if plan is None:
                raise ValueError("a multi-GPU context needs the halo plan (dist.halo_plan)")
            self.plan = plan
            self._ck(L.ga_set_csr_dist(self.h, _p(op.rowptr), _p(plan["col_local"]), _p(op.val), int(plan["nhalo"]),
                                       _p(plan["recv_counts"]), _p(plan["send_counts"]), _p(plan["send_idx"]),
                                       _p(plan["send_dest_off"])))
        else:
            self._ck(L.ga_set_csr(self.h, _p(op.rowptr), _p(op.col), _p(op.val)))
        self.iseed = np.array([1, 3, 5, 7], dtype=np.int64)  # src/macros.jl:291
        if dots is not None:
            self.set_dots(dots)

    def set_dots(self, mode):
        """GA_OPT_DOT_MODE: "exact" (Dot2, the default) or "plain" (FMA accumulation) in the Float64 sweeps."""
        if mode not in ("exact", "plain"):
            raise ValueError('dots must be "exact" or "plain"')
        self._ck(lib().ga_set_option(self.h, GA_OPT_DOT_MODE, GA_DOT_PLAIN if mode == "plain" else GA_DOT_EXACT))

    def get_dots(self):
        v = ctypes.c_longlong(0)
        self._ck(lib().ga_get_option(self.h, GA_OPT_DOT_MODE, ctypes.byref(v)))
        return "plain" if v.value == GA_DOT_PLAIN else "exact"

    def close(self):
        if getattr(self, "h", None):
            lib().ga_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            msg = lib().ga_last_error(self.h)
            raise ArpackException("garpack_b200 error %d: %s" % (rc, msg.decode() if msg else ""))
        return rc

    def _real(self, v=0.0):
        a = np.zeros(2)
        a[: np.size(v)] = np.ravel(v)
        return a

    def _real_out(self, a):
        return a.copy() if self.dtype == GA_DD else float(a[0])

    @property
    def sm_count(self):
        return int(lib().ga_device_sm_count(self.h))

    # --- data movement
    def upload_resid(self, x):
        x = to_elem(x, self.dtype)
        assert x.shape[0] == self.n
        self._ck(lib().ga_upload_resid(self.h, _p(x)))

    def download_resid(self, out=None):
        """out: optional preallocated (e.g. pinned) host array of the result's shape."""
        shape = (self.n,) + _eshape(self.dtype)
        x = np.zeros(shape, dtype=_np_dtype(self.dtype)) if out is None else out
        assert x.shape == shape and x.flags.c_contiguous
        self._ck(lib().ga_download_resid(self.h, _p(x)))
        return x

    def upload_v(self, V, col0=0):
        """V: n x k (column j of the basis = V[:, j])."""
        k = np.shape(V)[1]
        Vt = np.ascontiguousarray(np.swapaxes(to_elem(V, self.dtype), 0, 1))  # [col][row](,2)
        self._ck(lib().ga_upload_v(self.h, col0, k, _p(Vt), self.n))

    def download_v(self, col0=0, ncols=None, out=None):
        """out: optional preallocated (e.g. pinned) host array of shape (ncols, n[, 2]) = V columns as rows."""
        ncols = self.ncv - col0 if ncols is None else ncols
        shape = (ncols, self.n) + _eshape(self.dtype)
        Vt = np.zeros(shape, dtype=_np_dtype(self.dtype)) if out is None else out
        assert Vt.shape == shape and Vt.flags.c_contiguous
        self._ck(lib().ga_download_v(self.h, col0, ncols, _p(Vt), self.n))
        return np.swapaxes(Vt, 0, 1)

    def opx(self, x):
        x = to_elem(x, self.dtype)
        y = np.zeros_like(x)
        self._ck(lib().ga_opx(self.h, _p(x), _p(y)))
        return y

    def bx(self, x):
        """bx!(y, op, x) of the generalised problem: y = B x."""
        x = to_elem(x, self.dtype)
        y = np.zeros_like(x)
        self._ck(lib().ga_bx(self.h, _p(x), _p(y)))
        return y

    def solve_stats(self):
        """(solves with B, CG iterations in total, largest final relative residual)."""
        a, b, r = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_double(0.0)
        lib().ga_get_solve_stats(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(r))
        return a.value, b.value, r.value

    # --- hot path
    def getv0(self, initv=False, j=1):
        """dgetv0! : returns (ierr, rnorm)."""
        rn = np.zeros(2)
        rc = lib().ga_getv0(self.h, _p(self.iseed), int(initv), int(j), _p(rn))
        self._ck(rc, allow=(-1,))
        return rc, self._real_out(rn)

    def lanczos(self, k, np_, H, rnorm):
        """dsaitr! : H is an (ncv, 2[, 2]) array (col 0 sub-diagonal, col 1 diagonal), updated in
        place for rows k..k+np-1.  returns (info, rnorm)."""
        Hf = np.ascontiguousarray(np.swapaxes(H, 0, 1))  # [col][row]
        rn = self._real(rnorm)
        rc = lib().ga_lanczos(self.h, int(k), int(np_), _p(Hf), H.shape[0], _p(rn), _p(self.iseed))
        if rc < 0:
            self._ck(rc)
        H[...] = np.swapaxes(Hf, 0, 1)
        return rc, self._real_out(rn)

    def apply_shifts(self, kev, np_, Q, beta):
        """device tail of dsapps!: Q is (kplusp, kplusp[, 2]) real. returns rnorm."""
        Qf = np.ascontiguousarray(np.swapaxes(Q, 0, 1))
        b = self._real(beta)
        out = np.zeros(2)
        self._ck(lib().ga_apply_shifts(self.h, int(kev), int(np_), _p(Qf), Q.shape[0], _p(b), _p(out)))
        return self._real_out(out)

    def ritz_vectors(self, nconv, Qh, want_z=True):
        Qf = np.ascontiguousarray(np.swapaxes(Qh, 0, 1))
        Zt = np.zeros((max(nconv, 1), self.n) + _eshape(self.dtype), dtype=_np_dtype(self.dtype)) if want_z else None
        self._ck(lib().ga_ritz_vectors(self.h, int(nconv), _p(Qf), Qh.shape[0], _p(Zt), self.n))
        return np.swapaxes(Zt[:nconv], 0, 1) if want_z else None

    def svd_vectors(self, order):
        """eigenvecs_to_singvecs! + orth! on the device (ga_svd_vectors): the other side's singular
        vectors for the eigenvectors held in basis columns `order`.  Returns (rc, W, norms); W has
        this rank's rows of the other side (all of them on one GPU)."""
        nvec = len(order)
        if getattr(self, "mt_local", None) is not None:
            rows = self.mt_local
        else:
            m, n = self.op.shape
            rows = m if self.op.At_side == "left" else n
        Wt = np.zeros((nvec, rows) + _eshape(self.dtype), dtype=_np_dtype(self.dtype))
        norms = np.zeros((nvec, 2))
        perm = np.ascontiguousarray(order, dtype=np.int32)
        rc = lib().ga_svd_vectors(self.h, nvec, _p(perm), _p(Wt), max(rows, 1), _p(norms))
        nr = norms.reshape(-1)[:nvec].copy() if self.dtype != GA_DD else norms[:, 0].copy()
        return rc, np.swapaxes(Wt, 0, 1), nr

    def norm2_resid(self):
        out = np.zeros(2)
        self._ck(lib().ga_norm2_resid(self.h, _p(out)))
        return self._real_out(out)

    def bnorm_resid(self):
        """The norm dsaup2! recomputes after a restart: 2-norm (bmat = :I) or sqrt|r' B r| (bmat = :G)."""
        out = np.zeros(2)
        self._ck(lib().ga_bnorm_resid(self.h, _p(out)))
        return self._real_out(out)

    def stats(self):
        s = ga_stats()
        lib().ga_get_stats(self.h, ctypes.byref(s))
        return ArpackStats({n: getattr(s, n) for n, _ in ga_stats._fields_})

    def reset_stats(self):
        lib().ga_reset_stats(self.h)

    # --- multi-GPU plumbing (rendezvous is the caller's: torch.distributed / MPI)
    def comm_init(self, id128):
        buf = np.ascontiguousarray(id128, dtype=np.uint8)
        assert buf.size == 128
        self._ck(lib().ga_comm_init(self.h, _p(buf)))

    def ipc_export(self):
        buf = np.zeros(128, dtype=np.uint8)
        self._ck(lib().ga_comm_ipc_export(self.h, _p(buf)))
        return buf

    def ipc_import(self, all_handles):
        buf = np.ascontiguousarray(all_handles, dtype=np.uint8)
        self._ck(lib().ga_comm_ipc_import(self.h, _p(buf)))

    # --- measurement hooks
    def timer_start(self):
        self._ck(lib().ga_timer_start(self.h))

    def timer_stop(self):
        ms = ctypes.c_double(0.0)
        self._ck(lib().ga_timer_stop(self.h, ctypes.byref(ms)))
        return ms.value

    def debug_trace(self, read=False):
        if not read:
            self._ck(lib().ga_debug_trace(self.h, None))
            return None
        out = np.zeros(64, dtype=np.uint64)
        self._ck(lib().ga_debug_trace(self.h, _p(out)))
        return out

    def tune_spmv(self, groups, stages):
        self._ck(lib().ga_tune_spmv(self.h, int(groups), int(stages)))

    def debug_spmv_check(self, read=False):
        if not read:
            self._ck(lib().ga_debug_spmv_check(self.h, None))
            return None
        out = np.zeros(128, dtype=np.uint32)
        self._ck(lib().ga_debug_spmv_check(self.h, _p(out)))
        return out

    def bench_opx(self, reps=20):
        """average milliseconds of one OP*x on device-resident vectors (measurement hook)"""
        ms = ctypes.c_double(0.0)
        self._ck(lib().ga_bench_opx(self.h, int(reps), ctypes.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        self._ck(lib().ga_profile(self.h, int(enable)))

    def profile_gs_times(self, n):
        out = np.zeros(int(n), dtype=np.float32)
        self._ck(lib().ga_profile_gs_times(self.h, _p(out), int(n)))
        return out

    def profile_read(self):
        ms = ctypes.c_double(0.0)
        runs, units, launches = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        self._ck(lib().ga_profile_read(self.h, ctypes.byref(ms), ctypes.byref(runs), ctypes.byref(units),
                                       ctypes.byref(launches)))
        ngs = ctypes.c_int64(0)
        self._ck(lib().ga_profile_gs_launches(self.h, ctypes.byref(ngs)))
        return {"gs_ms": ms.value, "gs_runs": runs.value, "gs_units": units.value, "launches": launches.value,
                "gs_launches": ngs.value}

    # --- whole solve
    def symeigs_begin(self, nev, which="LM", tol=0.0, maxiter=300, v0=None):
        t = self._real(tol)
        v = None if v0 is None else to_elem(v0, self.dtype)
        return lib().ga_symeigs_begin(self.h, _WHICH.get(str(which).lstrip(":"), -1), int(nev), _p(t), int(maxiter), _p(v))

    def symeigs_iterate(self, max_restarts=-1):
        return lib().ga_symeigs_iterate(self.h, int(max_restarts))

    def symeigs_end(self, nev, ritzvec=True):
        es = _eshape(self.dtype)
        values = np.zeros((nev,) + es)
        bounds = np.zeros((nev,) + es)
        Zt = np.zeros((nev, self.n) + es, dtype=_np_dtype(self.dtype)) if ritzvec else None
        iparam = np.zeros(11, dtype=np.int64)
        rc = lib().ga_symeigs_end(self.h, int(ritzvec), _p(values), _p(bounds), _p(Zt), self.n, _p(iparam))
        nconv = int(iparam[4])
        vecs = np.swapaxes(Zt[:nconv], 0, 1) if ritzvec else None
        return rc, values[:nconv], bounds[:nconv], vecs, iparam


def trim_cache():
    """ga_trim_cache: give the device blocks the library keeps for the next context back to the driver."""
    L = lib()
    L.ga_trim_cache.restype = None
    L.ga_trim_cache()


def comm_unique_id(buf=None):
    """128-byte NCCL unique id (rank 0 creates it, the caller broadcasts it)."""
    if buf is None:
        buf = np.zeros(128, dtype=np.uint8)
    rc = lib().ga_comm_unique_id(_p(buf))
    if rc != 0:
        raise ArpackException("ga_comm_unique_id failed (%d): is libnccl.so.2 loadable?" % rc)
    return buf


class ArpackEigen:
    """Result of symeigs (src/interface.jl:173-213): iterates as (values, vectors)."""

    def __init__(self, which, iparam, values, vectors, bounds, op, ctx, stats, info):
        self.which, self.iparam, self.values, self.vectors = which, iparam, values, vectors
        self.bounds, self.op, self.ctx, self.stats, self.info = bounds, op, ctx, stats, info

    def __iter__(self):
        yield self.values
        yield self.vectors

    @property
    def V(self):
        return self.ctx.download_v()

    @property
    def resid(self):
        return self.ctx.download_resid()


def _type_arg(x):
    """element type given like the reference's leading TV / TF arguments, or None if x is not a type"""
    if isinstance(x, str):
        return _DTYPES.get(x.lower())
    if isinstance(x, (type, np.dtype)):
        try:
            return {np.dtype(np.float32): GA_F32, np.dtype(np.float64): GA_F64, np.dtype(np.complex128): GA_C64}.get(np.dtype(x))
        except TypeError:
            return None
    return None


def _as_op(A, dtype=None, normal=False, B=None):
    if isinstance(A, B200ArpackOp):
        return A
    if B is not None:
        return ArpackSymmetricGeneralizedOp(A, B, dtype=dtype)
    return (ArpackNormalOp if normal else ArpackSimpleOp)(A, dtype=dtype)


def symeigs(A, *args, which="LM", maxiter=None, ncv=None, v0=None, stats=None, tol=0.0, ritzvec=True,
            failonmaxiter=True, dtype=None, device=0, keep_context=False, dots=None):
    """symeigs(TV, TF, op, nev; which, maxiter, ncv, v0, stats, tol, ritzvec, failonmaxiter)
    (src/interface.jl:215-290) on the B200.  ``symeigs(A, nev)``: scipy sparse matrix (full storage) or a
    B200ArpackOp; ``symeigs(A, B, nev)``: the generalised problem A x = lambda B x through
    ArpackSymmetricGeneralizedOp (src/interface.jl:143-163).  Returns an ArpackEigen; values ascending like
    the reference.  Like the reference the element types may lead the argument list: ``symeigs(np.float32, A, nev)``
    (= eigs(Float32, A, k)) and ``symeigs(np.float32, np.float64, A, nev)`` (= eigs(Float32, Float64, A, k), the mixed
    precision of src/interface.jl:26-34; the factorisation type of Float32 vectors is always Float64 here)."""
    lead = []
    narrow_tf = False
    allargs = (A,) + tuple(args)
    while allargs and _type_arg(allargs[0]) is not None and len(lead) < 2:
        lead.append(_type_arg(allargs[0]))
        allargs = allargs[1:]
    if lead:
        if not allargs:
            raise TypeError("symeigs(TV, [TF,] A, nev)")
        tv = lead[0]
        ok_tf = {GA_F64: (GA_F64,), GA_C64: (GA_F64,), GA_DD: (GA_DD,), GA_F32: (GA_F64, GA_F32)}[tv]
        if len(lead) == 2 and lead[1] not in ok_tf:
            raise ValueError("unsupported factorisation type TF for these vectors")
        if dtype is not None and _DTYPES[dtype] != tv:
            raise ValueError("conflicting element types")
        dtype = tv
        # eigs(Float32, A, k) means TV = TF = Float32: the values come back as Float32 like the reference's
        # (the factorisation itself still runs in Float64 here)
        narrow_tf = tv == GA_F32 and (len(lead) == 1 or lead[1] == GA_F32)
        A, args = allargs[0], allargs[1:]
    if len(args) == 1:
        B, nev = None, args[0]
    elif len(args) == 2:
        B, nev = args
    else:
        raise TypeError("symeigs(A, nev) or symeigs(A, B, nev)")
    nev = int(nev)
    op = _as_op(A, dtype, B=B)
    n = op.size()
    if maxiter is None:
        maxiter = max(300, int(math.ceil(math.sqrt(n))))            # :217
    if ncv is None:
        ncv = min(n, max(2 * nev, 20))                              # :219
    if not ncv <= n:
        raise ValueError("ncv=%d must be at most n=%d" % (ncv, n))  # :231
    if not nev < ncv:
        raise ValueError("nev=%d must be strictly less than ncv=%d" % (nev, ncv))  # :232
    ctx = Context(op, ncv, device=device, dots=dots)
    try:
        rc = ctx.symeigs_begin(nev, which, tol, maxiter, v0)
        if rc != 0:
            raise ArpackException("symmetric aupd gave error code info=%d (%s)" % (rc, (lib().ga_last_error(ctx.h) or b"").decode()))
        rc = ctx.symeigs_iterate(-1)
        if rc != 99:
            raise ArpackException("symmetric aupd gave error code info=%d (%s)" % (rc, (lib().ga_last_error(ctx.h) or b"").decode()))
        info, values, bounds, vecs, iparam = ctx.symeigs_end(nev, ritzvec)
        if info == 1 and failonmaxiter:                             # :269-270
            raise ArpackException("symmetric aupd hit maxiter=%d with %d < %d eigenvalues converged" % (maxiter, int(iparam[4]), nev))
        if info not in (0, 1):                                      # :271-273, :285-287
            raise ArpackException("symmetric aupd/eupd gave error code info=%d" % info)
        st = ctx.stats()
        if isinstance(stats, dict):
            stats.update(st)
        if narrow_tf:
            values, bounds = values.astype(np.float32), bounds.astype(np.float32)
        return ArpackEigen(which, iparam, values, vecs, bounds, op, ctx if keep_context else None, st, info)
    finally:
        if not keep_context:
            ctx.close()


def eigs(A, *args, **kw):
    """eigs(Symmetric(A) / Hermitian(A), k; kwargs...) (src/interface.jl:133-139) and
    eigs(Symmetric(A), Symmetric(B), k; kwargs...) (:143-147)."""
    return symeigs(A, *args, **kw)


def hermeigs(A, *args, **kw):  # src/interface.jl:158-168
    kw.setdefault("dtype", "c64")
    return symeigs(A, *args, **kw)


class SVD:
    """LinearAlgebra.SVD(U, S, Vt) as returned by the reference's svds (src/interface.jl:336);
    iterates as (U, S, V) like Julia's destructuring."""

    def __init__(self, U, S, Vt, einfo=None):
        self.U, self.S, self.Vt, self.einfo = U, S, Vt, einfo

    @property
    def V(self):
        return None if self.Vt is None else np.conj(self.Vt).T

    def __iter__(self):
        yield self.U
        yield self.S
        yield self.V


def svds(A, k, *, which="LM", dtype=None, **kw):
    """svds(A, k; which, kwargs...) (src/interface.jl:304-337) via ArpackNormalOp.

    The Lanczos iteration on A'A (m > n) / AA' (m <= n) runs on the device; so does
    eigenvecs_to_singvecs! + orth! (src/idonow_ops.jl:279-333, C ABI ga_svd_vectors): the other side's
    vectors A v_i / A' u_i, their norms, scaling and the orthonormalisation.  Returns SVD(U, S, V')
    with singular values largest first for which in (LM, LA), smallest first for (SA, SM)."""
    import warnings

    op = _as_op(A, dtype, normal=True)
    if which not in ("LM", "LA", "SM", "SA", "BE"):
        raise ValueError("which = %s is not valid for SVD of ArpackNormalOp" % which)     # :252
    arwhich = {"LM": "LM", "LA": "LM", "SM": "SA", "SA": "SA", "BE": "BE"}[which]         # :240-253
    ritzvec = kw.get("ritzvec", True)
    einfo = symeigs(op, k, which=arwhich, keep_context=True, **kw)
    ctx = einfo.ctx
    try:
        vals = np.sqrt(np.maximum(dd_to_float(einfo.values) if op.dtype == GA_DD else einfo.values, 0.0))  # :255-260
        if which in ("LM", "LA"):                                         # src/interface.jl:323-329
            order = np.argsort(-vals, kind="stable")
        elif which in ("SA", "SM"):
            order = np.argsort(vals, kind="stable")
        else:
            order = np.arange(len(vals))
        S = vals[order]
        m, n = op.shape
        if not ritzvec or einfo.vectors is None or einfo.vectors.shape[1] == 0:
            es = _eshape(op.dtype)
            return SVD(np.zeros((m, 0) + es), S, np.zeros((0, n) + es), einfo)
        nvec = einfo.vectors.shape[1]
        order = order[:nvec]
        Z = einfo.vectors[:, order]
        rc, W, nr = ctx.svd_vectors(order)
        eps = 4.930380657631324e-32 if op.dtype == GA_DD else (np.finfo(np.float32).eps if op.dtype == GA_F32 else np.finfo(np.float64).eps)
        side = "U" if op.At_side == "left" else "V"
        if rc == GA_ERR_NUMERIC or (rc == 0 and np.any(nr <= eps / 2)):   # check_norm (src/idonow_ops.jl:300-306)
            raise ArpackException("encountered sigma ~ %g when computing %s subspace" % (float(np.min(nr)), side))
        ctx._ck(rc)
        if np.any(nr <= eps ** (2.0 / 3.0)):
            warnings.warn("%s subspace may be inaccurate due to small singular value (sigma ~ %g)" % (side, float(np.min(nr))))
        U, V = (W, Z) if op.At_side == "left" else (Z, W)
        if op.dtype == GA_C64:
            Vt = np.conj(V).T
        elif op.dtype == GA_DD:
            Vt = np.swapaxes(V, 0, 1)
        else:
            Vt = V.T
        return SVD(U, S, Vt, einfo)
    finally:
        ctx.close()
        einfo.ctx = None


def svd_residuals(A, U, s, V):
    """svd_residuals (src/interface.jl:388-407): ||A V[:,i] - U[:,i] s[i]|| (host helper, Float64/ComplexF64)."""
    return np.array([float(np.linalg.norm(A @ V[:, i] - U[:, i] * s[i])) for i in range(len(s))])
