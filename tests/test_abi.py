"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/garpack_b200.h declares; the Python mirror has the reference's entry points; and without a
GPU the product path fails loudly instead of falling back to a CPU implementation."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "garpack_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ga_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    pkg.build()
    L = pkg.lib()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libgarpack_b200.so does not export %s" % n
    assert sorted(pkg.ABI_SYMBOLS) == names


def test_python_mirror_surface(pkg):
    """names of the reference's interface for this path (src/interface.jl, src/idonow_ops.jl)"""
    for name in ("eigs", "symeigs", "hermeigs", "svds", "ArpackSimpleOp", "ArpackNormalOp", "ArpackException",
                 "ArpackEigen", "ArpackStats"):
        assert hasattr(pkg, name)
    op = pkg.ArpackSimpleOp(sp.identity(7).tocsr())
    assert op.size() == 7 and op.arpack_mode == 1 and op.bmat == "I" and op.is_arpack_mode_valid_for_op(1)
    nop = pkg.ArpackNormalOp(sp.random(9, 4, density=0.5, format="csr", random_state=np.random.RandomState(0)))
    assert nop.size() == 4 and nop.At_side == "left"
    with pytest.raises(ValueError):
        pkg.symeigs(op, 3, ncv=8)   # ncv > n (src/interface.jl:231)
    with pytest.raises(ValueError):
        pkg.symeigs(op, 5, ncv=5)   # nev >= ncv (src/interface.jl:232)


def test_python_mirror_types_and_function_ops(pkg):
    """leading element-type arguments like the reference's symeigs(TV, TF, op, nev) (src/interface.jl:215), parsed
    before any device work, and the function-operator types of src/idonow_ops.jl:360-431"""
    assert pkg._type_arg(np.float32) == pkg.GA_F32 and pkg._type_arg("Float32".lower()) == pkg.GA_F32
    assert pkg._type_arg(np.float64) == pkg.GA_F64 and pkg._type_arg(np.complex128) == pkg.GA_C64
    assert pkg._type_arg("dd") == pkg.GA_DD and pkg._type_arg(3) is None and pkg._type_arg(np.zeros(2)) is None
    A = sp.identity(30, format="csr")
    with pytest.raises(ValueError):
        pkg.eigs(np.float64, np.float32, A, 3)              # TF narrower than TV
    with pytest.raises(ValueError):
        pkg.eigs(np.float32, A, 3, dtype="f64")             # conflicting element types
    with pytest.raises(TypeError):
        pkg.eigs(np.float32)
    assert pkg.ArpackSimpleOp(A.astype(np.float32)).dtype == pkg.GA_F32
    assert pkg.ArpackSimpleOp(A, dtype="f32").val.dtype == np.float32
    fop = pkg.ArpackSimpleFunctionOp(lambda y, x: None, 12)
    assert fop.size() == 12 and fop.arpack_mode == 1 and fop.bmat == "I" and not fop.normal
    nop = pkg.ArpackNormalFunctionOp(lambda y, x: None, lambda y, x: None, 20, 8, dtype="c64")
    assert nop.size() == 8 and nop.At_side == "left" and nop.normal and nop.dtype == pkg.GA_C64
    gop = pkg.ArpackSymmetricGeneralizedOp(A, 2 * A)
    assert gop.arpack_mode == 2 and gop.bmat == "G" and gop.is_arpack_mode_valid_for_op(2) and not gop.is_arpack_mode_valid_for_op(1)
    v = pkg._DevVec(4096, 5, pkg.GA_C64).__cuda_array_interface__
    assert v["shape"] == (5,) and v["typestr"] == "<c16" and v["data"] == (4096, False)


def test_header_is_plain_c_and_example_links(pkg, tmp_path):
    """include/garpack_b200.h is a C header (C99, -pedantic -Werror) and a plain-C caller links against the library;
    without a device the example reports GA_ERR_CUDA and computes nothing (exit 0), with one it prints eigenvalues."""
    import shutil
    import subprocess

    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("no gcc")
    pkg.build()
    exe = str(tmp_path / "c_abi_example")
    libdir = os.path.join(ROOT, "genericarpack.jl_b200")
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", libdir, "-lgarpack_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    assert "GA_ERR_CUDA" in out.stdout or "lambda[2]" in out.stdout


def test_no_cpu_fallback(pkg):
    """Without a CUDA device ga_create must fail (GA_ERR_CUDA); eigs must raise, not compute."""
    L = pkg.lib()
    h = ctypes.c_void_p()
    rc = L.ga_create(ctypes.byref(h), 0, 10, 10, 0, 4, 0, 0, 1)
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = os.path.exists("/dev/nvidia0")
    if has_gpu:
        assert rc == 0
        L.ga_destroy(h)
    else:
        assert rc == -100 and not h
        with pytest.raises(RuntimeError):
            pkg.eigs(sp.identity(30).tocsr(), 3)


def test_product_never_imports_oracle():
    """the product package may not reference oracle/ (only tests/, smoke() and bench.py may)"""
    pk = os.path.join(ROOT, "genericarpack.jl_b200")
    for dirpath, _, files in os.walk(pk):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt and \
                    "oracle.hpp" not in txt and "oracle/" not in txt.replace("oracle/ ", ""), f


def test_synthetic_generators(syn):
    rp, col, val, shape = syn.laplacian2d(7)
    A = syn.to_scipy((rp, col, val, shape)).toarray()
    assert shape == (49, 49) and len(col) == 5 * 49 - 4 * 7
    assert np.allclose(A, A.T) and np.all(np.diag(A) == 4)
    w = np.linalg.eigvalsh(A)
    assert np.allclose(np.sort(w)[-5:], syn.laplacian2d_eigs(7, 5))
    part = syn.laplacian2d(7, row_range=(14, 30))
    assert np.allclose(syn.to_scipy(part).toarray(), A[14:30])
    S = syn.to_scipy(syn.sprand_sym(200, 5.0, seed=1)).toarray()
    assert np.allclose(S, S.T) and 7 < (S != 0).sum() / 200 < 13
    Hm = syn.to_scipy(syn.hermitian_stencil(6)).toarray()
    assert np.allclose(Hm, Hm.conj().T)
    T = syn.to_scipy(syn.tall_sparse(300, 40, 5))
    assert T.shape == (300, 40)
    assert np.array_equal(syn.splitmix64(1234, 3), syn.splitmix64(1234, 5)[:3])
