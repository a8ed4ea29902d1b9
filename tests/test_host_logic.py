"""CPU tests of the PRODUCT's host-side control logic (genericarpack.jl_b200/csrc/host/*.hpp, reached through the
device-free ga_host_* entry points of the C ABI) against the oracle -- an independent restatement of the same
reference routines -- and against the reference's own analytic known answers:
dstqrb! / simple_dsteqr! (test/dstqrb.jl:87-114, test/dsteqr-simple.jl), dseigt! (test/runtests.jl:141-162),
dsortr / dsgets / dsconv (test/runtests.jl:113-140, test/dsgets_simple.jl) and the bulge chase of dsapps!
(src/dsapps.jl:609-805).  No GPU involved: this is the O(ncv^2) work BASELINE.json keeps on the host."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp

WHICH = {"LM": 0, "SM": 1, "LA": 2, "SA": 3, "BE": 4}


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_tridiag(pkg, d, e, full=False, dd=False):
    d = np.array(d, dtype=np.float64, copy=True)
    n = d.shape[0]
    e = np.array(e, dtype=np.float64, copy=True)
    if e.shape[0] < max(n, 1):
        e = np.concatenate([e, np.zeros((max(n, 1) - e.shape[0],) + e.shape[1:])])
    es = (2,) if dd else ()
    z = np.zeros(((n, n) if full else (n,)) + es)
    rc = pkg.lib().ga_host_tridiag_eig(int(dd), n, _p(d), _p(e), _p(z), int(full), n)
    assert rc == 0
    return d, (np.swapaxes(z, 0, 1) if full else z)       # full: column-major -> Z[row, col]


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 13, 17, 64, 99, 128, 129, 500, 1025])
def test_tridiag_laplacian_analytic_and_oracle(pkg, O, n):
    """test/dstqrb.jl:104-112 on the product's solver, and the same bits as the oracle's dstqrb!."""
    d, z = host_tridiag(pkg, 2 * np.ones(n), -np.ones(max(n - 1, 0)))
    k = np.arange(1, n + 1)
    lams = np.sort(-2 * np.cos(np.pi * k / (n + 1)) + 2.0)
    assert np.max(np.abs(lams - d)) <= 2 * np.sqrt(n) * np.finfo(float).eps
    zexp = np.abs(np.sqrt(2 / (n + 1)) * np.sin(n * k * np.pi / (n + 1)))
    assert np.linalg.norm(np.abs(z) - zexp) <= 2 * n * np.finfo(float).eps
    do, zo = O.dstqrb(2 * np.ones(n), -np.ones(max(n - 1, 0)))
    assert np.max(np.abs(d - do)) <= 4 * np.finfo(float).eps * 4 and np.max(np.abs(np.abs(z) - np.abs(zo))) <= 1e-13


@pytest.mark.parametrize("n", [2, 8, 16, 128, 3, 9, 127, 1001])
def test_tridiag_ones(pkg, n):
    """test/dstqrb.jl:87-103."""
    d, z = host_tridiag(pkg, np.zeros(n), np.ones(n - 1))
    k = np.arange(1, n + 1)
    if n % 2 == 0:
        lams = np.sort(2 * np.cos(np.pi * k / (n + 1)))
        assert np.max(np.abs(d - lams) / np.spacing(np.abs(lams))) <= 2 * n
    else:
        assert abs(d[n // 2]) <= 10 * np.finfo(float).eps
    zexp = np.abs(np.sqrt(2 / (n + 1)) * np.sin(n * k * np.pi / (n + 1)))
    assert np.linalg.norm(np.abs(z) - zexp) <= 2 * n * np.finfo(float).eps


def test_tridiag_random_full_vectors_and_splits(pkg, O):
    """simple_dsteqr! (full eigenvectors, used by dseupd's small eigenproblem) on random tridiagonals incl. exact zeros on
    the off-diagonal (the matrix splits, test/dsteqr-compare.jl) against LAPACK (scipy) and the oracle."""
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 12, 40, 64):
        d0 = rng.standard_normal(n)
        e0 = rng.standard_normal(max(n - 1, 0))
        if n > 6:
            e0[[2, n // 2]] = 0.0
        d, Z = host_tridiag(pkg, d0, e0, full=True)
        T = np.diag(d0) + np.diag(e0, 1) + np.diag(e0, -1)
        assert np.all(np.diff(d) >= 0)
        assert np.allclose(d, np.linalg.eigvalsh(T), rtol=0, atol=1e-13 * max(1.0, np.max(np.abs(d))))
        assert np.max(np.abs(T @ Z - Z * d)) <= 1e-13 * max(1.0, np.max(np.abs(d))) * n
        assert np.max(np.abs(Z.T @ Z - np.eye(n))) <= 1e-13 * n
        do, Zo = O.dsteqr(d0, e0)
        assert np.max(np.abs(d - do)) <= 1e-14 * max(1.0, np.max(np.abs(d)))
        dl, zl = host_tridiag(pkg, d0, e0)                      # last-row variant: same eigenvalues, last row of Z
        assert np.array_equal(dl, d) and np.max(np.abs(np.abs(zl) - np.abs(Z[-1, :]))) <= 1e-13


def test_tridiag_zero_matrix_and_scaling(pkg):
    """all-zero input (test/dstqrb.jl) and the dlascl-protected ranges (huge / tiny entries)."""
    d, z = host_tridiag(pkg, np.zeros(6), np.zeros(5))
    assert np.all(d == 0) and abs(np.linalg.norm(z) - 1) <= 1e-15
    for scale in (1e150, 1e-150, 1e-300):
        d, z = host_tridiag(pkg, scale * np.array([2.0, 1.0, 3.0]), scale * np.array([1.0, 1.0]))
        ref = np.linalg.eigvalsh(np.array([[2.0, 1, 0], [1, 1, 1], [0, 1, 3.0]]))
        assert np.allclose(d / scale, ref, rtol=1e-13)


def test_tridiag_double64(pkg, O):
    """the Double64 instantiation: tridiagonal Laplacian against the analytic spectrum in 200-bit arithmetic (1e-28) and
    the oracle's Double64 dstqrb!."""
    import mpmath as mp

    mp.mp.prec = 240
    n = 24
    d0 = np.zeros((n, 2)); d0[:, 0] = 2.0
    e0 = np.zeros((n - 1, 2)); e0[:, 0] = -1.0
    d, z = host_tridiag(pkg, d0, e0, dd=True)
    exact = sorted(2 - 2 * mp.cos(k * mp.pi / (n + 1)) for k in range(1, n + 1))
    got = [mp.mpf(float(v[0])) + mp.mpf(float(v[1])) for v in d]
    assert max(abs(g - x) for g, x in zip(got, exact)) < mp.mpf(10) ** -28
    do, zo = O.dstqrb(d0, e0, dtype=O.DD)
    assert np.max(np.abs((d[:, 0] - do[:, 0]) + (d[:, 1] - do[:, 1]))) <= 1e-29


def test_ritz_and_bounds_vs_oracle(pkg, O):
    """dseigt!: Ritz values of the Lanczos tridiagonal and rnorm * |last eigenvector components|."""
    rng = np.random.default_rng(2)
    for n in (1, 2, 6, 20, 40):
        H = np.zeros((n, 2))
        H[:, 1] = rng.standard_normal(n) * 3
        H[1:, 0] = np.abs(rng.standard_normal(n - 1)) + 0.1
        rnorm = 0.37
        Hf = np.asfortranarray(H)
        eig, bounds, work = np.zeros(n), np.zeros(n), np.zeros(3 * n)
        rn = np.array([rnorm, 0.0])
        assert pkg.lib().ga_host_ritz_and_bounds(0, _p(rn), n, _p(Hf), n, _p(eig), _p(bounds), _p(work)) == 0
        eo, bo = O.dseigt(rnorm, H)
        T = np.diag(H[:, 1]) + np.diag(H[1:, 0], 1) + np.diag(H[1:, 0], -1)
        w, Q = np.linalg.eigh(T)
        assert np.allclose(eig, w, atol=1e-13 * max(1, np.max(np.abs(w))))
        assert np.allclose(bounds, rnorm * np.abs(Q[-1, :]), atol=1e-13)
        assert np.max(np.abs(eig - eo)) <= 1e-14 * max(1, np.max(np.abs(w))) and np.max(np.abs(bounds - bo)) <= 1e-14


@pytest.mark.parametrize("which", ["LM", "SM", "LA", "SA"])
def test_sort_pair_vs_oracle(pkg, O, which):
    rng = np.random.default_rng(WHICH[which])
    for n in (0, 1, 2, 5, 17, 40):
        x1 = rng.standard_normal(n) * 10
        x1[rng.integers(0, max(n, 1), size=n // 4)] *= -1
        x2 = np.arange(float(n))
        a1, a2 = x1.copy(), x2.copy()
        assert pkg.lib().ga_host_sort_pair(0, WHICH[which], n, _p(a1), _p(a2)) == 0
        o1, o2 = O.dsortr(which, True, x1, x2)
        assert np.array_equal(a1, o1) and np.array_equal(a2, o2)
        key = np.abs(a1) if which in ("LM", "SM") else a1
        assert np.all(np.diff(key) >= 0) if which in ("LM", "LA") else np.all(np.diff(key) <= 0)   # wanted end LAST


@pytest.mark.parametrize("which", ["LM", "SM", "LA", "SA", "BE"])
def test_select_shifts_and_count_converged_vs_oracle(pkg, O, which):
    """dsgets (incl. the :BE swap) and dsconv, bit for bit against the oracle; test/dsgets_simple.jl's properties."""
    rng = np.random.default_rng(10 + WHICH[which])
    for kev, np_ in ((2, 3), (4, 6), (10, 30), (5, 1), (7, 0), (1, 5)):
        n = kev + np_
        ritz = rng.standard_normal(n) * 5
        bounds = np.abs(rng.standard_normal(n)) * 1e-3
        r, b, sh = ritz.copy(), bounds.copy(), np.zeros(max(np_, 1))
        assert pkg.lib().ga_host_select_shifts(0, 1, WHICH[which], kev, np_, _p(r), _p(b), _p(sh)) == 0
        ro, bo, so = O.dsgets(1, which, kev, np_, ritz, bounds)
        assert np.array_equal(r, ro) and np.array_equal(b, bo) and np.array_equal(sh[:np_], so)
        assert sorted(r) == sorted(ritz)                               # a permutation
        if np_ > 0:
            assert np.array_equal(sh[:np_], r[:np_]) and list(b[:np_]) == sorted(b[:np_], reverse=True)
        if which == "LM":
            assert np.min(np.abs(r[np_:])) >= np.max(np.abs(r[:np_])) if np_ > 0 else True   # wanted = largest magnitude
        tol = np.array([1e-3, 0.0])
        got = pkg.lib().ga_host_count_converged(0, n, _p(r), _p(b), _p(tol))
        assert got == O.dsconv(r, b, 1e-3)
    ones, zeros, tol = np.ones(10), np.zeros(10), np.array([1e-4, 0.0])
    assert pkg.lib().ga_host_count_converged(0, 10, _p(ones), _p(zeros), _p(tol)) == 10     # test/runtests.jl:113-121
    assert pkg.lib().ga_host_count_converged(0, 10, _p(ones), _p(ones), _p(tol)) == 0


def test_chase_shifts_vs_oracle_and_similarity(pkg, O):
    """the host part of dsapps!: np implicit QR shifts on the Lanczos tridiagonal.  Q is orthogonal, Q' T Q is the updated
    tridiagonal in its leading kev x kev block (plus the kev+1 sub-diagonal), and H, Q equal the oracle's dsapps!."""
    rng = np.random.default_rng(3)
    for kev, np_ in ((3, 4), (6, 6), (10, 30), (1, 7), (12, 1)):
        k = kev + np_
        A = sp.diags(np.arange(1.0, k + 5)).tocsr()
        o = O.Oracle(A, k)
        H = np.zeros((k, 2))
        H[:, 1] = rng.standard_normal(k) * 2
        H[1:, 0] = np.abs(rng.standard_normal(k - 1)) + 0.2
        T = np.diag(H[:, 1]) + np.diag(H[1:, 0], 1) + np.diag(H[1:, 0], -1)
        shifts = np.sort(np.linalg.eigvalsh(T))[:np_].copy()
        o.H[...] = H
        o.resid[:] = rng.standard_normal(o.n)
        o.dsapps(kev, np_, shifts)
        Hf = np.asfortranarray(H.copy())
        Q = np.zeros((k, k), order="F")
        assert pkg.lib().ga_host_chase_shifts(0, kev, np_, _p(shifts), _p(Hf), k, _p(Q), k) == 0
        assert np.max(np.abs(Q.T @ Q - np.eye(k))) <= 1e-13
        Tn = Q.T @ T @ Q
        lead = np.diag(Hf[:kev, 1]) + np.diag(Hf[1:kev, 0], 1) + np.diag(Hf[1:kev, 0], -1)
        assert np.max(np.abs(Tn[:kev, :kev] - lead)) <= 1e-12 * max(1.0, np.max(np.abs(T)))
        Ho, Qo = np.asarray(o.H), np.asarray(o.Q)
        assert np.array_equal(Hf[:kev + 1], Ho[:kev + 1])              # same operation order as the reference: same bits
        assert np.array_equal(Q[:, :kev + 1], Qo[:, :kev + 1])


def test_host_entry_points_reject_bad_arguments(pkg):
    L = pkg.lib()
    x = np.zeros(4)
    assert L.ga_host_tridiag_eig(0, 0, _p(x), _p(x), _p(x), 0, 1) == -101
    assert L.ga_host_sort_pair(0, 4, 3, _p(x), _p(x)) == -101            # :BE is not a sort order
    assert L.ga_host_chase_shifts(0, 0, 2, _p(x), _p(x), 2, _p(x), 2) == -101


def test_tridiag_fuzz_bitwise_vs_oracle(pkg, O):
    """400 random tridiagonals (sizes 1..59, entries scaled by 1e-200..1e200, exact zeros that split the matrix, tiny
    couplings, constant diagonals): the product's dstqrb! and the oracle's -- two independent restatements of
    src/dstqrb.jl -- give identical bits for the eigenvalues and, up to sign, the last-row components."""
    rng = np.random.default_rng(123)
    for _ in range(400):
        n = int(rng.integers(1, 60))
        scale = 10.0 ** rng.integers(-200, 200) if rng.random() < 0.3 else 1.0
        d0 = rng.standard_normal(n) * scale
        e0 = rng.standard_normal(max(n - 1, 0)) * scale
        if n > 3 and rng.random() < 0.5:
            e0[rng.integers(0, n - 1, size=2)] = 0.0
        if rng.random() < 0.2:
            e0 *= 1e-9
        if rng.random() < 0.1:
            d0[:] = d0[0]
        d, z = host_tridiag(pkg, d0, e0)
        do, zo = O.dstqrb(d0, e0)
        assert np.array_equal(d, do) and np.array_equal(np.abs(z), np.abs(zo))


def test_chase_shifts_fuzz_bitwise_vs_oracle(pkg, O):
    """300 random Lanczos tridiagonals (exact zeros and 1e-18 couplings that trigger the splitting / deflation logic of
    src/dsapps.jl:661-673,781-805, diagonal scales 1e-3..1e3, exact or arbitrary shifts): H and the first kev+1 columns of Q
    from the product's bulge chase equal the oracle's dsapps! bit for bit."""
    rng = np.random.default_rng(7)
    oracles = {}
    for _ in range(300):
        kev, np_ = int(rng.integers(1, 20)), int(rng.integers(1, 25))
        k = kev + np_
        if k not in oracles:
            oracles[k] = O.Oracle(sp.diags(np.arange(1.0, k + 3)).tocsr(), k)
        o = oracles[k]
        H = np.zeros((k, 2))
        H[:, 1] = rng.standard_normal(k) * rng.choice([1, 1e3, 1e-3])
        H[1:, 0] = np.abs(rng.standard_normal(k - 1))
        if rng.random() < 0.4:
            H[rng.integers(1, k, size=2), 0] = 0.0
        if rng.random() < 0.3:
            H[rng.integers(1, k, size=2), 0] = 1e-18
        T = np.diag(H[:, 1]) + np.diag(H[1:, 0], 1) + np.diag(H[1:, 0], -1)
        sh = (np.sort(np.linalg.eigvalsh(T))[:np_] if rng.random() < 0.5 else rng.standard_normal(np_)).copy()
        o.H[...] = H
        o.resid[:] = 1.0
        o.dsapps(kev, np_, sh)
        Hf = np.asfortranarray(H.copy())
        Q = np.zeros((k, k), order="F")
        assert pkg.lib().ga_host_chase_shifts(0, kev, np_, _p(sh), _p(Hf), k, _p(Q), k) == 0
        assert np.array_equal(Hf[:kev + 1], np.asarray(o.H)[:kev + 1])
        assert np.array_equal(Q[:, :kev + 1], np.asarray(o.Q)[:, :kev + 1])
        assert np.max(np.abs(Q.T @ Q - np.eye(k))) <= 1e-13

