"""Portable synthetic inputs for BASELINE.json's configs.

All generators are pure numpy and seeded by a SplitMix64 stream, so the Julia reference, the
CPU oracle and the GPU path can be fed bit-identical matrices (dump_csr writes the flat binary
file bench/reference_cpu.jl loads).  Matrices are returned as (rowptr int64, col int32, val, shape)
tuples in full (both triangles) CSR with sorted columns.
"""
import numpy as np


def splitmix64(seed, n):
    """n uint64 values of the SplitMix64 stream started at `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed, n):
    return (splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def laplacian2d(m, dtype=np.float64, row_range=None):
    """C2: 5-point stencil on an m x m grid, diag 4, off-diagonals -1, natural ordering.
    n = m*m, nnz = 5n - 4m.  row_range=(r0, r1) builds only those rows (global column indices)."""
    n = m * m
    r0, r1 = (0, n) if row_range is None else row_range
    rows = np.arange(r0, r1, dtype=np.int64)
    gx = rows % m
    has = np.stack([rows >= m, gx > 0, np.ones_like(rows, dtype=bool), gx < m - 1, rows < n - m], axis=1)
    cols = np.stack([rows - m, rows - 1, rows, rows + 1, rows + m], axis=1)
    vals = np.array([-1.0, -1.0, 4.0, -1.0, -1.0], dtype=dtype)
    counts = has.sum(axis=1)
    rowptr = np.zeros(r1 - r0 + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    col = cols[has].astype(np.int32)
    val = np.broadcast_to(vals, has.shape)[has].copy()
    return rowptr, col, val, (r1 - r0, n)


def laplacian2d_eigs(m, k):
    """Largest k analytic eigenvalues of laplacian2d(m), ascending."""
    c = 2.0 - 2.0 * np.cos(np.arange(1, m + 1) * np.pi / (m + 1))
    top = np.sort(c)[-min(m, 4 * k + 8):]
    lam = np.sort((top[:, None] + top[None, :]).ravel())
    return lam[-k:]


def _sym_from_coo(n, i, j, v):
    import scipy.sparse as sp

    A = sp.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    A = (A + A.conj().T).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def sprand_sym(n, avg_nnz_per_row=5.0, seed=1234):
    """C1: statistically what Symmetric(sprand(n,n,p) + A') is: ~p*n^2 uniform(0,1) entries at
    uniformly random positions, symmetrised by A + A' (so ~2*avg non-zeros per row)."""
    nnz = int(round(avg_nnz_per_row * n))
    i = (splitmix64(seed, nnz) % np.uint64(n)).astype(np.int64)
    j = (splitmix64(seed + 1, nnz) % np.uint64(n)).astype(np.int64)
    v = uniform01(seed + 2, nnz)
    A = _sym_from_coo(n, i, j, v)
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.copy(), A.shape


def hermitian_stencil(m, seed=99):
    """C4: Hermitian 5-point stencil on an m x m grid: real diagonal U(1,2), unit-modulus
    off-diagonals e^{i theta} (conjugated in the lower triangle)."""
    import scipy.sparse as sp

    n = m * m
    rows = np.arange(n, dtype=np.int64)
    gx = rows % m
    d = 1.0 + uniform01(seed, n)
    th1 = 2 * np.pi * uniform01(seed + 1, n)
    thm = 2 * np.pi * uniform01(seed + 2, n)
    right = gx < m - 1
    down = rows < n - m
    U = sp.coo_matrix((np.concatenate([np.exp(1j * th1[right]), np.exp(1j * thm[down])]),
                       (np.concatenate([rows[right], rows[down]]), np.concatenate([rows[right] + 1, rows[down] + m]))),
                      shape=(n, n)).tocsr()
    A = (U + U.conj().T + sp.diags(d.astype(np.complex128))).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.copy(), A.shape


def tall_sparse(m, n, nnz_per_row=10, seed=7):
    """C5: m x n, nnz_per_row uniform random columns per row, values U(0,1)."""
    import scipy.sparse as sp

    nnz = m * nnz_per_row
    col = (splitmix64(seed, nnz) % np.uint64(n)).astype(np.int64)
    val = uniform01(seed + 1, nnz)
    row = np.repeat(np.arange(m, dtype=np.int64), nnz_per_row)
    A = sp.coo_matrix((val, (row, col)), shape=(m, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.copy(), A.shape


def tall_sparse_stream(m, n, nnz_per_row=10, seed=7, chunk_rows=1 << 22, row_range=None):
    """C5 at full size (50M x 2M, nnz = 5e8) without the COO detour of tall_sparse: rows are generated chunk by chunk
    straight into CSR (same SplitMix64 streams, so entry k of row i is the same (column, value) pair as tall_sparse's),
    columns sorted within each row, duplicate columns of a row kept as separate entries (a valid CSR operator; tall_sparse
    sums them, which changes values in ~2e-5 of the rows only).  Peak host memory ~ nnz * 12 bytes + one chunk."""
    ra, rb = (0, m) if row_range is None else row_range        # row_range: only these rows of the same matrix (sharded runs)
    nnz = (rb - ra) * nnz_per_row
    base = ra * nnz_per_row
    rowptr = np.arange(0, nnz + 1, nnz_per_row, dtype=np.int64)
    col = np.empty(nnz, dtype=np.int32)
    val = np.empty(nnz, dtype=np.float64)
    with np.errstate(over="ignore"):
        for r0 in range(ra, rb, chunk_rows):
            r1 = min(rb, r0 + chunk_rows)
            k0, k1 = r0 * nnz_per_row, r1 * nnz_per_row
            idx = np.arange(k0 + 1, k1 + 1, dtype=np.uint64)

            def mix(sd):
                z = np.uint64(sd) + idx * np.uint64(0x9E3779B97F4A7C15)
                z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
                z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
                return z ^ (z >> np.uint64(31))

            c = (mix(seed) % np.uint64(n)).astype(np.int32).reshape(r1 - r0, nnz_per_row)
            v = ((mix(seed + 1) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)).reshape(r1 - r0, nnz_per_row)
            order = np.argsort(c, axis=1, kind="stable")
            col[k0 - base:k1 - base] = np.take_along_axis(c, order, axis=1).ravel()
            val[k0 - base:k1 - base] = np.take_along_axis(v, order, axis=1).ravel()
    return rowptr, col, val, (rb - ra, n)


def to_scipy(csr):
    import scipy.sparse as sp

    rowptr, col, val, shape = csr
    return sp.csr_matrix((val, col, rowptr), shape=shape)


def dump_csr(path, csr):
    """Flat binary: int64 nrows, ncols, nnz, elem_bytes; int64 rowptr; int32 col; val."""
    rowptr, col, val, shape = csr
    with open(path, "wb") as f:
        np.array([shape[0], shape[1], len(col), val.dtype.itemsize], dtype=np.int64).tofile(f)
        rowptr.astype(np.int64).tofile(f)
        col.astype(np.int32).tofile(f)
        val.tofile(f)
