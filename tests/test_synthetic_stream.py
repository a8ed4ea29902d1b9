"""CPU checks of the streaming C5 generator (genericarpack.jl_b200/synthetic.py:tall_sparse_stream) used by the full-size GPU test and by
the sharded run: same matrix as the COO-based generator, and a row range is exactly that slice of the whole."""
import numpy as np


def test_stream_generator_matches_coo_generator(syn):
    a = syn.to_scipy(syn.tall_sparse(3000, 120, 10))
    b = syn.to_scipy(syn.tall_sparse_stream(3000, 120, 10, chunk_rows=700))
    b.sum_duplicates()
    assert abs(a - b).max() == 0.0
    rp, col, val, shape = syn.tall_sparse_stream(3000, 120, 10, chunk_rows=700)
    assert shape == (3000, 120) and rp[-1] == len(col) == len(val) == 30000
    assert np.all(np.diff(col.reshape(3000, 10), axis=1) >= 0)          # columns sorted inside every row


def test_stream_generator_row_range_is_a_slice(syn):
    rp, col, val, _ = syn.tall_sparse_stream(1000, 50, 10, chunk_rows=128)
    rp2, col2, val2, shape2 = syn.tall_sparse_stream(1000, 50, 10, chunk_rows=100, row_range=(300, 700))
    assert shape2 == (400, 50) and len(rp2) == 401 and rp2[0] == 0
    assert np.array_equal(col[3000:7000], col2) and np.array_equal(val[3000:7000], val2)
