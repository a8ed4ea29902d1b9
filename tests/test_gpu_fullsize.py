"""BASELINE.json configs at their stated sizes that do not fit the oracle's time budget: checked through the
size-independent properties the reference's own tests use (check_svd, test/utility.jl:244-253)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c5_svds_full_size_one_gpu(pkg, syn):
    """C5: svds on the synthetic tall sparse 50M x 2M matrix (nnz = 5e8) via the normal equations, k = 20, ncv = 60, on ONE
    B200 (6 GB of CSR twice -- A and its explicit adjoint -- plus a 1 GB basis and the 8 GB of left vectors fit easily).
    Properties: Ritz estimates converged, sigma_i^2 are eigenvalues of A'A to the Lanczos tolerance (||A'(A v) - s^2 v||),
    ||A v_i - s_i u_i|| at rounding level, U and V orthonormal (src/idonow_ops.jl:279-333, src/interface.jl:304-337)."""
    import psutil
    import scipy.sparse as sp

    if os.environ.get("GA_TEST_SKIP_C5_FULL") == "1":
        pytest.skip("GA_TEST_SKIP_C5_FULL=1")
    if psutil.virtual_memory().available < 48 * 2 ** 30:
        pytest.skip("needs ~30 GB of host memory for the 5e8-non-zero CSR, its adjoint and the 50M x 20 left vectors")
    m, n, k, ncv = 50_000_000, 2_000_000, 20, 60
    t0 = time.time()
    csr = syn.tall_sparse_stream(m, n, 10)
    t_gen = time.time() - t0
    t0 = time.time()
    U, S, V = pkg.svds(pkg.ArpackNormalOp(csr), k, ncv=ncv)
    t_svds = time.time() - t0
    assert len(S) == k and np.all(np.diff(S) <= 0) and S[-1] > 0
    assert U.shape == (m, k) and V.shape == (n, k)
    assert np.max(np.abs(V.T @ V - np.eye(k))) <= 1e-12
    assert np.max(np.abs(U.T @ U - np.eye(k))) <= 1e-12
    A = sp.csr_matrix((csr[2], csr[1], csr[0].astype(np.int32)), shape=(m, n))
    for i in (0, k // 2, k - 1):
        av = A @ V[:, i]
        assert np.linalg.norm(av - U[:, i] * S[i]) <= 1e-11 * S[0]
        assert np.linalg.norm(A.T @ av - S[i] ** 2 * V[:, i]) <= 1e-9 * S[0] ** 2
    print("C5 full size: generation %.0f s, svds %.1f s (upload + adjoint build + solve + vectors), sigma_1 = %.6f, sigma_20 = %.6f"
          % (t_gen, t_svds, S[0], S[-1]))


@pytest.mark.parametrize("script,args", [("dist_gpu_check.py", ["300"]), ("dist_gpu_check_normal.py", []), ("dist_gpu_check_f32.py", []), ("dist_gpu_check_general.py", [])])
def test_two_gpu_parity_under_torchrun(script, args):
    """The row-sharded path (halo exchange + fused peer-memory slot reduction) against the single-GPU path and the oracle:
    tests/dist_gpu_check*.py under torchrun with two ranks.  Needs two visible GPUs (skipped on a one-GPU box; bench.py's
    parity_gate is the check that runs with every `bench.py --gpus N`)."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29533 + sum(map(ord, script)) % 40), os.path.join(here, script)] + args   # one port per script
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
