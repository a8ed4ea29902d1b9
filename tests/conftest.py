import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def syn(pkg):
    return importlib.import_module(entry.PKG_NAME + ".synthetic")


@pytest.fixture(scope="session")
def O():
    return entry.load_oracle()


# Golden vectors literal in the reference's own tests, committed as a fixture (tests/golden/reference_kats.json,
# regenerated / audited from /root/reference/test/*.jl by tests/golden/make_golden.py).
import json  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as _f:
    GOLD = json.load(_f)

# the literal 15 x 15 matrix of test/interface_simple.jl:7 (CSC, 1-based)
KAT15_COLPTR = GOLD["kat15"]["colptr"]
KAT15_ROWVAL = GOLD["kat15"]["rowval"]
KAT15_NZVAL = GOLD["kat15"]["nzval"]
# golden start vector of test/dgetv0_simple.jl:6-15 (seed (1,3,5,7), n = 10)
GOLDEN_V0 = GOLD["dgetv0_start_vector"]["values"]
# dsdrv3 generalised problem, n = 100 (test/interface_high.jl:45,72,102)
DSDRV3_GOLDEN = GOLD["dsdrv3_generalized_n100"]["values"]


def kat15():
    import numpy as np
    import scipy.sparse as sp

    return sp.csc_matrix((KAT15_NZVAL, np.array(KAT15_ROWVAL) - 1, np.array(KAT15_COLPTR) - 1), shape=(15, 15)).tocsr()


def dsdrv3(n):
    """arpack_dsdrv3 (test/utility.jl:268-272): 1-D FEM Laplacian, stiffness A and mass matrix B."""
    import numpy as np
    import scipy.sparse as sp

    A = (sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) * (n + 1)).tocsr()
    B = (sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]) * (1.0 / (6 * (n + 1)))).tocsr()
    return A, B
