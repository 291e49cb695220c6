import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _product_library():
    """The C-ABI library is a build artefact (git-ignored): build it once if a fresh checkout runs the tests before
    `__graft_entry__.build()`.  nvcc cross-compiles without a GPU; the oracle's C library builds itself on demand."""
    from ir_ads_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()


GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "msda_golden.npz")
GOLDEN_CASES = ["ref_test", "d30", "d32", "d64", "d71", "d1025", "edge", "edge_d32", "enc_mini",
                "dec_mini", "stress_mini"]


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the reference's multi_scale_deformable_attn_pytorch (oracle/make_golden.py)."""
    blob = np.load(GOLDEN_PATH)

    def case(name):
        keys = [k for k in blob.files if k.startswith(name + "/")]
        return {k[len(name) + 1:]: blob[k] for k in keys}

    return case
