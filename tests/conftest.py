import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "lazybandedmatrices.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def kats():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json"), encoding="utf-8") as fh:
        return json.load(fh)["kats"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def L():
    """The product package; GPU tests fail loudly (not skip) if the CUDA extension is missing."""
    import lbm_b200
    if not os.path.exists(lbm_b200.SO_PATH):
        # not built yet in this checkout: build in-tree (nvcc cross-compiles without a GPU).  This is
        # building the product, not falling back: if the build fails the tests fail.
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "lazybandedmatrices.jl_b200"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    lbm_b200.lib()
    return lbm_b200
