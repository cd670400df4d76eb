import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

REFERENCE_MODEL_DIR = "/root/reference/model"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_MODEL_DIR, "segment_anything"))


@pytest.fixture(scope="session")
def ref_sa():
    """The reference's own segment_anything package (only in the build container; never on the GPU box)."""
    if not have_reference():
        pytest.skip("reference tree not present")
    if REFERENCE_MODEL_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_MODEL_DIR)
    import segment_anything  # noqa

    return segment_anything
