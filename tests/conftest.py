import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "nn-sdp_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def _ensure_library():
    """The tests go through the C ABI: build the shared library if a fresh checkout has none yet
    (same command as __graft_entry__.build(); nvcc cross-compiles without a GPU)."""
    lib = os.path.join(ROOT, "nn-sdp_b200", "lib", "libnnsdp_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.run(["make", "-C", os.path.join(ROOT, "nn-sdp_b200"), "-j8"], check=True)


_ensure_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import nnsdp_b200 as nb

    if nb.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    c = nb.Context([0])
    from nnsdp_b200 import reference_api

    reference_api.set_default_context(c)
    return c
