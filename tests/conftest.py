import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (HERE, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference behind tests/harness (oracle/_ref/libvxh_ref.so, prebuilt by oracle/Makefile)."""
    import harness
    if not os.path.exists(harness.REF_LIB):
        pytest.skip("oracle/_ref/libvxh_ref.so not built (needs the reference checkout: make -C oracle ref)")
    return harness.reference()


@pytest.fixture(scope="session")
def restatement():
    import restate
    if not os.path.exists(restate.RESTATE_LIB):
        pytest.skip("build/oracle/libvxr_restate.so not built (make -C oracle restate)")
    return restate.Restate()


@pytest.fixture(scope="session")
def gpu_context():
    import voxels_b200
    ctx = voxels_b200.Context(0)  # raises without a CUDA device or without the built library: no fallback
    yield ctx
    ctx.close()
