import os
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture
def oracle_engine():
    """Engine whose op lists are evaluated by the NumPy oracle on host buffers (no GPU)."""
    import _oracle_backend
    from ramba_b200 import ramba
    from ramba_b200.runtime import RT

    ramba.deferred_op.ramba_deferred_ops = None
    RT.reset()
    _oracle_backend.install()
    yield
    ramba.deferred_op.ramba_deferred_ops = None
    RT.reset()


@pytest.fixture
def gpu_engine():
    """Product configuration: op lists go through libramba_b200.so on cuda:0."""
    from ramba_b200 import ramba
    from ramba_b200.runtime import RT

    ramba.deferred_op.ramba_deferred_ops = None
    RT.reset()
    if os.environ.get("RB200_DRY_GPU_TESTS"):
        _dry_gpu()
    yield
    ramba.sync()
    RT.reset()


def _dry_gpu():
    """RB200_DRY_GPU_TESTS=1 (no GPU needed, proves nothing about the kernels): run the PYTHON of the -m gpu tests with the
    oracle backend standing in for the CUDA library, to check the tests' own logic against the engine as it is now - the
    round-1 failure was a bug in a GPU test that had never been executed."""
    import _oracle_backend
    from ramba_b200 import _cabi
    from ramba_b200 import runtime
    from ramba_b200.runtime import RT

    _oracle_backend.install()
    be = RT.backend
    if not hasattr(_cabi, "_dry_count"):
        _cabi._dry_count = [0]
        _cabi.launch_count = lambda: _cabi._dry_count[0]
        runtime.Runtime.is_cuda = property(lambda self: True)
    orig = be.run

    def run(fop, stream=None):
        _cabi._dry_count[0] += 1
        return orig(fop, stream)

    be.run = run
