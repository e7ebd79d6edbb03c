"""pytest configuration: registers the `gpu` marker and puts the product package (bare-name
modules, like the reference's script directory) and the oracle on sys.path."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reinformcement-optimized-video-reconstruction_b200")
ORACLE = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, ORACLE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the PyTorch checker must be true fp32: no TF32 in its convolutions / matmuls
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure librovr_b200.so exists (nvcc cross-compiles on the CPU box)."""
    import build_native
    build_native.build()
    yield


@pytest.fixture(autouse=True)
def _no_kernel_hang(request):
    """After every GPU test: no kernel may have hit an mbarrier timeout."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import torch
    if not torch.cuda.is_available():
        return
    import _native
    torch.cuda.synchronize()
    code = _native.hang_code()
    assert code == 0, f"a kernel timed out on mbarrier code 0x{code:x}"
