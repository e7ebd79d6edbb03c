"""CPU: the C-ABI library builds, loads, and exports exactly the symbols include/rovr_b200.h
declares; the ctypes table covers all of them; compute calls fail loudly without a GPU."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rovr_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rovr_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import _native
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(rovr_[a-zA-Z0-9_]+)", out)))
    declared = _declared()
    assert declared, "no declarations parsed"
    assert [s for s in declared if s not in exported] == [], "declared but not exported"
    assert [s for s in exported if s not in declared] == [], "exported but not declared"


def test_ctypes_table_covers_header():
    import _native
    assert sorted(_native.SIGNATURES.keys()) == _declared()
    assert _native.lib.rovr_abi_version() == 1


def test_library_is_sm100a_tcgen05():
    """The shipped SASS must contain the Blackwell tensor-core / TMA / TMEM instructions."""
    import _native
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass  # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import _native
    assert _native.lib.rovr_device_check() != 0
    assert "no CPU fallback" in _native.last_error() or "sm_" in _native.last_error()
    from local_net import LocalNetworkUNetNorm
    net = LocalNetworkUNetNorm()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 3, 8, 8))


def test_host_selftest_index_arithmetic():
    """The multiply-shift division used for tile indices (host-computed magic numbers, device-side
    umulhi + add + shift) agrees with integer division — checked on the host, no GPU involved."""
    import _native
    assert _native.lib.rovr_host_selftest() == 0
