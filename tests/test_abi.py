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


def test_workspace_sizes_are_host_arithmetic():
    """The `rovr_*_workspace` functions are pure host arithmetic (no device needed): the per-frame BatchNorm
    workspace is frames x blocks x 2C floats with 1..16 blocks per frame, never grows when the same pixels are
    spread over more frames than blocks are useful for, and rejects nonsense shapes with 0."""
    import _native
    lib = _native.lib
    for C in (64, 256, 2048):
        assert lib.rovr_bn_workspace(C) >= 2 * C * 4
        for frames, pix in ((1, 49), (6, 3136), (25, 12544), (500, 49), (4000, 196)):
            n = lib.rovr_bn_frames_workspace(C, frames, pix)
            blocks, rem = divmod(n, frames * 2 * C * 4)
            assert rem == 0 and 1 <= blocks <= 16, (C, frames, pix, n)
        # many frames: one or two blocks per frame are enough to fill the machine
        assert lib.rovr_bn_frames_workspace(C, 4000, 196) == 4000 * 2 * C * 4
    assert lib.rovr_bn_frames_workspace(0, 4, 49) == 0
    assert lib.rovr_bn_frames_workspace(64, 0, 49) == 0
    assert lib.rovr_bn_frames_workspace(64, 4, 0) == 0
