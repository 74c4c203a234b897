"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/fplplus_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fplplus_b200 import lib as L
    return L.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "fplplus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_binding_table_matches_header(lib):
    from fplplus_b200 import lib as L
    assert sorted(L._SIGNATURES) == _declared()


def test_version_and_error_string(lib):
    assert lib.fpl_version() >= 100
    assert isinstance(lib.fpl_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    from fplplus_b200 import lib as L
    # channel counts that are not multiples of 8 are rejected before any launch
    with pytest.raises(L.FplError):
        L.call("fpl_dsbn_finalize", None, 10, None, None, None, None, None, 0.1, 1e-5, 1, None, None, None, None, 12, None)
    assert b"multiple of 8" in lib.fpl_last_error()
    assert lib.fpl_conv3d_weight_image_bytes(16, 16, 3) == 16 * 16 * 27 * 2
    assert lib.fpl_conv3d_weight_image_bytes(12, 16, 3) == -1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG/UBLKCP, tcgen05.ld -> LDTM (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from fplplus_b200 import build
    obj = os.path.join(os.path.dirname(build.LIB), "build", "conv_tc.o")
    build.build()
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM"):
        assert mnemonic in sass, mnemonic
