"""CPU-only: the C-ABI library loads and exports every symbol include/dirb200.h
declares; argument validation works without a GPU (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "dirb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dirb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import _lib
    import _convlib, resnet  # noqa: F401  (register the conv-stack / runner bindings)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dirb200.h but not exported"
    # and the Python binding table covers the header
    assert set(syms) <= set(_lib.exported_symbols()), set(syms) - set(_lib.exported_symbols())


def test_version_and_error_channel():
    import _lib
    assert _lib.raw("dirb200_version")() >= 100
    # invalid argument is rejected before any CUDA call
    rc = _lib.raw("dirb200_fds_smooth_tables")(None, 10, 4, None, 5, None, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = _lib.raw("dirb200_loss_fwd_bwd")(99, None, None, None, 4, 0.0, 1.0, 0, 1.0, None, None, None, 0, None)
    assert rc == -1 and "kind" in _lib.last_error()


def test_no_cpu_fallback():
    import pytest
    import torch
    import _lib
    from loss import weighted_l1_loss
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.Dirb200Error):
        weighted_l1_loss(torch.zeros(4, 1), torch.zeros(4, 1))
