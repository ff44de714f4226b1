"""The C-ABI library loads and exports every symbol declared in include/wbk.h (no compute calls)."""

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wbk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wbk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for name in ("wbk_smooth", "wbk_contours", "wbk_index_run", "wbk_events_raster", "wbk_rasterize_rings",
                 "wbk_track_overlap", "wbk_create", "wbk_destroy", "wbk_last_error"):
        assert name in syms


def test_cuda_library_exports_every_declared_symbol():
    from wavebreaking_b200 import _build, _lib

    path = _build.build()  # nvcc cross-compiles without a GPU
    cdll = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(cdll, s)]
    assert not missing, missing
    # the Python binding declares a signature for every entry point it uses
    assert set(_lib._SIGNATURES) <= set(declared_symbols())
    cdll.wbk_version.restype = ctypes.c_int
    assert cdll.wbk_version() >= 100


def test_product_loader_refuses_to_run_without_cuda():
    import pytest
    import torch

    from wavebreaking_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    prev = _lib._LIB
    _lib._LIB = None
    try:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _lib.get()
    finally:
        _lib._LIB = prev
