"""CPU: the C-ABI library builds, loads and exports every symbol include/cvb200.h declares."""
import os
import re

from clairvoyante_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cvb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libcvb200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "no ctypes signature for %s" % s
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_string(lib):
    assert lib.cvb_version() >= 100
    assert isinstance(lib.cvb_last_error(), bytes)


def test_no_cpu_fallback_without_gpu(lib):
    """On a box without a GPU the model constructor must raise, not fall back."""
    import ctypes
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clairvoyante_b200 import clairvoyante_v3 as cv
    with pytest.raises(RuntimeError):
        cv.Clairvoyante()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "clairvoyante_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "/root/reference" not in txt, f


def test_mode_enums_match_the_python_tables():
    """CVB_COMPUTE_* / CVB_TRAIN_* in the header are the values model.py passes through ctypes"""
    from clairvoyante_b200 import model
    src = open(os.path.join(ROOT, "include", "cvb200.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(CVB_[A-Z0-9_]+)\s*=\s*(\d+)", src)}
    assert {"CVB_COMPUTE_" + k.upper(): v for k, v in model.COMPUTE_MODES.items()} == \
        {k: v for k, v in enums.items() if k.startswith("CVB_COMPUTE_")}
    assert {"CVB_TRAIN_" + k.upper(): v for k, v in model.TRAIN_MODES.items()} == \
        {k: v for k, v in enums.items() if k.startswith("CVB_TRAIN_")}
    assert enums["CVB_V3"] == model._VARIANT_ID["v3"] and enums["CVB_V3_SLIM"] == model._VARIANT_ID["v3_slim"]
