"""CPU-side checks of the C-ABI boundary: the library builds, loads without a GPU and
exports every symbol include/gamd_b200.h declares; no compute call is made."""
import os
import re

import pytest

from gamd_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gamd_b200 import build
    build.build()
    return _capi.load_library()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gamd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gamd_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_capi.SIGNATURES) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_version_and_enum_values(lib):
    assert b"sm_100a" in lib.gamd_version()
    src = open(os.path.join(ROOT, "include", "gamd_b200.h")).read()
    for name, val in (("GAMD_EINVAL", _capi.EINVAL), ("GAMD_ECUDA", _capi.ECUDA),
                      ("GAMD_EUNSUPPORTED", _capi.EUNSUPPORTED), ("GAMD_ECAPACITY", _capi.ECAPACITY),
                      ("GAMD_ESTATE", _capi.ESTATE), ("GAMD_ENOGPU", _capi.ENOGPU),
                      ("GAMD_NBR_LE", _capi.NBR_LE), ("GAMD_NBR_SELF", _capi.NBR_SELF),
                      ("GAMD_NBR_NOWRAP", _capi.NBR_NOWRAP), ("GAMD_MODEL_WATER", _capi.MODEL_WATER),
                      ("GAMD_PREC_BF16X3", _capi.PREC_BF16X3)):
        m = re.search(name + r"\s*=\s*(-?\d+)", src)
        assert m and int(m.group(1)) == val, name


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.GamdError) as ei:
        _capi.Context()
    assert ei.value.code == _capi.ENOGPU
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gamd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
