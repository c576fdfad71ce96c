"""The C-ABI libraries load without a GPU and export every symbol declared in
include/*.h; GPU entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import supersampler_b200 as S
from supersampler_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    S.build()


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(spsph?_[a-z0-9_]+)\s*\(", txt)))


def test_device_abi_symbols():
    L = S.device_lib()
    names = declared("spsp.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert L.spsp_abi_version() == 1


def test_host_abi_symbols():
    L = S.host_lib()
    names = [n for n in declared("spsp_host.h") if n.startswith("spsph_")]
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(S.SpspError):
        S.DeviceContext(31, 11, 1 << 50)
    with pytest.raises(S.SpspError):
        S.sketch_buffers([b">x\nACGT\n"])
    with pytest.raises(S.SpspError):
        S.compare_buffers([b"51 11 0 1000.000000\n"])


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: the product must not reference it."""
    pkg = os.path.join(ROOT, "supersampler_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("test infrastructure", ""), os.path.join(dp, f)
