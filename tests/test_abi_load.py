"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ssf/ssf.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_gpu
from ssf_gpu import capi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ssf", "ssf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in ssf.h but not exported by libssf_gpu.so"
    assert set(names) == set(capi.EXPORTS)


def test_struct_layouts_match_header():
    # sizes the C compiler gives the two POD structs of ssf.h (all members 4-byte)
    assert ctypes.sizeof(capi.IcpParams) == 32
    assert ctypes.sizeof(capi.IcpResult) == 64 + 9 * 4


def test_version_and_error_string():
    L = capi.lib()
    assert b"sm_100a" in L.ssf_version()
    assert L.ssf_ctx_create(0, None) == -1  # SSF_ERR_INVALID, no crash
    assert b"NULL" in L.ssf_last_error()


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    import ssf_gpu
    with pytest.raises(ssf_gpu.SsfError) as e:
        ssf_gpu.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "slam-sensor-fusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "ssf_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
