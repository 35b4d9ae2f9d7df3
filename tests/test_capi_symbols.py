"""The C-ABI library loads without a GPU, exports every symbol the headers under include/ declare, and fails loudly
(no CPU fallback) when asked to compute without a CUDA device."""
import os
import re
import ctypes
import pytest

from positionbaseddynamics_b200 import _capi, model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pbdm?_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = declared("pbd_b200.h") + declared("pbd_b200_model.h")
    assert len(names) > 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_capi.SYMBOLS) == declared("pbd_b200.h")
    assert sorted(model.MODEL_SYMBOLS) == declared("pbd_b200_model.h")


def test_type_tables():
    assert [_capi.num_bodies(t) for t in range(15)] == [2, 2, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 2, 2]
    assert [_capi.num_params(t) for t in range(15)] == [2, 2, 2, 17, 17, 10, 9, 2, 2, 12, 12, 13, 24, 12, 6]


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_capi.PbdError) as ei:
        _capi.Engine(0)
    assert "no CPU fallback" in str(ei.value)
    m = model.HostModel(); m.add_regular_triangle_model(3, 3)
    with pytest.raises(_capi.PbdError):
        m.step(1)
    m.close()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, load or link it."""
    pkg = os.path.join(ROOT, "positionbaseddynamics_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liboracle", r"libpbdref", r"pyoracle\s*\.", r"#include\s+\"[^\"]*oracle"):
                    assert not re.search(pat, text, flags=re.M), (os.path.join(dp, f), pat)
