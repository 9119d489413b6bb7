"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol
include/gm_b200.h declares, and refuses loudly to run without a B200."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import gm_b200
    L = gm_b200.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert L.gm_version() >= 100


def test_no_cpu_fallback():
    import torch
    import gm_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gm_b200.GmError):
        gm_b200.ctx()
    with pytest.raises(gm_b200.GmError):
        gm_b200.GanEngine()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "generative-models_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("the oracle", ""), os.path.join(dp, f)
