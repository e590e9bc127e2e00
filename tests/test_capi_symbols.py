"""CPU suite, part 3: the C-ABI library loads and exports every symbol that
include/ltp_b200.h declares (no compute calls -- there is no GPU here), and the product
has no CPU route: creating a planner without a device fails loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ltp_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from longtermplanner_b200 import _capi
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_capi.LIB)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ltp_b200.h but not exported"
    assert set(names) == set(_capi.EXPORTS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from longtermplanner_b200 import LongTermPlanner, _capi, workloads as W
    with pytest.raises(RuntimeError):
        LongTermPlanner(7, 0.001, *W.FRANKA7.arrays())
    h = ctypes.c_void_p()
    lim = [x.ctypes.data_as(ctypes.c_void_p) for x in W.FRANKA7.arrays()]
    rc = _capi.create(ctypes.byref(h), 0, 7, 0.001, *lim)
    assert rc == _capi.LTP_ERR_CUDA and not h.value


def test_product_sources_do_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "longtermplanner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text.replace("oracle/ltp_oracle.h (the reference", "") or f == "ltp_math.cuh", f
                assert "import oracle" not in text and "from oracle" not in text, f
