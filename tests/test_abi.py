"""The C-ABI library loads and exports every symbol include/xrft_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "xrft_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(xrftb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound():
    from xrft_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/xrft_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, "xrft_b200/_lib.py EXPORTS out of sync with the header"
    bound = _lib.load()
    assert bound.xrftb_version() == 100
    assert isinstance(bound.xrftb_last_error(), bytes)
    assert bound.xrftb_launch_count(1) >= 0


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly; it never routes through the oracle."""
    import torch
    import numpy as np
    import xrft_b200 as xrft
    from xrft_b200._lib import XrftbError

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    da = xrft.DataArray(np.random.rand(8, 8), dims=["y", "x"])
    with pytest.raises(XrftbError):
        xrft.power_spectrum(da)
    for f in os.listdir(os.path.join(ROOT, "xrft_b200")):
        if f.endswith(".py"):
            code = open(os.path.join(ROOT, "xrft_b200", f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", code, flags=re.M), f"{f} imports the oracle"
