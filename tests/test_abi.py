"""The C-ABI library loads on a CPU-only box and exports every symbol include/lkflow_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "lkflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lkf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_typed():
    from lk_optical_flow_method_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    hdr = open(os.path.join(ROOT, "include", "lkflow_b200.h")).read()
    assert lib.lkf_abi_version() == int(re.search(r"#define LKF_ABI_VERSION (\d+)", hdr).group(1)) == _lib.ABI_VERSION
    assert lib.lkf_strerror(0) == b"ok" and lib.lkf_strerror(-1) == b"bad argument"


def test_struct_layouts_match_header():
    from lk_optical_flow_method_b200 import _lib
    assert ctypes.sizeof(_lib.LkfImage) == 8 + 4 * 4 + 8 + 8
    assert ctypes.sizeof(_lib.LkfLkParams) == 6 * 4 + 2 * 8
    assert ctypes.sizeof(_lib.LkfGfttParams) == 4 * 4 + 2 * 8
    assert _lib.LkfLkParams.epsilon.offset == 24 and _lib.LkfGfttParams.quality_level.offset == 16
    # the compiled structs (host arithmetic inside the library) agree with the ctypes mirrors
    lib = _lib.load()
    for which, cls in enumerate((_lib.LkfImage, _lib.LkfLkParams, _lib.LkfGfttParams, _lib.LkfVpParams, _lib.LkfVpState)):
        assert lib.lkf_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.lkf_struct_size(99) == -1


def test_pure_host_entry_points():
    """lkf_effective_max_level is host arithmetic: cv2 reduces maxLevel when a level would be <= winSize."""
    from lk_optical_flow_method_b200 import _lib
    from oracle import oracle as O
    lib = _lib.load()
    for (H, W, win, ml) in [(483, 860, 15, 2), (120, 160, 31, 4), (1080, 1920, 21, 2), (2160, 3840, 31, 3), (40, 40, 21, 3)]:
        assert lib.lkf_effective_max_level(H, W, win, win, ml) == O.effective_max_level(H, W, (win, win), ml)
    assert lib.lkf_effective_max_level(120, 160, 31, 31, 4) == 1          # SURVEY A.3 probe


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of computing on the host."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import lk_optical_flow_method_b200.cv as cv
    with pytest.raises(RuntimeError):
        cv.cvtColor(np.zeros((8, 8, 3), np.uint8), cv.COLOR_BGR2GRAY)
