import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    """GPU-marked tests skip themselves on a box without CUDA (a plain `pytest tests` stays green on CPU)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


_CLIPS = {}


def synth_frames(W, H, n=2, seed=0, t0=3):
    """n consecutive BGR frames of the deterministic zoom clip (cached per process)."""
    from lk_optical_flow_method_b200 import synth
    key = (W, H, n, seed, t0)
    if key not in _CLIPS:
        tex = synth.make_texture(W, H, seed)
        _CLIPS[key] = [synth.zoom_frame(tex, W, H, t0 + i) for i in range(n)]
    return _CLIPS[key]


def processed_pair(W, H, seed=0, t0=3):
    """(prev, next) uint8 gray+blur frames computed by the ORACLE (test infrastructure)."""
    from oracle import oracle as O
    f = synth_frames(W, H, 2, seed, t0)
    return tuple(O.blur3(O.bgr2gray(x)) for x in f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O
