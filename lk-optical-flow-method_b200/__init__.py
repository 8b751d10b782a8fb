"""B200-native optical-flow hot path: drop-in for the cv2 calls of LK-Optical-Flow-Method.

Importable as ``lk_optical_flow_method_b200`` (see the alias package next to this directory).

    import lk_optical_flow_method_b200.cv as cv       # cvtColor / GaussianBlur / goodFeaturesToTrack /
                                                       # calcOpticalFlowPyrLK with cv2 signatures
    from lk_optical_flow_method_b200 import ops        # batched device-resident operators
    from lk_optical_flow_method_b200 import pipeline   # frame-pair pipeline (CUDA graph), clip runner

The arithmetic is hand-written CUDA for sm_100a in csrc/, reached through the C ABI of
include/lkflow_b200.h.  There is no CPU fallback: importing works anywhere, calling needs a B200.
"""
import importlib as _importlib

__version__ = "0.1.0"
_SUBMODULES = ("_lib", "device", "ops", "cv", "pipeline", "vp", "shard", "synth")


def __getattr__(name):
    if name in _SUBMODULES:
        return _importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
