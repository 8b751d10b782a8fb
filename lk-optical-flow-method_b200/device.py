"""Device-side plumbing: torch tensors as HBM buffers, image descriptors, per-device workspace.

PyTorch is used for memory, streams and (in bench.py) torch.distributed only; every arithmetic step
is a kernel of liblkflow_b200.so called through the C ABI (see _lib.py).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib
from ._lib import LkfImage, check


_cuda_ok = False


def require_cuda():
    """Raises without a CUDA device (no CPU fallback).  The positive answer is remembered: torch.cuda.is_available() costs
    ~4 us and sits on every per-call path of the drop-in functions."""
    global _cuda_ok
    if _cuda_ok:
        return
    if not torch.cuda.is_available():
        raise RuntimeError("lk_optical_flow_method_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    _cuda_ok = True


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw-handle query: 14 us less per call than
    building a torch.cuda.Stream object)."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_ELEM = {torch.uint8: 1, torch.int16: 2, torch.float32: 4}


def image_desc(t, channels=1):
    """LkfImage for a CUDA tensor shaped [B,H,W], [H,W] (channels==1 or packed last dim) or
    [B,H,W,C] / [H,W,C].  Rows must be dense (stride of x == C elements); row/image pitch are free."""
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor")
    es = _ELEM[t.dtype]
    nd = t.dim()
    if t.is_contiguous() and ((channels == 1 and nd == 3) or (channels > 1 and nd == 4 and t.shape[3] == channels)):
        # the common case on the per-call paths: dense planes, pitches follow from the shape (no stride queries)
        B, H, W = t.shape[0], t.shape[1], t.shape[2]
        rp = W * channels * es
        return LkfImage(t.data_ptr(), B, H, W, channels, rp, rp * H)
    if channels == 1:
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.dim() != 3 or t.stride(2) != 1:
            raise ValueError("expected [B,H,W] with unit x stride")
        B, H, W = t.shape
    else:
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if t.dim() != 4 or t.shape[3] != channels or t.stride(3) != 1 or t.stride(2) != channels:
            raise ValueError(f"expected [B,H,W,{channels}] interleaved")
        B, H, W, _ = t.shape
    return LkfImage(t.data_ptr(), B, H, W, channels, t.stride(1) * es, t.stride(0) * es)


def as_host_array(x):
    """numpy view of a host input (numpy array or CPU torch tensor)."""
    if isinstance(x, np.ndarray):
        return x
    if isinstance(x, torch.Tensor) and not x.is_cuda:
        return x.numpy()
    raise TypeError(f"unsupported host array type {type(x)!r}")


def is_device(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


class Workspace:
    """One per (device): the opaque lkf_ctx plus cached device buffers keyed by (tag, shape, dtype)."""

    _tls = threading.local()

    def __init__(self, device):
        require_cuda()
        self.device = torch.device("cuda", device)
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        check(self.lib.lkf_ctx_create(device, ctypes.byref(h)), "lkf_ctx_create")
        self.ctx = h
        self._bufs = {}

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.lkf_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def buf(self, tag, shape, dtype, pitch_align=None):
        key = (tag, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def launches(self):
        return int(self.lib.lkf_ctx_launch_count(None))

    @classmethod
    def private(cls, device=None):
        """A workspace of its own (own lkf_ctx, own scratch) for an object that captures CUDA graphs: a graph bakes the
        scratch addresses in, so they must not be shared with calls that could make the scratch grow (abi.cu: lkf_scratch)."""
        require_cuda()
        return cls(torch.cuda.current_device() if device is None else device)

    def use(self):
        """Context manager: inside it Workspace.current() (and so every ops.* call) is this workspace."""
        return _Use(self)

    @classmethod
    def current(cls):
        require_cuda()
        stack = getattr(cls._tls, "stack", None)
        if stack:
            return stack[-1]
        dev = torch.cuda.current_device()
        d = getattr(cls._tls, "ws", None)
        if d is None:
            d = cls._tls.ws = {}
        if dev not in d:
            d[dev] = cls(dev)
        return d[dev]


class _Use:
    def __init__(self, ws):
        self.ws = ws

    def __enter__(self):
        st = getattr(Workspace._tls, "stack", None)
        if st is None:
            st = Workspace._tls.stack = []
        st.append(self.ws)
        return self.ws

    def __exit__(self, *exc):
        Workspace._tls.stack.pop()
        return False


def pyramid_shapes(H, W, n_levels):
    out = [(H, W)]
    for _ in range(1, n_levels):
        H, W = (H + 1) // 2, (W + 1) // 2
        out.append((H, W))
    return out
