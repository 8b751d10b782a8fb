"""Batched, device-resident operators: thin typed wrappers over the C ABI (one kernel launch, or a fixed
short sequence, per call; nothing here synchronises).  Tensors are torch CUDA tensors:
images [B,H,W] uint8 (BGR: [B,H,W,3]), points [B,N,2] float32.
"""
import ctypes

import torch

from . import _lib
from ._lib import LkfGfttParams, LkfImage, LkfLkParams, check
from .device import Workspace, image_desc, pyramid_shapes, stream_ptr

TERM_CRITERIA_COUNT = 1
TERM_CRITERIA_MAX_ITER = 1
TERM_CRITERIA_EPS = 2
OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_LK_GET_MIN_EIGENVALS = 8


def _b3(t):
    return t if t.dim() == 3 else t.unsqueeze(0)


def bgr2gray(bgr, out=None):
    """a1: cv.cvtColor(img, COLOR_BGR2GRAY) (VP.py:406) on [B,H,W,3] -> [B,H,W]."""
    ws = Workspace.current()
    b4 = bgr if bgr.dim() == 4 else bgr.unsqueeze(0)
    if out is None:
        out = torch.empty(b4.shape[:3], dtype=torch.uint8, device=bgr.device)
    s, d = image_desc(b4, 3), image_desc(out)
    check(ws.lib.lkf_bgr2gray_u8(ctypes.byref(s), ctypes.byref(d), stream_ptr()), "lkf_bgr2gray_u8")
    return out


def gaussian_blur3(gray, out=None):
    """a2: cv.GaussianBlur(img, (3,3), 0) (VP.py:420) on [B,H,W]."""
    ws = Workspace.current()
    g = _b3(gray)
    if out is None:
        out = torch.empty_like(g, memory_format=torch.contiguous_format)
    s, d = image_desc(g), image_desc(out)
    check(ws.lib.lkf_gaussian_blur3_u8(ctypes.byref(s), ctypes.byref(d), stream_ptr()), "lkf_gaussian_blur3_u8")
    return out


def process_img(bgr, out=None):
    """a1+a2 fused: process_img() of VP.py:404-425 on [B,H,W,3] -> [B,H,W]."""
    ws = Workspace.current()
    b4 = bgr if bgr.dim() == 4 else bgr.unsqueeze(0)
    if out is None:
        out = torch.empty(b4.shape[:3], dtype=torch.uint8, device=bgr.device)
    s, d = image_desc(b4, 3), image_desc(out)
    check(ws.lib.lkf_process_img_u8(ctypes.byref(s), ctypes.byref(d), stream_ptr()), "lkf_process_img_u8")
    return out


def pyr_down(gray, out=None):
    """a3: one pyrDown level, [B,H,W] -> [B,(H+1)//2,(W+1)//2]."""
    ws = Workspace.current()
    g = _b3(gray)
    B, H, W = g.shape
    if out is None:
        out = torch.empty((B, (H + 1) // 2, (W + 1) // 2), dtype=torch.uint8, device=g.device)
    s, d = image_desc(g), image_desc(out)
    check(ws.lib.lkf_pyr_down_u8(ctypes.byref(s), ctypes.byref(d), stream_ptr()), "lkf_pyr_down_u8")
    return out


def effective_levels(H, W, win, max_level):
    """Number of pyramid levels cv2 uses (maxLevel+1 unless a level would be <= winSize)."""
    return _lib.load().lkf_effective_max_level(int(H), int(W), int(win[0]), int(win[1]), int(max_level)) + 1


def build_pyramid(gray, n_levels, out=None):
    """[level0, level1, ...] (level0 is `gray` itself)."""
    g = _b3(gray)
    levels = [g]
    for l in range(1, n_levels):
        levels.append(pyr_down(levels[-1], None if out is None else out[l]))
    return levels


def scharr(gray, out=None):
    """a4: int16 [B,H,W,2] = (Ix, Iy), unnormalised Scharr, REFLECT_101."""
    ws = Workspace.current()
    g = _b3(gray)
    if out is None:
        out = torch.empty(tuple(g.shape) + (2,), dtype=torch.int16, device=g.device)
    s, d = image_desc(g), image_desc(out, 2)
    check(ws.lib.lkf_scharr_s16(ctypes.byref(s), ctypes.byref(d), stream_ptr()), "lkf_scharr_s16")
    return out


def lk_params_struct(winSize=(21, 21), maxLevel=3, criteria=(TERM_CRITERIA_COUNT | TERM_CRITERIA_EPS, 30, 0.01),
                     flags=0, minEigThreshold=1e-4):
    return LkfLkParams(int(winSize[0]), int(winSize[1]), int(maxLevel), int(criteria[0]), int(criteria[1]),
                       int(flags), float(criteria[2]), float(minEigThreshold))


def pyrlk(prev_pyr, next_pyr, pts, params, counts=None, next_pts=None, want_iters=False,
          out=None):
    """a7 on pyramids: pts [B,N,2] f32 -> (next_pts [B,N,2], status [B,N] u8, err [B,N] f32[, iters])."""
    ws = Workspace.current()
    n_levels = len(prev_pyr)
    B, N, _ = pts.shape
    dev = pts.device
    if out is not None:
        nxt, st, err = out
        if next_pts is not None and next_pts.data_ptr() != nxt.data_ptr():
            nxt.copy_(next_pts)
    else:
        nxt = next_pts.clone() if next_pts is not None else torch.empty_like(pts)
        st = torch.empty((B, N), dtype=torch.uint8, device=dev)
        err = torch.empty((B, N), dtype=torch.float32, device=dev)
    iters = torch.empty((B, N), dtype=torch.int32, device=dev) if want_iters else None
    Arr = LkfImage * n_levels
    pa = Arr(*[image_desc(t) for t in prev_pyr])
    na = Arr(*[image_desc(t) for t in next_pyr])
    if not pts.is_contiguous():
        raise ValueError("prevPts must be contiguous")        # cv2 asserts the same
    check(ws.lib.lkf_pyrlk_sparse(ws.ctx, ctypes.byref(params), n_levels, pa, na, pts.data_ptr(),
                                  None if counts is None else counts.data_ptr(), N, nxt.data_ptr(),
                                  st.data_ptr(), err.data_ptr(),
                                  None if iters is None else iters.data_ptr(), stream_ptr()),
          "lkf_pyrlk_sparse")
    return (nxt, st, err, iters) if want_iters else (nxt, st, err)


def min_eigen(gray, block_size=3, dy_tail_cols=-1, out=None):
    """a5: cv.cornerMinEigenVal(img, blockSize, ksize=3) as goodFeaturesToTrack computes it."""
    ws = Workspace.current()
    g = _b3(gray)
    if out is None:
        out = torch.empty(g.shape, dtype=torch.float32, device=g.device)
    s, d = image_desc(g), image_desc(out)
    check(ws.lib.lkf_corner_min_eigen_f32(ctypes.byref(s), ctypes.byref(d), int(block_size), int(dy_tail_cols),
                                          stream_ptr()), "lkf_corner_min_eigen_f32")
    return out


def gftt_select(eig, masks, max_corners, quality, min_distance, max_out, n_masks=1, out=None):
    """a6: selection on an eig map.  masks: None or [n_masks,H,W] / [B*n_masks,H,W] uint8.
    -> corners [B*n_masks, max_out, 2] f32 (strongest first), counts [B*n_masks] int32
    (a negative count means a capacity overflow: -1-count)."""
    ws = Workspace.current()
    e = _b3(eig)
    B = e.shape[0]
    lists = B * n_masks
    if out is None:
        corners = torch.empty((lists, max_out, 2), dtype=torch.float32, device=e.device)
        counts = torch.empty((lists,), dtype=torch.int32, device=e.device)
    else:
        corners, counts = out
    p = LkfGfttParams(int(max_corners), 0, -1, 0, float(quality), float(min_distance))
    ed = image_desc(e)
    md = None if masks is None else image_desc(_b3(masks))
    check(ws.lib.lkf_gftt_select(ws.ctx, ctypes.byref(p), ctypes.byref(ed),
                                 None if md is None else ctypes.byref(md), int(n_masks), corners.data_ptr(),
                                 counts.data_ptr(), int(max_out), stream_ptr()), "lkf_gftt_select")
    return corners, counts
