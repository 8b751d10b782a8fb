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


def process_pyramid(bgr, out_levels):
    """a1+a2+a3: out_levels[0] = process_img(bgr), out_levels[l] = pyrDown(out_levels[l-1]) (one ABI call)."""
    ws = Workspace.current()
    b4 = bgr if bgr.dim() == 4 else bgr.unsqueeze(0)
    n = len(out_levels)
    Arr = LkfImage * n
    la = Arr(*[image_desc(_b3(t)) for t in out_levels])
    s = image_desc(b4, 3)
    check(ws.lib.lkf_process_pyramid_u8(ctypes.byref(s), la, n, stream_ptr()), "lkf_process_pyramid_u8")
    return out_levels


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
    iters = torch.zeros((B, N), dtype=torch.int32, device=dev) if want_iters else None      # points beyond counts[b] stay 0
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


def pyrlk_dense(prev_pyr, next_pyr, params, want_iters=False, out=None):
    """a7 for prevPts = every pixel of level 0 in raster order (config C4): -> (next_pts [B,H*W,2], status [B,H*W] u8,
    err [B,H*W] f32[, iters]); same values as pyrlk() on that grid, no point array."""
    ws = Workspace.current()
    n_levels = len(prev_pyr)
    B, H, W = prev_pyr[0].shape
    dev = prev_pyr[0].device
    if out is not None:
        nxt, st, err = out
    else:
        nxt = torch.empty((B, H * W, 2), dtype=torch.float32, device=dev)
        st = torch.empty((B, H * W), dtype=torch.uint8, device=dev)
        err = torch.empty((B, H * W), dtype=torch.float32, device=dev)
    iters = torch.zeros((B, H * W), dtype=torch.int32, device=dev) if want_iters else None
    Arr = LkfImage * n_levels
    pa = Arr(*[image_desc(t) for t in prev_pyr])
    na = Arr(*[image_desc(t) for t in next_pyr])
    check(ws.lib.lkf_pyrlk_dense(ws.ctx, ctypes.byref(params), n_levels, pa, na, nxt.data_ptr(), st.data_ptr(), err.data_ptr(),
                                 None if iters is None else iters.data_ptr(), stream_ptr()), "lkf_pyrlk_dense")
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


def gftt_detect(gray, masks, max_corners, quality, min_distance, block_size, max_out, n_masks=1, enable=None,
                out=None):
    """a5+a6 fused for a batch (eig map in the workspace); `enable`: optional int32 [B] device predicate."""
    ws = Workspace.current()
    g = _b3(gray)
    B = g.shape[0]
    lists = B * n_masks
    if out is None:
        corners = torch.empty((lists, max_out, 2), dtype=torch.float32, device=g.device)
        counts = torch.zeros((lists,), dtype=torch.int32, device=g.device)
    else:
        corners, counts = out
    p = LkfGfttParams(int(max_corners), int(block_size), -1, 0, float(quality), float(min_distance))
    sd = image_desc(g)
    md = None if masks is None else image_desc(_b3(masks))
    check(ws.lib.lkf_gftt_detect(ws.ctx, ctypes.byref(p), ctypes.byref(sd), None if md is None else ctypes.byref(md),
                                 int(n_masks), None if enable is None else enable.data_ptr(), corners.data_ptr(),
                                 counts.data_ptr(), int(max_out), stream_ptr()), "lkf_gftt_detect")
    return corners, counts


def area_tables(ssize, dsize):
    """cv2's computeResizeAreaTab for one axis: (ofs int32 [dsize+1], src index int32, alpha float32)."""
    import math

    import numpy as np
    scale = ssize / dsize
    ofs, si, al = [0], [], []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cw = min(scale, ssize - f1)
        s1 = math.ceil(f1)
        s2 = min(math.floor(f2), ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            si.append(s1 - 1); al.append(np.float32((s1 - f1) / cw))
        for s in range(s1, s2):
            si.append(s); al.append(np.float32(1.0 / cw))
        if f2 - s2 > 1e-3:
            si.append(s2); al.append(np.float32(min(min(f2 - s2, 1.0), cw) / cw))
        ofs.append(len(si))
    return (np.asarray(ofs, np.int32), np.asarray(si, np.int32), np.asarray(al, np.float32))


_AREA_TABS = {}


def resize_area(img, dsize, out=None):
    """cv2.resize(img, (Wd, Hd), interpolation=INTER_AREA), non-integer down-scale; [B,H,W,3] or [B,H,W] uint8."""
    ws = Workspace.current()
    Wd, Hd = int(dsize[0]), int(dsize[1])
    ch = 3 if (img.dim() == 4 or (img.dim() == 3 and img.shape[-1] == 3 and img.stride(-1) == 1 and img.stride(-2) == 3)) else 1
    t = img if (img.dim() == 4 or ch == 1 and img.dim() == 3) else img.unsqueeze(0)
    if ch == 1 and t.dim() == 2:
        t = t.unsqueeze(0)
    B, H, W = t.shape[0], t.shape[1], t.shape[2]
    key = (H, W, Hd, Wd, t.device.index)
    if key not in _AREA_TABS:
        xt, yt = area_tables(W, Wd), area_tables(H, Hd)
        _AREA_TABS[key] = [torch.from_numpy(a).to(t.device) for a in (*xt, *yt)]
    xo, xs, xa, yo, ys, ya = _AREA_TABS[key]
    if out is None:
        out = torch.empty((B, Hd, Wd, 3) if ch == 3 else (B, Hd, Wd), dtype=torch.uint8, device=t.device)
    s, d = image_desc(t, ch), image_desc(out, ch)
    check(ws.lib.lkf_resize_area_u8(ctypes.byref(s), ctypes.byref(d), xo.data_ptr(), xs.data_ptr(), xa.data_ptr(),
                                    yo.data_ptr(), ys.data_ptr(), ya.data_ptr(), stream_ptr()), "lkf_resize_area_u8")
    return out
