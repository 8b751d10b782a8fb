"""ctypes front end of the CPU oracle (oracle/lkflow_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product package never does.

Each function restates one cv2 call of the reference's hot path:
  bgr2gray   <- cv.cvtColor(img, COLOR_BGR2GRAY)        VP_detection_using_optical_flow.py:406
  blur3      <- cv.GaussianBlur(g, (3,3), 0)            VP_detection_using_optical_flow.py:420
  pyrdown / scharr / pyrlk <- cv.calcOpticalFlowPyrLK   VP_detection_using_optical_flow.py:545-546
  min_eigen / gftt_select / good_features_to_track <- cv.goodFeaturesToTrack  VP...py:503,707
  resize_area <- imutils.resize(frame, width=860)       VP_detection_using_optical_flow.py:433,531
Pinned against cv2 4.13.0 by tests/test_oracle_vs_cv2.py and the fixtures in tests/golden/.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblkflow_oracle.so")

TERM_CRITERIA_COUNT = 1
TERM_CRITERIA_EPS = 2
OPTFLOW_USE_INITIAL_FLOW = 4
OPTFLOW_LK_GET_MIN_EIGENVALS = 8


def build(force=False):
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "lkflow_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.lko_gftt_select.restype = ctypes.c_int
        _lib.lko_pyrlk.restype = ctypes.c_int
        _lib.lko_effective_max_level.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def bgr2gray(bgr):
    bgr = _u8(bgr)
    H, W, _ = bgr.shape
    out = np.empty((H, W), np.uint8)
    lib().lko_bgr2gray(_p(bgr), H, W, _p(out))
    return out


def blur3(g):
    g = _u8(g)
    H, W = g.shape
    out = np.empty((H, W), np.uint8)
    lib().lko_blur3(_p(g), H, W, _p(out))
    return out


def pyrdown(g):
    g = _u8(g)
    H, W = g.shape
    out = np.empty(((H + 1) // 2, (W + 1) // 2), np.uint8)
    lib().lko_pyrdown(_p(g), H, W, _p(out))
    return out


def scharr(g):
    """-> int16 (H, W, 2) interleaved (Ix, Iy), unnormalised."""
    g = _u8(g)
    H, W = g.shape
    out = np.empty((H, W, 2), np.int16)
    lib().lko_scharr(_p(g), H, W, _p(out))
    return out


def min_eigen(g, block_size=7, tail_cols=-1):
    g = _u8(g)
    H, W = g.shape
    out = np.empty((H, W), np.float32)
    lib().lko_min_eigen(_p(g), H, W, int(block_size), int(tail_cols), _p(out))
    return out


def gftt_select(eig, mask=None, max_corners=0, quality=0.01, min_distance=1.0):
    eig = np.ascontiguousarray(eig, dtype=np.float32)
    H, W = eig.shape
    cap = H * W if max_corners <= 0 else int(max_corners)
    out = np.empty((max(cap, 1), 2), np.float32)
    m = None if mask is None else _u8(mask)
    mx = ctypes.c_double(0)
    n = lib().lko_gftt_select(_p(eig), None if m is None else _p(m), H, W, int(max_corners),
                              ctypes.c_double(quality), ctypes.c_double(min_distance), _p(out), cap,
                              ctypes.byref(mx))
    n = min(n, cap)
    if n == 0:
        return None
    return out[:n].reshape(-1, 1, 2).copy()


def good_features_to_track(image, maxCorners, qualityLevel, minDistance, mask=None, blockSize=3,
                           tail_cols=-1):
    return gftt_select(min_eigen(image, blockSize, tail_cols), mask, maxCorners, qualityLevel,
                       minDistance)


def effective_max_level(H, W, win, max_level):
    return lib().lko_effective_max_level(H, W, int(win[0]), int(win[1]), int(max_level))


def pyrlk(prev, nxt, pts, winSize=(21, 21), maxLevel=3,
          criteria=(TERM_CRITERIA_COUNT | TERM_CRITERIA_EPS, 30, 0.01), flags=0,
          minEigThreshold=1e-4, nextPts=None, return_iters=False):
    """-> (nextPts, status, err) shaped like cv2 ((N,1,2) f32, (N,1) u8, (N,1) f32)."""
    prev = _u8(prev)
    nxt = _u8(nxt)
    H, W = prev.shape
    p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
    n = p.shape[0]
    if n == 0:
        return None, None, None
    out = np.zeros((n, 2), np.float32)
    if nextPts is not None:
        out[:] = np.asarray(nextPts, np.float32).reshape(-1, 2)
    st = np.zeros((n,), np.uint8)
    err = np.zeros((n,), np.float32)
    iters = np.zeros((n,), np.int32)
    r = lib().lko_pyrlk(_p(prev), _p(nxt), H, W, _p(p), n, int(winSize[0]), int(winSize[1]),
                        int(maxLevel), int(criteria[0]), int(criteria[1]),
                        ctypes.c_double(criteria[2]), int(flags), ctypes.c_double(minEigThreshold),
                        _p(out), _p(st), _p(err), _p(iters))
    if r < 0:
        raise ValueError("oracle pyrlk: bad arguments")
    res = (out.reshape(np.asarray(pts).shape), st.reshape(-1, 1), err.reshape(-1, 1))
    return res + (iters,) if return_iters else res


def resize_area(img, dsize):
    """cv2.resize(img, dsize=(Wd,Hd), interpolation=INTER_AREA) for uint8 HxWxC (non-integer scale)."""
    img = _u8(img)
    H, W = img.shape[:2]
    cn = 1 if img.ndim == 2 else img.shape[2]
    Wd, Hd = int(dsize[0]), int(dsize[1])
    out = np.empty((Hd, Wd) if img.ndim == 2 else (Hd, Wd, cn), np.uint8)
    lib().lko_resize_area(_p(img), H, W, cn, _p(out), Hd, Wd)
    return out


def imutils_resize(img, width):
    """imutils.resize(image, width=W'): dim=(W', int(h*W'/w)), INTER_AREA."""
    h, w = img.shape[:2]
    r = width / float(w)
    return resize_area(img, (width, int(h * r)))


class OracleCV:
    """The oracle behind the four cv2 signatures the scripts call (for oracle/vp_oracle.run_clip)."""
    COLOR_BGR2GRAY = 6
    TERM_CRITERIA_EPS = 2
    TERM_CRITERIA_COUNT = 1

    @staticmethod
    def cvtColor(img, code):
        assert code == 6
        return bgr2gray(img)

    @staticmethod
    def GaussianBlur(img, ksize, sigma):
        assert tuple(ksize) == (3, 3) and sigma == 0
        return blur3(img)

    @staticmethod
    def goodFeaturesToTrack(image, maxCorners, qualityLevel, minDistance, mask=None, blockSize=3):
        return good_features_to_track(image, maxCorners, qualityLevel, minDistance, mask=mask, blockSize=blockSize)

    @staticmethod
    def calcOpticalFlowPyrLK(prev, nxt, pts, nextPts, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01), flags=0,
                             minEigThreshold=1e-4):
        return pyrlk(prev, nxt, pts, winSize, maxLevel, criteria, flags, minEigThreshold, nextPts)
