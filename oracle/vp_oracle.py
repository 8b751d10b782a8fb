"""CPU restatement of the flow -> vanishing-point glue of the reference scripts (rows a8-a13 of SURVEY.md 8a).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Plain Python + numpy scalars, written from the
behavioural notes in SURVEY.md Appendix B.6, following -- including their quirks --

  checkInside + `st[~filter] = 0`          VP_detection_using_optical_flow.py:326-349, 552-553   (B.1)
  FlowLine / angle_between                 VP_detection_using_optical_flow.py:109-125, 276-283
  line acceptance + EMA                    VP_detection_using_optical_flow.py:565-574
  pair loop + cross_point                  VP_detection_using_optical_flow.py:296-324, 581-595   (B.5: float32)
  VP init / update / hide, CSV rows        VP_detection_using_optical_flow.py:157-194, 596-653   (B.2-B.4)
  corner refresh                           VP_detection_using_optical_flow.py:700-718
  save_csv                                 VP_detection_using_optical_flow.py:388-392
  vanishing-line pair (class VL)           VP_detection_using_optical_flow.py:196-250, 654       (drawing only)

`variant="LKF"` applies the LK_Final.py differences.  PINNING: oracle/make_golden.py runs
the UNMODIFIED scripts on synthetic clips and stores their CSVs (and the state of their `vl` object after every
update call, vl_*.npz) in tests/golden/; with cv2 as the `cvm` argument run_clip() reproduces those CSVs byte for
byte and the vanishing-line states to 1e-9 (tests/test_vp_oracle_golden.py).

No class or function of the reference is reused: state is a dict, lines and points are tuples.
"""
import csv
import io
import itertools
import math

import numpy as np

DEFAULTS = dict(WID=860, TP_NUM=20, VP_REF_NUM=10, VP_UPDATE_RATE=0.5, FL_UPDATE_RATE=0.05, TP_UPDATE_RATE=0.3,
                TP_UPDATE_TIME=10, MIN_ANG_DIF=25, MAX_CP_STD=1.0, MIN_FL_LEN=1.0, CP_THOLD=1 / 15, HIDE_VP_THOLD=50,
                START_X_FILTER=True, RESET_AVG_ON_HIDE=True, VP_REF=300)
LKF_OVERRIDES = dict(VP_REF_NUM=15, MIN_FL_LEN=1.5, START_X_FILTER=False, RESET_AVG_ON_HIDE=False)


def params(variant="VP", **kw):
    p = dict(DEFAULTS)
    if variant == "LKF":
        p.update(LKF_OVERRIDES)
    p.update(kw)
    return p


def feature_params(p):
    return dict(maxCorners=int(p["TP_NUM"] / 4), qualityLevel=0.3, minDistance=7, blockSize=7)


LK_PARAMS = dict(winSize=(15, 15), maxLevel=2, criteria=(2 | 1, 10, 0.03))


# ---------------------------------------------------------------------------------------------------
# ROI masks (VP.py:442-487) -- one-time setup, kept on cv2.fillPoly like the reference
# ---------------------------------------------------------------------------------------------------
def roi_polygons(WID, Hei):
    oL, oU, oR, oD = int(WID * 0.2), int(Hei * 0.65), int(WID * 0.8), int(Hei * 0.8)
    iL, iU, iR = int(WID * 0.47), int(Hei * 0.65), int(WID * 0.52)
    mid = (oD + iU) // 2
    P = [(WID // 2, mid), (oL, oD), (WID // 2, oD), (oR, oD), ((oR + iR) // 2, mid), (iR, iU), (WID // 2, iU),
         (iL, iU), ((oL + iL) // 2, mid)]
    full = [P[1], P[3], P[5], P[7]]
    small = [[P[0], P[8], P[1], P[2]], [P[0], P[2], P[3], P[4]], [P[0], P[4], P[5], P[6]], [P[0], P[6], P[7], P[8]]]
    return full, small


def roi_masks(WID, Hei):
    import cv2
    full, small = roi_polygons(WID, Hei)
    m = cv2.fillPoly(np.zeros((Hei, WID), np.uint8), [np.array(full)], 255)
    sm = [cv2.fillPoly(np.zeros((Hei, WID), np.uint8), [np.array(q)], 255) for q in small]
    return m, sm


# ---------------------------------------------------------------------------------------------------
# a8: ROI test with the dtype flip (B.1).  Mutates st like `st[~filter] = 0`.
# ---------------------------------------------------------------------------------------------------
def apply_roi_filter(p1, st, mask):
    H, W = mask.shape
    flags = []
    for i in range(len(p1)):
        fy, fx = math.floor(p1[i, 0, 1]), math.floor(p1[i, 0, 0])
        if st[i, 0] == 0 or fy > H or fx > W:
            flags.append(None)                       # the script appends the int 0 here
        else:
            flags.append(bool(mask[fy, fx] > 0))     # IndexError when fy == H, negative indices wrap
    if all(f is not None for f in flags):
        for i, f in enumerate(flags):                # boolean mask: the intended behaviour
            if not f:
                st[i] = 0
    else:                                            # int64 array: ~0 = -1, ~1 = -2 -> integer fancy indexing
        vals = [0 if f is None else int(f) for f in flags]
        if any(v == 0 for v in vals):
            st[-1] = 0
        if any(v == 1 for v in vals):
            st[-2] = 0                               # IndexError for a single point cannot occur (needs both kinds)
    return st


# ---------------------------------------------------------------------------------------------------
# a9: flow line geometry -> (start(f32 x2), stop(f32 x2), length, angle)
# ---------------------------------------------------------------------------------------------------
def flow_line(old_pt, new_pt):
    c, d = old_pt
    a, b = new_pt
    start = np.array([c, d])
    stop = np.array([a, b])
    vec = np.subtract(np.multiply(stop, [1, -1]), np.multiply(start, [1, -1]))      # float64, y flipped
    nrm = np.linalg.norm(vec)
    length = np.round(nrm, 2)
    u = vec / np.linalg.norm(vec)
    ang = np.arccos(np.clip(np.dot(u, np.array([1, 0]) / np.linalg.norm([1, 0])), -1.0, 1.0)) / np.pi * 180
    if vec[1] < 0:
        ang = 360 - ang
    return (start, stop, length, ang)


# ---------------------------------------------------------------------------------------------------
# a11: intersection in float32 scalar arithmetic (numpy >= 2 semantics, B.5); first argument is "line1"
# ---------------------------------------------------------------------------------------------------
def intersect(l1, l2):
    x1, y1, x2, y2 = l1
    x3, y3, x4, y4 = l2
    with np.errstate(all="ignore"):
        k1 = (y2 - y1) * 1.0 / (x2 - x1)
        b1 = y1 * 1.0 - x1 * k1 * 1.0
        if (x4 - x3) == 0:
            x = x3
        else:
            k2 = (y4 - y3) * 1.0 / (x4 - x3)
            b2 = y3 * 1.0 - x3 * k2 * 1.0
            if k1 - k2 == 0:
                return None
            x = (b2 - b1) * 1.0 / (k1 - k2)
        y = k1 * x * 1.0 + b1 * 1.0
    return x, y


def new_state(p):
    return dict(vp=None, recent=[], avg=[p["MIN_FL_LEN"], p["MIN_FL_LEN"]], vp_ult=0, tp_ult=0, rows=[], n_lines=0,
                n_cps=0, moved=False, vl=None, vl_rows=[])


# ---------------------------------------------------------------------------------------------------
# vanishing-line pair (VP.py:196-250): least-squares lines y(x) and x(y) through the last VP_REF history rows,
# re-anchored at the current VP.  scipy.stats.linregress (1.18) computes slope = (dx.dy / n) / (dx.dx / n) on the
# de-meaned samples and intercept = ymean - slope * xmean; restated with numpy only.
# ---------------------------------------------------------------------------------------------------
VL_FIELDS = ("m", "intercept", "mv", "interceptv", "bp_x", "bp_y", "lp_y", "rp_y", "up_x", "dp_x")


def linfit(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    n = len(x)
    xm, ym = np.mean(x), np.mean(y)
    dx, dy = x - xm, y - ym
    with np.errstate(all="ignore"):
        slope = (np.dot(dx, dy) / n) / (np.dot(dx, dx) / n)
    return slope, ym - slope * xm


def vl_update(S, p, Hei):
    """VL.update(all_vp[-VP_REF:], vp) (VP.py:224-245, called at VP.py:654): nothing until the VP has moved once."""
    if S["moved"]:
        WID = p["WID"]
        hist = S["rows"][-p["VP_REF"]:]
        x, y = [r[0] for r in hist], [r[1] for r in hist]
        m, b = linfit(x, y)
        mv, bv = linfit(y, x)
        bx, by = S["vp"][0], S["vp"][1]
        S["vl"] = dict(m=m, intercept=b, mv=mv, interceptv=bv, bp_x=bx, bp_y=by, lp_y=by - bx * m,
                       rp_y=by + (WID - 1 - bx) * m, up_x=bx - by * mv, dp_x=bx + (Hei - 1 - by) * mv)
    S["vl_rows"].append([float("nan")] * len(VL_FIELDS) if S["vl"] is None else [S["vl"][k] for k in VL_FIELDS])


def vl_endpoints(vl, WID, Hei, mode="best_point"):
    """VL.__call__ (VP.py:213-221, 245-250): (left, right, up, down, st); st is always True once initialised (the
    script's `is not nan` compares object identity)."""
    if vl is None:
        return None, None, None, None, None
    if mode == "best_point":
        return (0, vl["lp_y"]), (WID - 1, vl["rp_y"]), (vl["up_x"], 0), (vl["dp_x"], Hei - 1), True
    return ((0, vl["intercept"]), (WID - 1, vl["intercept"] + (WID - 1) * vl["m"]), (vl["interceptv"], 0),
            (vl["interceptv"] + (Hei - 1) * vl["mv"], Hei - 1), True)


# ---------------------------------------------------------------------------------------------------
# a9-a12 for one frame: tracks = [(good_old (K,2) f32, good_new (K,2) f32) per group]; appends CSV rows
# ---------------------------------------------------------------------------------------------------
def vp_frame(S, tracks, p, Hei):
    WID = p["WID"]
    lines = []
    for n, (good_old, good_new) in enumerate(tracks):
        for new, old in zip(good_new, good_old):
            a, b = new.ravel()
            c, d = old.ravel()
            if a == c and b == d:
                continue
            ln = flow_line((c, d), (a, b))
            if ln[3] > 180 and ln[2] > p["MIN_FL_LEN"]:
                S["avg"][n] = (S["avg"][n] + ln[2] * p["FL_UPDATE_RATE"]) / (1 + p["FL_UPDATE_RATE"])
                if ln[2] > S["avg"][n]:
                    lines.append(ln)
    S["n_lines"] += len(lines)
    for la, lb in itertools.combinations(lines, 2):
        adiff = abs(la[3] - lb[3])
        if adiff < p["MIN_ANG_DIF"] or adiff > 360 - p["MIN_ANG_DIF"]:
            continue
        if p["START_X_FILTER"] and abs(la[0][0] - lb[0][0]) < WID * 0.05:
            continue
        r = intersect(np.concatenate([lb[0], lb[1]]), np.concatenate([la[0], la[1]]))
        if r is None:
            continue
        x, y = r
        if y > la[0][1] or y > lb[0][1]:
            continue
        vp = S["vp"]
        if vp is not None:
            ok = (np.abs(np.array([vp[0] - x, vp[1] - y])) < np.array([WID * p["CP_THOLD"], Hei * p["CP_THOLD"]])).all()
            if not ok:
                continue
        cp = [x, y]                                   # mutable: the init quirk re-values the LAST cp in place (B.2)
        S["recent"].append(cp)
        S["n_cps"] += 1
        if vp is not None:
            dif = [np.array([q[0] - vp[0], q[1] - vp[1]]) for q in S["recent"][-p["VP_REF_NUM"]:]]
            mean = np.mean(dif, axis=0)
            std = np.std(dif, axis=0)
            acc = np.array([0., 0.])
            cnt = 0
            for dd in dif:
                if np.less_equal(dd, mean + std * p["MAX_CP_STD"]).all() and \
                        np.greater_equal(dd, mean - std * p["MAX_CP_STD"]).all():
                    acc = acc + dd
                    cnt += 1
            if cnt != 0:
                acc = acc / cnt
                vp[0] = vp[0] + acc[0] * p["VP_UPDATE_RATE"]
                vp[1] = vp[1] + acc[1] * p["VP_UPDATE_RATE"]
                S["moved"] = True                     # Point.update sets _moved (VP.py:180)
                S["rows"].append((vp[0], vp[1]))
                S["vp_ult"] = 0
        elif len(S["recent"]) >= p["VP_REF_NUM"]:
            acc = np.array([0., 0.])
            for q in S["recent"]:
                acc = acc + np.array([q[0], q[1]])
            acc = acc / p["VP_REF_NUM"]
            cp[0], cp[1] = acc[0], acc[1]             # the last CP object BECOMES the VP (aliasing kept)
            S["vp"] = cp
            S["moved"] = False                        # ... a cross-point object that has never been update()d
            S["vp_ult"] = 0
    if S["vp"] is not None:
        if S["vp_ult"] > p["HIDE_VP_THOLD"]:
            S["vp"] = None
            S["moved"] = False                        # a fresh Point(isVP=True); the VL object keeps its last lines
            S["recent"] = []
            if p["RESET_AVG_ON_HIDE"]:
                S["avg"] = [p["MIN_FL_LEN"], p["MIN_FL_LEN"]]
        else:
            S["rows"].append((S["vp"][0], S["vp"][1]))
            vl_update(S, p, Hei)
    return S


def detect_groups(cvm, img, small_masks, fp):
    groups = []
    for j in range(2):
        acc = []
        for i in range(2):
            r = cvm.goodFeaturesToTrack(img, mask=small_masks[j * 2 + i], **fp)
            if r is not None:
                acc.extend(r)
        groups.append(np.array(acc, dtype=np.float32))
    return groups


def run_clip(frames, cvm, variant="VP", resize=None, masks=None, on_frame=None, **kw):
    """frames: iterable of BGR uint8 frames (full resolution).  cvm: module/object offering cvtColor,
    GaussianBlur, goodFeaturesToTrack, calcOpticalFlowPyrLK with cv2 semantics and COLOR_BGR2GRAY.
    resize(frame, width) -> frame (default: oracle INTER_AREA).  Returns the state dict (rows = CSV)."""
    p = params(variant, **kw)
    if resize is None:
        from . import oracle as O
        resize = O.imutils_resize
    it = iter(frames)
    f0 = next(it)
    h, w = f0.shape[:2]
    WID = p["WID"]
    Hei = int(WID * (float(h) / float(w)))
    mask, small = masks if masks is not None else roi_masks(WID, Hei)
    fp = feature_params(p)

    def prep(f):
        return cvm.GaussianBlur(cvm.cvtColor(resize(f, WID), cvm.COLOR_BGR2GRAY), (3, 3), 0)

    old = prep(f0)
    p0s = detect_groups(cvm, old, small, fp)
    S = new_state(p)
    for frame in it:
        new = prep(frame)
        tracks = []
        if len(p0s[0]) != 0 or len(p0s[1]) != 0:
            for n in range(2):
                p0 = p0s[n]
                if len(p0) == 0:
                    tracks.append((np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32)))
                    continue
                p1, st, err = cvm.calcOpticalFlowPyrLK(old, new, p0, None, **LK_PARAMS)
                good_new, good_old = [], []
                if p1 is not None:
                    st = apply_roi_filter(p1, st, mask)
                    good_new.extend(p1[st == 1])
                    good_old.extend(p0[st == 1])
                good_new = np.array(good_new)
                good_old = np.array(good_old)
                p0s[n] = good_new.reshape(-1, 1, 2)
                tracks.append((good_old, good_new))
            vp_frame(S, tracks, p, Hei)
        else:
            vp_frame(S, [], p, Hei)
        if on_frame is not None:
            on_frame(S, tracks)
        old = new
        if (len(p0s[0]) + len(p0s[1])) < p["TP_NUM"] * p["TP_UPDATE_RATE"] or S["tp_ult"] == p["TP_UPDATE_TIME"]:
            S["tp_ult"] = 0
            fresh = detect_groups(cvm, old, small, fp)
            if len(fresh[0]) != 0 and len(fresh[1]) != 0:
                p0s = fresh
        S["tp_ult"] += 1
        S["vp_ult"] += 1
    return S


def csv_text(rows):
    """save_csv (VP.py:388-392): header x,y then one row per history entry, csv default dialect (\\r\\n)."""
    buf = io.StringIO(newline="")
    wr = csv.writer(buf)
    wr.writerow(["x", "y"])
    wr.writerows(rows)
    return buf.getvalue()
