"""Clip runner: the frame loop of the reference scripts (Run(), VP.py:427-737 / LKF.py:423-726) for a BATCH of
clips in lock-step, device resident, no host synchronisation per frame.

Per frame (all clips at once, a fixed launch sequence suitable for a CUDA graph):
    resize (INTER_AREA, VP.py:531) -> process_img (a1+a2) -> pyramid (a3) -> pyramidal LK for both point
    groups (a4+a7) -> vp_frame kernel (a8-a12: ROI filter, flow lines, cross points, VP state machine, CSV rows,
    refresh decision, and the vanishing-line pair of class VL, VP.py:196-250) -> goodFeaturesToTrack with the 4 ROI sub-masks, executed on the device only for the clips
    whose refresh flag is set (a5+a6) -> REPlace of the point groups (VP.py:713-715).
The previous frame's pyramid is kept (ping-pong), so every frame is converted and down-sampled once.

Host code here is configuration, buffer ownership and the CSV writer (a13); the ROI masks are a one-time
set-up kept on cv2.fillPoly exactly like the reference (VP.py:442-487, out of the hot path).
"""
import csv
import ctypes
import io
import os

import numpy as np
import torch

from . import ops
from ._lib import LkfVpParams, LkfVpState, check
from .device import Workspace, image_desc, pyramid_shapes, stream_ptr

# module constants of VP.py:21-57 ; LK_Final.py differences per SURVEY Appendix C
DEFAULTS = dict(WID=860, TP_NUM=20, VP_REF_NUM=10, VP_UPDATE_RATE=0.5, FL_UPDATE_RATE=0.05, TP_UPDATE_RATE=0.3,
                TP_UPDATE_TIME=10, MIN_ANG_DIF=25, MAX_CP_STD=1.0, MIN_FL_LEN=1.0, CP_THOLD=1 / 15, HIDE_VP_THOLD=50,
                START_X_FILTER=True, RESET_AVG_ON_HIDE=True, VP_REF=300)
LKF_OVERRIDES = dict(VP_REF_NUM=15, MIN_FL_LEN=1.5, START_X_FILTER=False, RESET_AVG_ON_HIDE=False)
LK_PARAMS = dict(winSize=(15, 15), maxLevel=2, criteria=(ops.TERM_CRITERIA_EPS | ops.TERM_CRITERIA_COUNT, 10, 0.03))


def config(variant="VP", **overrides):
    p = dict(DEFAULTS)
    if variant == "LKF":
        p.update(LKF_OVERRIDES)
    elif variant != "VP":
        raise ValueError("variant must be 'VP' (VP_detection_using_optical_flow.py) or 'LKF' (LK_Final.py)")
    p.update(overrides)
    return p


def feature_params(cfg):
    """feature_params of VP.py:92-95."""
    return dict(maxCorners=int(cfg["TP_NUM"] / 4), qualityLevel=0.3, minDistance=7, blockSize=7)


def roi_polygons(WID, Hei):
    """Vertices of the ROI trapezoid and its four sub-quadrilaterals (VP.py:442-471)."""
    oL, oR, oD = int(WID * 0.2), int(WID * 0.8), int(Hei * 0.8)
    iL, iU, iR = int(WID * 0.47), int(Hei * 0.65), int(WID * 0.52)
    mid = (oD + iU) // 2
    P = [(WID // 2, mid), (oL, oD), (WID // 2, oD), (oR, oD), ((oR + iR) // 2, mid), (iR, iU), (WID // 2, iU), (iL, iU),
         ((oL + iL) // 2, mid)]
    return [P[1], P[3], P[5], P[7]], [[P[0], P[8], P[1], P[2]], [P[0], P[2], P[3], P[4]], [P[0], P[4], P[5], P[6]],
                                      [P[0], P[6], P[7], P[8]]]


def roi_masks(WID, Hei):
    """(mask, [4 sub-masks]) uint8, filled with cv2.fillPoly like the reference (one-time set-up)."""
    import cv2
    full, small = roi_polygons(WID, Hei)
    m = cv2.fillPoly(np.zeros((Hei, WID), np.uint8), [np.array(full)], 255)
    return m, [cv2.fillPoly(np.zeros((Hei, WID), np.uint8), [np.array(q)], 255) for q in small]


def csv_text(rows):
    """save_csv (VP.py:388-392): header `x,y`, one row per VP history entry, csv default dialect (\\r\\n).
    csv.writer writes str(np.float64) for every field, which is the shortest round-trip repr -- the same text as repr() of
    the Python float (tests/test_host_logic.py compares the two on edge values) -- so the rows are formatted directly
    (10x faster than csv.writer on the 3 000 rows per second of a 64-clip batch)."""
    a = np.asarray(rows, dtype=np.float64).reshape(-1, 2)
    if a.size == 0:
        return "x,y\r\n"
    return "x,y\r\n" + "\r\n".join(f"{x!r},{y!r}" for x, y in a.tolist()) + "\r\n"


def csv_text_reference(rows):
    """The same text through csv.writer exactly as the scripts write it (VP.py:388-392); kept as the checker of csv_text."""
    buf = io.StringIO(newline="")
    wr = csv.writer(buf)
    wr.writerow(["x", "y"])
    wr.writerows([(np.float64(x), np.float64(y)) for x, y in rows])
    return buf.getvalue()


VL_FIELDS = ("m", "intercept", "mv", "interceptv", "bp_x", "bp_y", "lp_y", "rp_y", "up_x", "dp_x")


def vl_of_state(s):
    """The vanishing-line pair held in an LkfVpState (class VL, VP.py:196-250) as a dict keyed by VL_FIELDS, or None
    while the clip's VL object is not initialised (the VP has not moved yet)."""
    if not s.vl_init:
        return None
    return dict(m=s.vl_m, intercept=s.vl_b, mv=s.vl_mv, interceptv=s.vl_bv, bp_x=s.vl_bp[0], bp_y=s.vl_bp[1],
                lp_y=s.vl_lp_y, rp_y=s.vl_rp_y, up_x=s.vl_up_x, dp_x=s.vl_dp_x)


def vanishing_lines(vl, WID, Hei, mode="best_point"):
    """VL.__call__ (VP.py:213-221, 245-250): (left, right, up, down, st) end points of the two lines; all None
    before the first update.  mode 'best_point': the fitted slopes through the current VP; otherwise the fitted
    lines themselves.  st is True once initialised (the script's `is not nan` test compares identity)."""
    if vl is None:
        return None, None, None, None, None
    if mode == "best_point":
        return (0, vl["lp_y"]), (WID - 1, vl["rp_y"]), (vl["up_x"], 0), (vl["dp_x"], Hei - 1), True
    return ((0, vl["intercept"]), (WID - 1, vl["intercept"] + (WID - 1) * vl["m"]), (vl["interceptv"], 0),
            (vl["interceptv"] + (Hei - 1) * vl["mv"], Hei - 1), True)


def save_csv(rows, video_name, out_dir="./vps"):
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"vps_{video_name}.csv")
    with open(path, "w", newline="") as f:
        f.write(csv_text(rows))
    return path


class ClipBatchRunner:
    """B clips of equal frame size processed in lock-step on the current CUDA device."""

    def __init__(self, n_clips, frame_hw, variant="VP", masks=None, row_capacity=None, expected_frames=4096, **overrides):
        self.cfg = cfg = config(variant, **overrides)
        self.B = int(n_clips)
        h, w = int(frame_hw[0]), int(frame_hw[1])
        self.in_hw = (h, w)
        self.WID = int(cfg["WID"])
        self.Hei = int(self.WID * (float(h) / float(w)))                      # VP.py:430-432
        self.resize_needed = (self.Hei, self.WID) != (h, w)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.ws = Workspace.private()      # own ctx: step_graph's graphs hold this workspace's scratch addresses
        B, H, W = self.B, self.Hei, self.WID
        self.fp = feature_params(cfg)
        self.maxc = max(1, self.fp["maxCorners"])
        self.maxp = 2 * self.maxc                                             # two masks feed one point group
        mask, small = masks if masks is not None else roi_masks(W, H)
        self.mask = torch.from_numpy(np.ascontiguousarray(mask)).to(self.dev)
        self.small = torch.from_numpy(np.ascontiguousarray(np.stack(small))).to(self.dev)
        self.n_levels = ops.effective_levels(H, W, LK_PARAMS["winSize"], LK_PARAMS["maxLevel"])
        self.lk = ops.lk_params_struct(**LK_PARAMS)
        shapes = pyramid_shapes(H, W, self.n_levels)
        self.pyr = [[torch.zeros((B, hh, ww), dtype=torch.uint8, device=self.dev) for (hh, ww) in shapes] for _ in range(2)]
        self.cur = 0                                                          # index of the "old" pyramid
        self.resized = torch.zeros((B, H, W, 3), dtype=torch.uint8, device=self.dev) if self.resize_needed else None
        self.p0 = torch.zeros((B, 2 * self.maxp, 2), dtype=torch.float32, device=self.dev)    # lists 2b+g
        self.p1 = torch.zeros_like(self.p0)
        self.st = torch.zeros((B, 2 * self.maxp), dtype=torch.uint8, device=self.dev)
        self.err = torch.zeros((B, 2 * self.maxp), dtype=torch.float32, device=self.dev)
        self.counts = torch.zeros((B * 2,), dtype=torch.int32, device=self.dev)
        self.corners = torch.zeros((B * 4, self.maxc, 2), dtype=torch.float32, device=self.dev)
        self.ccounts = torch.zeros((B * 4,), dtype=torch.int32, device=self.dev)
        self.refresh = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        # rows per frame <= 1 + C(L,2) with L <= 2*maxp accepted lines; size for the expected clip length
        per_frame = 1 + (2 * self.maxp) * (2 * self.maxp - 1) // 2
        self.row_cap = int(row_capacity) if row_capacity else int(min(per_frame, 256) * expected_frames)
        self.rows = torch.zeros((B, self.row_cap, 2), dtype=torch.float64, device=self.dev)
        st0 = LkfVpState()
        st0.avg_len[0] = st0.avg_len[1] = float(cfg["MIN_FL_LEN"])
        raw = np.frombuffer(bytes(st0), dtype=np.uint8)
        self._state0 = torch.from_numpy(np.tile(raw, (B, 1)).copy()).to(self.dev)
        self.states = self._state0.clone()                                                # [B, sizeof(state)] bytes
        self.vp_params = LkfVpParams(
            cp_gate_x=self.WID * cfg["CP_THOLD"], cp_gate_y=self.Hei * cfg["CP_THOLD"], min_ang_dif=float(cfg["MIN_ANG_DIF"]),
            max_cp_std=float(cfg["MAX_CP_STD"]), vp_update_rate=float(cfg["VP_UPDATE_RATE"]),
            fl_update_rate=float(cfg["FL_UPDATE_RATE"]), min_fl_len=float(cfg["MIN_FL_LEN"]),
            tp_refresh_below=cfg["TP_NUM"] * cfg["TP_UPDATE_RATE"],
            start_x_sep=float(np.float32(self.WID * 0.05)) if cfg["START_X_FILTER"] else -1.0,
            vp_ref_num=int(cfg["VP_REF_NUM"]), hide_vp_thold=int(cfg["HIDE_VP_THOLD"]),
            reset_avg_on_hide=1 if cfg["RESET_AVG_ON_HIDE"] else 0, tp_update_time=int(cfg["TP_UPDATE_TIME"]),
            vl_ref=int(cfg["VP_REF"]))
        self.frames_done = 0
        self._graphs = [None, None]
        self._graph_in = None

    # -- building blocks --------------------------------------------------------------------------
    def _prepare(self, frames, slot):
        """resize + process_img + pyramid of a batch of BGR frames into pyramid `slot`."""
        src = frames
        if self.resize_needed:
            ops.resize_area(frames, (self.WID, self.Hei), out=self.resized)
            src = self.resized
        ops.process_img(src, out=self.pyr[slot][0])
        for l in range(1, self.n_levels):
            ops.pyr_down(self.pyr[slot][l - 1], out=self.pyr[slot][l])

    def _detect(self, slot, enable):
        ops.gftt_detect(self.pyr[slot][0], self.small, self.fp["maxCorners"], self.fp["qualityLevel"], self.fp["minDistance"],
                        self.fp["blockSize"], self.maxc, n_masks=4, enable=enable, out=(self.corners, self.ccounts))

    def _apply_refresh(self, force):
        check(self.ws.lib.lkf_vp_apply_refresh(self.states.data_ptr(), self.B, self.corners.data_ptr(), self.ccounts.data_ptr(),
                                               self.maxc, self.p0.data_ptr(), self.counts.data_ptr(), self.maxp,
                                               1 if force else 0, stream_ptr()), "lkf_vp_apply_refresh")

    def reset(self):
        """Back to the state before start(): a new batch of clips of the same geometry on the same buffers (the CUDA
        graphs of step_graph stay valid: they hold buffer addresses, not clip state)."""
        self.states.copy_(self._state0)
        for t in (self.p0, self.p1, self.st, self.err, self.counts, self.refresh, self.ccounts):
            t.zero_()
        self.cur = 0
        self.frames_done = 0

    # -- the frame loop -----------------------------------------------------------------------------
    def start(self, first_frames):
        """Step 1-2 of Run(): first frame -> corners of the four ROI sub-masks (VP.py:494-507)."""
        with self.ws.use():
            self._prepare(first_frames, self.cur)
            self._detect(self.cur, None)
            self._apply_refresh(force=True)

    def _step_launches(self, frames):
        with self.ws.use():
            self._step_launches_ws(frames)

    def _step_launches_ws(self, frames):
        old, new = self.cur, self.cur ^ 1
        self._prepare(frames, new)
        ops.pyrlk(self.pyr[old], self.pyr[new], self.p0, self.lk, out=(self.p1, self.st, self.err))
        md = image_desc(self.mask)
        check(self.ws.lib.lkf_vp_frame(ctypes.byref(self.vp_params), self.states.data_ptr(), self.B, self.p0.data_ptr(),
                                       self.p1.data_ptr(), self.st.data_ptr(), self.counts.data_ptr(), self.maxp,
                                       ctypes.byref(md), self.rows.data_ptr(), self.row_cap, self.refresh.data_ptr(), None,
                                       stream_ptr()), "lkf_vp_frame")
        self._detect(new, self.refresh)
        self._apply_refresh(force=False)

    def step(self, frames):
        """One frame of every clip: frames [B,h,w,3] uint8 CUDA tensor."""
        self._step_launches(frames)
        self.cur ^= 1
        self.frames_done += 1

    def step_graph(self, frames_buf):
        """Same as step() but replays a CUDA graph captured on the persistent input buffer `frames_buf`
        (two graphs, one per ping-pong parity)."""
        if self._graph_in is None or self._graph_in.data_ptr() != frames_buf.data_ptr():
            self._graphs, self._graph_in = [None, None], frames_buf
        g = self._graphs[self.cur]
        if g is None:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # capture must not change state: snapshot and restore everything the launches mutate
            snap = [t.clone() for t in (self.p0, self.counts, self.states, self.rows, self.refresh)]
            with torch.cuda.graph(g):
                self._step_launches(frames_buf)
            torch.cuda.synchronize()
            for t, s in zip((self.p0, self.counts, self.states, self.rows, self.refresh), snap):
                t.copy_(s)
            self._graphs[self.cur] = g
        g.replay()
        self.cur ^= 1
        self.frames_done += 1

    # -- results ------------------------------------------------------------------------------------
    def state_structs(self):
        raw = self.states.cpu().numpy()
        return [LkfVpState.from_buffer_copy(raw[b].tobytes()) for b in range(self.B)]

    def rows_host(self):
        """List (per clip) of (n_rows, 2) float64 arrays = the vps CSV rows.  Synchronises."""
        sts = self.state_structs()
        rows = self.rows.cpu().numpy()
        out = []
        for b, s in enumerate(sts):
            if s.row_overflow & 2:
                raise RuntimeError(f"clip {b}: the corner selection of a refresh overflowed its candidate capacity")
            if s.row_overflow:
                raise RuntimeError(f"clip {b}: CSV row capacity {self.row_cap} exceeded; pass a larger row_capacity")
            out.append(rows[b, :s.n_rows].copy())
        return out

    def vl_host(self):
        """Per clip: the current vanishing-line pair (dict keyed by VL_FIELDS) or None.  Synchronises."""
        return [vl_of_state(s) for s in self.state_structs()]

    def tracks_host(self):
        """Current point groups per clip: list of (group0 (K,1,2), group1 (K,1,2)) float32 arrays."""
        p0 = self.p0.cpu().numpy().reshape(self.B, 2, self.maxp, 2)
        c = self.counts.cpu().numpy().reshape(self.B, 2)
        return [tuple(p0[b, g, :c[b, g]].reshape(-1, 1, 2).copy() for g in range(2)) for b in range(self.B)]


def run_clips(frame_batches, variant="VP", use_graph=False, **kw):
    """frame_batches: iterable of [B,h,w,3] uint8 CUDA tensors (frame t of every clip).  Returns per-clip CSV rows."""
    it = iter(frame_batches)
    first = next(it)
    r = ClipBatchRunner(first.shape[0], first.shape[1:3], variant=variant, **kw)
    r.start(first)
    buf = torch.empty_like(first) if use_graph else None
    for fr in it:
        if use_graph:
            buf.copy_(fr)
            r.step_graph(buf)
        else:
            r.step(fr)
    return r.rows_host(), r
