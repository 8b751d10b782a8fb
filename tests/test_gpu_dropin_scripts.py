"""The drop-in claim, executed: the UNMODIFIED reference scripts (staged by oracle/make_golden.py under the git-ignored
baseline/_ref/) run headlessly with their four hot-path cv2 calls swapped for the B200 path by
lk_optical_flow_method_b200.cv.patch(cv2) (INTEGRATION.md level 1), and the vps/*.csv they write is compared with the CSV
the same script wrote on stock cv2 (tests/golden/*.csv): same number of rows, every VP within 1 px (north_star).
Also the work cache at the boundary: launch counts per script frame and equality with the uncached path."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
STAGE = os.path.join(ROOT, "baseline", "_ref")


def _clip(meta):
    from lk_optical_flow_method_b200 import synth
    tex = synth.make_texture(meta["W"], meta["H"], meta["clip"].get("seed", 0))
    kw = {k: v for k, v in meta["clip"].items() if k != "seed"}
    return [synth.zoom_frame(tex, meta["W"], meta["H"], t, **kw) for t in range(meta["frames"])]


@pytest.mark.parametrize("name", ["vp_zoom60", "lkf_zoom60", "vp_fast40"])
def test_unmodified_script_under_patch(name, tmp_path):
    import cv2
    from lk_optical_flow_method_b200 import cv as b200
    from lk_optical_flow_method_b200.device import Workspace
    from oracle import run_reference_script as R
    meta = json.load(open(os.path.join(GOLD, f"{name}.json")))
    script = os.path.join(STAGE, meta["script"])
    if not os.path.exists(script):
        pytest.skip("baseline/_ref/ is not staged (run oracle/make_golden.py stage_ref where /root/reference exists)")
    for t, f in enumerate(_clip(meta)):
        cv2.imwrite(str(tmp_path / f"f_{t:04d}.png"), f)
    lib = Workspace.current().lib
    b200.cache_enable(True)
    n0 = lib.lkf_ctx_launch_count(None)
    saved = b200.patch(cv2)
    try:
        text = R.run(script, "f_%04d.png", str(tmp_path), overrides=meta["overrides"] or None)
    finally:
        b200.unpatch(cv2, saved)
    assert lib.lkf_ctx_launch_count(None) > n0          # the CUDA path really ran inside the script
    hits = b200.cache_stats()["image_hits"]
    b200.cache_enable(False)
    assert hits > meta["frames"]
    assert text.startswith("x,y\r\n")
    got = np.loadtxt(text.splitlines()[1:], delimiter=",").reshape(-1, 2) if text.count("\n") > 1 else np.zeros((0, 2))
    gold = np.loadtxt(os.path.join(GOLD, f"{name}.csv"), delimiter=",", skiprows=1).reshape(-1, 2)
    assert got.shape == gold.shape, (got.shape, gold.shape)
    if len(gold):
        assert np.abs(got - gold).max() <= 1.0


def _script_frames(n):
    from lk_optical_flow_method_b200 import synth
    tex = synth.make_texture(860, 483, 7)
    return [synth.zoom_frame(tex, 860, 483, t) for t in range(n)]


def test_work_cache_launches_per_script_frame():
    """Four masked goodFeaturesToTrack calls -> one min-eigenvalue pass; two calcOpticalFlowPyrLK calls -> one pyramid
    pair; `processed_old_frame = processed_frame.copy()` carries next -> prev (VP.py:500-507, 545-546, 682)."""
    from lk_optical_flow_method_b200 import cv
    from lk_optical_flow_method_b200.device import Workspace
    lib = Workspace.current().lib
    z = np.load(os.path.join(GOLD, "roi_masks_860x483.npz"))
    masks = [np.ascontiguousarray(m) for m in z["small"]]
    fp = dict(maxCorners=5, qualityLevel=0.3, minDistance=7, blockSize=7)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))
    frames = _script_frames(3)

    def count(fn):
        a = lib.lkf_ctx_launch_count(None)
        r = fn()
        return lib.lkf_ctx_launch_count(None) - a, r

    def process(f):
        return cv.GaussianBlur(cv.cvtColor(f, cv.COLOR_BGR2GRAY), (3, 3), 0)

    def run(cached):
        cv.cache_enable(cached)
        out, launches = [], []
        old = process(frames[0])
        n_g = []
        p0 = []
        for m in masks:
            d, c = count(lambda: cv.goodFeaturesToTrack(old, mask=m, **fp))
            n_g.append(d)
            if c is not None:
                p0.extend(c)
        p0 = np.array(p0, dtype=np.float32)
        out.append(p0)
        groups = [np.ascontiguousarray(p0[: len(p0) // 2]), np.ascontiguousarray(p0[len(p0) // 2:])]
        for f in frames[1:]:
            new = process(f)
            n_l = []
            for g in groups:
                d, (p1, st, err) = count(lambda: cv.calcOpticalFlowPyrLK(old, new, g, None, **lk))
                n_l.append(d)
                out += [p1, st, err]
            launches.append(n_l)
            old = new.copy()                      # VP.py:682
        return out, n_g, launches

    ref, g0, l0 = run(False)
    got, g1, l1 = run(True)
    st = cv.cache_stats()
    cv.cache_enable(False)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    assert len(ref[0]) >= 8
    # uncached: every GFTT call pays the eig pass, every PyrLK call two pyramids (2 levels each)
    assert g0[0] == g0[1] == g0[2] == g0[3] and l0[0][0] == l0[0][1] == l0[1][0]
    # cached: the 2nd..4th GFTT call skip the eig pass
    assert g1[0] == g0[0] and g1[1] == g1[2] == g1[3] == g0[0] - 1
    # cached: first call of frame 1 builds both pyramids, the second none; frame 2 builds only the new frame's
    # (`old` is a .copy() of the previous `new`: found by content)
    lk_only = l1[0][1]
    assert lk_only >= 1 and l1[0][0] == l0[0][0] and l0[0][0] - lk_only == 4
    assert l1[1][0] - lk_only == 2 and l1[1][1] == lk_only
    assert st["image_hits"] > 0 and st["mask_misses"] == 4


def test_cache_never_returns_a_stale_image():
    """An array mutated in place is a different image: the content comparison must miss."""
    from lk_optical_flow_method_b200 import cv
    cv.cache_enable(True)
    f = _script_frames(1)[0]
    g = cv.GaussianBlur(cv.cvtColor(f, cv.COLOR_BGR2GRAY), (3, 3), 0)
    a = cv.goodFeaturesToTrack(g, 50, 0.05, 5, blockSize=3)
    g2 = g.copy()
    g2[100:200, 100:300] = 0
    b = cv.goodFeaturesToTrack(g2, 50, 0.05, 5, blockSize=3)
    g[100:200, 100:300] = 0                      # same object, new content
    c = cv.goodFeaturesToTrack(g, 50, 0.05, 5, blockSize=3)
    cv.cache_enable(False)
    d = cv.goodFeaturesToTrack(g, 50, 0.05, 5, blockSize=3)
    assert not np.array_equal(a, b) and np.array_equal(b, c) and np.array_equal(c, d)


def test_max_corners_zero_means_all_and_errors_are_cv2_errors():
    import cv2
    from lk_optical_flow_method_b200 import cv
    f = _script_frames(1)[0]
    g = cv2.GaussianBlur(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 0)
    want = cv2.goodFeaturesToTrack(g, 0, 0.001, 2, blockSize=3)
    got = cv.goodFeaturesToTrack(g, 0, 0.001, 2, blockSize=3)
    assert len(want) > 4096 and got.shape == want.shape and np.array_equal(got, want)
    with pytest.raises(cv2.error):
        cv.goodFeaturesToTrack(g, 10, -1.0, 2)
    with pytest.raises(cv2.error):
        cv.calcOpticalFlowPyrLK(g, g, np.zeros((4, 1, 2), np.float32)[::2], None)
