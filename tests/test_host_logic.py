"""Host-side logic that needs no GPU: argument validation of the cv2-signature boundary, the synthetic
clip generator, pyramid geometry."""
import numpy as np
import pytest


def test_boundary_argument_errors_before_any_device_work():
    import lk_optical_flow_method_b200.cv as cv
    img = np.zeros((48, 64), np.uint8)
    pts = np.zeros((4, 1, 2), np.float32)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))
    assert cv.calcOpticalFlowPyrLK(img, img, np.zeros((0, 1, 2), np.float32), None, **lk) == (None, None, None)
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(img, img, pts.astype(np.float64), None, **lk)
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(img, img, np.zeros((8, 1, 2), np.float32)[::2], None, **lk)
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(img, img[:-1], pts, None, **lk)
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(img, img, pts, None, winSize=(2, 9))
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(img.astype(np.float32), img, pts, None, **lk)
    with pytest.raises(cv.error):
        cv.goodFeaturesToTrack(img, 10, 0.0, 7)
    with pytest.raises(cv.error):
        cv.goodFeaturesToTrack(img, 10, 0.1, -1)
    with pytest.raises(cv.error):
        cv.goodFeaturesToTrack(img, 10, 0.1, 7, useHarrisDetector=True)
    with pytest.raises(cv.error):
        cv.goodFeaturesToTrack(img, 10, 0.1, 7, mask=np.zeros((3, 3), np.uint8))
    with pytest.raises(cv.error):
        cv.cvtColor(np.zeros((8, 8, 3), np.uint8), 4)
    with pytest.raises(cv.error):
        cv.GaussianBlur(img, (5, 5), 0)
    assert cv.COLOR_BGR2GRAY == 6 and cv.TERM_CRITERIA_EPS == 2 and cv.TERM_CRITERIA_COUNT == 1


def test_forwarding_of_non_hot_path_attributes():
    cv2 = pytest.importorskip("cv2")
    import lk_optical_flow_method_b200.cv as cv
    assert cv.fillPoly is cv2.fillPoly and cv.VideoCapture is cv2.VideoCapture


def test_synth_is_deterministic_and_device_independent():
    import torch
    from lk_optical_flow_method_b200 import synth
    tex = synth.make_texture(160, 96, 5)
    assert tex.shape == (192, 320, 3) and tex.dtype == np.uint8
    assert np.array_equal(tex, synth.make_texture(160, 96, 5))
    f = synth.zoom_frame(tex, 160, 96, 7)
    ft = synth.zoom_frame(torch.from_numpy(tex), 160, 96, 7)
    assert np.array_equal(f, ft.numpy())
    # the focus of expansion maps to the same texture point in every frame of a period
    W, H = 160, 96
    fx, fy = synth.foe_px(W, H)
    a = synth.zoom_frame(tex, W, H, 0)[int(fy), int(fx)].astype(int)
    b = synth.zoom_frame(tex, W, H, 20)[int(fy), int(fx)].astype(int)
    assert np.abs(a - b).max() <= 24


def test_pyramid_shapes():
    from lk_optical_flow_method_b200.device import pyramid_shapes
    assert pyramid_shapes(483, 860, 3) == [(483, 860), (242, 430), (121, 215)]
    assert pyramid_shapes(1080, 1920, 3) == [(1080, 1920), (540, 960), (270, 480)]


def test_vanishing_line_endpoints_match_the_oracle_restatement():
    """vp.vanishing_lines (host mirror of VL.__call__, VP.py:213-221, 245-250) against oracle.vl_endpoints in both modes,
    on vanishing-line states recorded from the unmodified script (tests/golden/vl_*.npz)."""
    import os
    from conftest import ROOT
    from lk_optical_flow_method_b200 import vp as V
    from oracle import vp_oracle as O
    assert V.VL_FIELDS == O.VL_FIELDS
    assert V.vanishing_lines(None, 860, 483) == (None, None, None, None, None)
    z = np.load(os.path.join(ROOT, "tests", "golden", "vl_vp_zoom300.npz"))
    rows = z["rows"][~np.isnan(z["rows"][:, 0])]
    assert len(rows) > 100
    for r in rows[::37]:
        vl = dict(zip(V.VL_FIELDS, (float(x) for x in r)))
        for mode in ("best_point", "other"):
            assert V.vanishing_lines(vl, 860, 483, mode) == O.vl_endpoints(vl, 860, 483, mode)
        lp, rp, up, dp, st = V.vanishing_lines(vl, 860, 483)
        # the recorded end points are the script's own (VP.py:241-244)
        assert st is True and lp == (0, vl["lp_y"]) and rp == (859, vl["rp_y"]) and up == (vl["up_x"], 0) and dp == (vl["dp_x"], 482)
        assert abs(vl["lp_y"] - (vl["bp_y"] - vl["bp_x"] * vl["m"])) < 1e-9


def test_csv_text_equals_csv_writer_on_edge_values():
    """csv_text formats the rows itself; it must give the bytes csv.writer gives for np.float64 fields (save_csv, VP.py:388-392)."""
    from lk_optical_flow_method_b200 import vp as V
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.uniform(-1e3, 1e3, 4000), rng.uniform(-1, 1, 500) * 10.0 ** rng.integers(-30, 30, 500),
                           np.array([0.0, -0.0, 1.0, 1e16, 1e15, 123456789012345678.0, 1e-4, 1e-5, 9.999e-5, np.inf, -np.inf, np.nan,
                                     412.7707922, 289.99299469, 5e-324, 1.7976931348623157e308, 0.1, 1 / 3, 2 ** 53, 1e22, 1e21])])
    if len(vals) % 2:
        vals = vals[:-1]
    rows = vals.reshape(-1, 2)
    assert V.csv_text(rows) == V.csv_text_reference(rows)
    assert V.csv_text(np.zeros((0, 2))) == V.csv_text_reference([]) == "x,y\r\n"
    assert V.csv_text([(1.5, 2.0)]) == "x,y\r\n1.5,2.0\r\n"


def test_cv_resize_forwards_everything_but_the_area_downscale():
    """cv.resize sends every use except the INTER_AREA non-integer uint8 down-scale to cv2 (no CUDA needed for those)."""
    cv2 = pytest.importorskip("cv2")
    import lk_optical_flow_method_b200.cv as cv
    a = np.random.default_rng(0).integers(0, 255, (72, 128, 3), dtype=np.uint8)
    assert cv.INTER_AREA == cv2.INTER_AREA
    for args, kw in [(((64, 36),), dict(interpolation=cv2.INTER_AREA)),          # integer factor: cv2's fast path
                     (((100, 50),), {}),                                           # INTER_LINEAR
                     (((200, 100),), dict(interpolation=cv2.INTER_AREA)),        # enlargement
                     ((None,), dict(fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)),
                     (((100, 50),), dict(interpolation=cv2.INTER_CUBIC))]:
        assert np.array_equal(cv.resize(a, *args, **kw), cv2.resize(a, *args, **kw))
    f = a.astype(np.float32)
    assert np.array_equal(cv.resize(f, (100, 50), interpolation=cv2.INTER_AREA), cv2.resize(f, (100, 50), interpolation=cv2.INTER_AREA))
