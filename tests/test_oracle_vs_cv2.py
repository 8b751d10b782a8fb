"""Pins the oracle directly against cv2 (the third-party module that holds the reference's arithmetic)
on edge-case inputs: odd / tiny sizes, random noise, masks, window clamping.  Skipped without cv2."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("W,H", [(67, 35), (64, 48), (5, 4), (131, 97), (320, 240)])
def test_integer_stages_bit_exact(oracle, W, H):
    rng = np.random.default_rng(W * H)
    f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    g = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(oracle.bgr2gray(f), g)
    b = cv2.GaussianBlur(g, (3, 3), 0)
    assert np.array_equal(oracle.blur3(g), b)
    assert np.array_equal(oracle.pyrdown(b), cv2.pyrDown(b))
    s = oracle.scharr(b)
    assert np.array_equal(s[..., 0], cv2.Scharr(b, cv2.CV_16S, 1, 0))
    assert np.array_equal(s[..., 1], cv2.Scharr(b, cv2.CV_16S, 0, 1))


@pytest.mark.parametrize("W,H,bs", [(320, 240, 7), (131, 97, 3), (96, 64, 5), (200, 120, 2)])
def test_corner_pipeline_identical(oracle, W, H, bs):
    from lk_optical_flow_method_b200 import synth
    img = cv2.GaussianBlur(cv2.cvtColor(synth.zoom_frame(synth.make_texture(W, H, bs), W, H, 2), cv2.COLOR_BGR2GRAY), (3, 3), 0)
    e = oracle.min_eigen(img, bs)
    ec = cv2.cornerMinEigenVal(img, bs, ksize=3)
    assert (e != ec).sum() <= 8 and np.abs(e - ec).max() <= 1e-8
    mask = np.zeros((H, W), np.uint8)
    mask[H // 4:, : W // 2] = 7
    for m in (None, mask):
        for (mc, q, d) in [(5, 0.3, 7), (200, 0.01, 3), (0, 0.05, 0), (50, 0.01, 2.5), (50, 0.01, 10.4)]:
            got = oracle.good_features_to_track(img, mc, q, d, mask=m, blockSize=bs)
            ref = cv2.goodFeaturesToTrack(img, mc, q, d, mask=m, blockSize=bs)
            assert (got is None) == (ref is None)
            if ref is not None:
                assert np.array_equal(got, ref)


@pytest.mark.parametrize("win,ml,crit", [((15, 15), 2, (3, 10, 0.03)), ((21, 21), 3, (3, 30, 0.01)), ((31, 31), 4, (3, 10, 0.03)),
                                         ((5, 5), 2, (3, 10, 0.03)), ((9, 13), 2, (1, 5, 0)), ((7, 7), 0, (2, 0, 0.1))])
def test_pyrlk_within_tolerance(oracle, win, ml, crit):
    from lk_optical_flow_method_b200 import synth
    W, H = 320, 240
    tex = synth.make_texture(W, H, 9)
    a = cv2.GaussianBlur(cv2.cvtColor(synth.zoom_frame(tex, W, H, 4), cv2.COLOR_BGR2GRAY), (3, 3), 0)
    b = cv2.GaussianBlur(cv2.cvtColor(synth.zoom_frame(tex, W, H, 5), cv2.COLOR_BGR2GRAY), (3, 3), 0)
    c = cv2.goodFeaturesToTrack(a, 150, 0.01, 5).reshape(-1, 2)
    pts = np.concatenate([c, c[:30] + np.float32(0.37), synth.adversarial_points(W, H, 16)]).astype(np.float32).reshape(-1, 1, 2)
    r = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=win, maxLevel=ml, criteria=crit)
    o = oracle.pyrlk(a, b, pts, win, ml, crit)
    both = ((r[1] == 1) & (o[1] == 1)).ravel()
    assert (r[1] == o[1]).mean() >= 0.995
    assert np.abs(r[0] - o[0]).reshape(-1, 2)[both].max() <= 1e-2
    assert np.abs(r[2] - o[2]).ravel()[both].max() <= 1e-2
    # status-0 points keep the last iterate in nextPts (deterministic in cv2 too)
    z = (r[1] == 0).ravel() & (o[1] == 0).ravel()
    if z.any():
        assert np.abs(r[0] - o[0]).reshape(-1, 2)[z].max() <= 1e-2


@pytest.mark.parametrize("seed", range(60))
def test_randomised_sizes_and_parameters(oracle, seed):
    """Seeded sweep over frame sizes (incl. sizes where cv2 has to cut maxLevel), windows, block sizes, quality levels
    and min distances: integer stages and corner lists identical, LK within the north_star tolerances."""
    from lk_optical_flow_method_b200 import synth
    rng = np.random.default_rng(1000 + seed)
    W, H = int(rng.integers(40, 260)), int(rng.integers(36, 200))
    tex = synth.make_texture(W, H, seed)
    f0, f1 = synth.zoom_frame(tex, W, H, 3), synth.zoom_frame(tex, W, H, 4)
    a, b = (cv2.GaussianBlur(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 0) for f in (f0, f1))
    assert np.array_equal(oracle.blur3(oracle.bgr2gray(f0)), a)
    assert np.array_equal(oracle.pyrdown(a), cv2.pyrDown(a))
    bs = int(rng.choice([2, 3, 5, 7]))
    q = float(rng.choice([0.3, 0.05, 0.01]))
    d = float(rng.choice([0, 1, 3.5, 7, 12]))
    mc = int(rng.choice([0, 5, 40, 400]))
    mask = None
    if rng.uniform() < 0.5:
        mask = np.zeros((H, W), np.uint8)
        mask[int(H * 0.2):int(H * 0.9), int(W * 0.1):int(W * 0.8)] = 255
    got = oracle.good_features_to_track(a, mc, q, d, mask=mask, blockSize=bs)
    ref = cv2.goodFeaturesToTrack(a, mc, q, d, mask=mask, blockSize=bs)
    assert (got is None) == (ref is None)
    if ref is not None:
        assert np.array_equal(got, ref)
    win = (int(rng.choice([3, 5, 9, 15, 21, 31])), int(rng.choice([3, 5, 9, 15, 21, 31])))
    ml = int(rng.integers(0, 5))
    crit = [(3, 10, 0.03), (3, 30, 0.01), (1, 4, 0), (2, 0, 0.05)][int(rng.integers(0, 4))]
    base = ref.reshape(-1, 2) if ref is not None and len(ref) else np.zeros((0, 2), np.float32)
    pts = np.concatenate([base[:60], rng.uniform(-3, max(W, H) + 3, (40, 2)).astype(np.float32),
                          synth.adversarial_points(W, H, 8, seed)]).astype(np.float32).reshape(-1, 1, 2)
    r = cv2.calcOpticalFlowPyrLK(a, b, pts, None, winSize=win, maxLevel=ml, criteria=crit)
    o = oracle.pyrlk(a, b, pts, win, ml, crit)
    both = ((r[1] == 1) & (o[1] == 1)).ravel()
    assert (r[1] == o[1]).mean() >= 0.99, (W, H, win, ml, crit)
    if both.any():
        assert np.abs(r[0] - o[0]).reshape(-1, 2)[both].max() <= 1e-2, (W, H, win, ml, crit)
        assert np.abs(r[2] - o[2]).ravel()[both].max() <= 1e-2


@pytest.mark.parametrize("seed", range(24))
def test_resize_area_randomised(oracle, seed):
    """INTER_AREA down-scaling (imutils.resize, VP.py:433) on random geometries, 1 and 3 channels: bit-exact."""
    rng = np.random.default_rng(500 + seed)
    W, H = int(rng.integers(30, 400)), int(rng.integers(20, 300))
    s = float(rng.uniform(1.05, 3.9))
    wd, hd = max(2, int(W / s)), max(2, int(H / s))
    if W % wd == 0 and H % hd == 0:
        wd -= 1                                           # keep off cv2's integer-ratio fast path (not on the scripts' path)
    f = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    assert np.array_equal(oracle.resize_area(f, (wd, hd)), cv2.resize(f, (wd, hd), interpolation=cv2.INTER_AREA))
    g = np.ascontiguousarray(f[..., 0])
    assert np.array_equal(oracle.resize_area(g, (wd, hd)), cv2.resize(g, (wd, hd), interpolation=cv2.INTER_AREA))
