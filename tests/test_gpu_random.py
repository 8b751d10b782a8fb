"""Seeded randomised sweep of the CUDA path (through the cv2-signature boundary) against the oracle: odd frame sizes
(unaligned rows, levels cut by the window), square and non-square windows (specialised and generic trackers), every
block size, masks, quality levels and min distances.  Integer stages and corner lists identical; LK within the
north_star tolerances (flow <= 1e-2 px, status agreement >= 99.5 %)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(64))
def test_random_geometry_and_parameters(oracle, seed):
    from lk_optical_flow_method_b200 import cv, synth
    rng = np.random.default_rng(2000 + seed)
    W, H = int(rng.integers(40, 300)), int(rng.integers(36, 220))
    tex = synth.make_texture(W, H, seed)
    f0, f1 = synth.zoom_frame(tex, W, H, 3), synth.zoom_frame(tex, W, H, 4)
    a = cv.GaussianBlur(cv.cvtColor(f0, cv.COLOR_BGR2GRAY), (3, 3), 0)
    b = cv.GaussianBlur(cv.cvtColor(f1, cv.COLOR_BGR2GRAY), (3, 3), 0)
    assert np.array_equal(a, oracle.blur3(oracle.bgr2gray(f0))) and np.array_equal(b, oracle.blur3(oracle.bgr2gray(f1)))
    bs = int(rng.choice([3, 5, 7]))
    q = float(rng.choice([0.3, 0.05, 0.01]))
    d = float(rng.choice([0, 1, 3.5, 7, 12]))
    mc = int(rng.choice([0, 5, 40, 400]))
    mask = None
    if rng.uniform() < 0.5:
        mask = np.zeros((H, W), np.uint8)
        mask[int(H * 0.2):int(H * 0.9), int(W * 0.1):int(W * 0.8)] = 255
    got = cv.goodFeaturesToTrack(a, mc, q, d, mask=mask, blockSize=bs)
    ref = oracle.good_features_to_track(a, mc, q, d, mask=mask, blockSize=bs)
    assert (got is None) == (ref is None)
    if ref is not None:
        assert np.array_equal(got, ref)
    if rng.uniform() < 0.6:
        w = int(rng.choice([3, 5, 7, 15, 21, 31]))
        win = (w, w)
    else:
        win = (int(rng.choice([3, 5, 9, 15, 21])), int(rng.choice([5, 9, 13, 31])))
    ml = int(rng.integers(0, 5))
    crit = [(3, 10, 0.03), (3, 30, 0.01), (1, 4, 0), (2, 0, 0.05)][int(rng.integers(0, 4))]
    base = ref.reshape(-1, 2) if ref is not None and len(ref) else np.zeros((0, 2), np.float32)
    pts = np.concatenate([base[:60], rng.uniform(-3, max(W, H) + 3, (40, 2)).astype(np.float32),
                          synth.adversarial_points(W, H, 8, seed)]).astype(np.float32).reshape(-1, 1, 2)
    n1, st, err = cv.calcOpticalFlowPyrLK(a, b, pts, None, winSize=win, maxLevel=ml, criteria=crit)
    o = oracle.pyrlk(a, b, pts, win, ml, crit)
    both = ((st == 1) & (o[1] == 1)).ravel()
    assert (st == o[1]).mean() >= 0.995, (W, H, win, ml, crit)
    if both.any():
        assert np.abs(n1 - o[0]).reshape(-1, 2)[both].max() <= 1e-2, (W, H, win, ml, crit)
        assert np.abs(err - o[2]).ravel()[both].max() <= 1e-2
