"""GPU parity, a5/a6 (Shi-Tomasi corner scoring + selection) against the oracle: the corner list must
be identical INCLUDING ORDER; the eig map bit-identical up to a handful of 1e-9-level pixels."""
import numpy as np
import pytest

from conftest import processed_pair

pytestmark = pytest.mark.gpu


def _masks(W, H):
    m = np.zeros((4, H, W), np.uint8)
    m[0, H // 2:, : W // 2] = 255
    m[1, H // 2:, W // 2:] = 255
    m[2, : H // 2, W // 2:] = 1          # any non-zero value counts
    m[3, H // 3: 2 * H // 3, W // 3: 2 * W // 3] = 255
    return m


@pytest.mark.parametrize("W,H,bs", [(860, 483, 7), (1920, 1080, 7), (641, 359, 3), (320, 240, 5), (64, 48, 2)])
def test_min_eigen_map(oracle, W, H, bs):
    import torch
    from lk_optical_flow_method_b200 import ops
    prev, _ = processed_pair(W, H)
    got = ops.min_eigen(torch.from_numpy(prev).cuda()[None], bs)[0].cpu().numpy()
    ref = oracle.min_eigen(prev, bs)
    bad = got != ref
    assert bad.sum() <= 64, bad.sum()
    assert np.abs(got - ref).max() <= 1e-8


CASES = [
    (860, 483, dict(maxCorners=5, qualityLevel=0.3, minDistance=7, blockSize=7)),       # feature_params VP.py:92-95
    (860, 483, dict(maxCorners=500, qualityLevel=0.01, minDistance=7, blockSize=7)),
    (1920, 1080, dict(maxCorners=2000, qualityLevel=0.01, minDistance=7, blockSize=7)),  # C2 points
    (3840, 2160, dict(maxCorners=10000, qualityLevel=0.001, minDistance=3, blockSize=3)),  # C3 points
    (641, 359, dict(maxCorners=1000, qualityLevel=0.001, minDistance=3, blockSize=3)),
    (641, 359, dict(maxCorners=0, qualityLevel=0.05, minDistance=0, blockSize=3)),      # unlimited, no min distance
    (320, 240, dict(maxCorners=300, qualityLevel=0.01, minDistance=2.5, blockSize=5)),  # cvRound(2.5) = 2
    (320, 240, dict(maxCorners=300, qualityLevel=0.01, minDistance=7.4, blockSize=3)),
]


@pytest.mark.parametrize("W,H,fp", CASES)
def test_good_features_identical_to_oracle(oracle, W, H, fp):
    import lk_optical_flow_method_b200.cv as cv
    prev, _ = processed_pair(W, H)
    masks = _masks(W, H)
    for mask in [None, masks[0], masks[2], masks[3]]:
        got = cv.goodFeaturesToTrack(prev, mask=mask, **fp)
        ref = oracle.good_features_to_track(prev, fp["maxCorners"], fp["qualityLevel"], fp["minDistance"], mask=mask,
                                            blockSize=fp["blockSize"])
        assert (got is None) == (ref is None)
        if ref is not None:
            assert got.dtype == np.float32 and got.shape == ref.shape, (got.shape, ref.shape)
            assert np.array_equal(got, ref)


def test_good_features_edge_cases(oracle):
    import lk_optical_flow_method_b200.cv as cv
    prev, _ = processed_pair(320, 240)
    fp = dict(maxCorners=10, qualityLevel=0.3, minDistance=7, blockSize=7)
    assert cv.goodFeaturesToTrack(prev, mask=np.zeros_like(prev), **fp) is None          # empty mask -> None
    assert cv.goodFeaturesToTrack(np.full((64, 64), 128, np.uint8), **fp) is None         # flat image -> None
    # exact ties: periodic image; order must follow "larger raster index first"
    t = np.zeros((96, 128), np.uint8)
    t[::16, ::16] = 255
    got = cv.goodFeaturesToTrack(t, maxCorners=0, qualityLevel=0.5, minDistance=1, blockSize=3)
    ref = oracle.good_features_to_track(t, 0, 0.5, 1, blockSize=3)
    assert np.array_equal(got, ref)
    with pytest.raises(cv.error):
        cv.goodFeaturesToTrack(prev, 10, 0.0, 7)
    with pytest.raises(cv.error):
        cv.goodFeaturesToTrack(prev, 10, 0.1, 7, mask=np.zeros((3, 3), np.uint8))


def test_good_features_multi_mask_batched(oracle):
    """Device API: one eig map per image, several masks per image in one pass (the scripts call GFTT
    4x per refresh on the same image with 4 ROI masks, VP.py:500-507)."""
    import torch
    from lk_optical_flow_method_b200 import ops
    W, H = 860, 483
    imgs = [processed_pair(W, H, seed=s)[0] for s in range(2)]
    masks = _masks(W, H)
    eig = ops.min_eigen(torch.from_numpy(np.stack(imgs)).cuda(), 7)
    cor, cnt = ops.gftt_select(eig, torch.from_numpy(masks).cuda(), 5, 0.3, 7, 5, n_masks=4)
    cor, cnt = cor.cpu().numpy(), cnt.cpu().numpy()
    for b in range(2):
        for m in range(4):
            ref = oracle.good_features_to_track(imgs[b], 5, 0.3, 7, mask=masks[m], blockSize=7)
            k = cnt[b * 4 + m]
            assert k == (0 if ref is None else len(ref))
            if ref is not None:
                assert np.array_equal(cor[b * 4 + m, :k], ref.reshape(-1, 2))


def test_gftt_against_cv2_when_available(oracle):
    cv2 = pytest.importorskip("cv2")
    import lk_optical_flow_method_b200.cv as cv
    prev, _ = processed_pair(1920, 1080)
    fp = dict(maxCorners=2000, qualityLevel=0.01, minDistance=7, blockSize=7)
    got = cv.goodFeaturesToTrack(prev, **fp)
    ref = cv2.goodFeaturesToTrack(prev, **fp)
    assert np.array_equal(got, ref)


def _same_as_oracle(oracle, img, **fp):
    import lk_optical_flow_method_b200.cv as cv
    got = cv.goodFeaturesToTrack(img, mask=None, **fp)
    ref = oracle.good_features_to_track(img, fp["maxCorners"], fp["qualityLevel"], fp["minDistance"], blockSize=fp["blockSize"])
    assert (got is None) == (ref is None)
    if ref is not None:
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert np.array_equal(got, ref)
    return 0 if ref is None else len(ref)


def test_selection_paths_that_leave_the_fast_kernel(oracle):
    """The rank-by-bin selection kernel (k_sort_select_fast) hands lists it cannot finish to the general bitonic-sort kernel:
    (i) thousands of EXACTLY equal scores (one histogram bin holds them all: ties broken by raster index),
    (ii) more kept candidates than its shared memory ranks (maxCorners = 12 000 -> 18 256 to keep),
    (iii) the kept strongest candidates do not yield maxCorners corners under a large minDistance (full list needed).
    Each must still give the oracle's list, order included."""
    t = np.zeros((512, 640), np.uint8)
    t[4::8, 4::8] = 255                                          # 64 x 80 identical corners
    n = _same_as_oracle(oracle, t, maxCorners=3000, qualityLevel=0.5, minDistance=1, blockSize=3)
    assert n == 3000
    n = _same_as_oracle(oracle, t, maxCorners=0, qualityLevel=0.5, minDistance=3, blockSize=3)
    assert n > 3000
    prev, _ = processed_pair(1920, 1080)
    n = _same_as_oracle(oracle, prev, maxCorners=12000, qualityLevel=0.0001, minDistance=1, blockSize=3)
    assert n == 12000
    prev, _ = processed_pair(860, 483)
    n = _same_as_oracle(oracle, prev, maxCorners=300, qualityLevel=0.001, minDistance=40, blockSize=3)
    assert 50 < n < 300                                          # fewer than maxCorners exist at that spacing
