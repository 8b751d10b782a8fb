"""GPU parity, integer stages (a1-a4): bit-exact against the oracle (and cv2 when importable)."""
import numpy as np
import pytest

from conftest import synth_frames

pytestmark = pytest.mark.gpu

SIZES = [(860, 483), (1920, 1080), (641, 359), (67, 35), (5, 4)]


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("W,H", SIZES)
def test_gray_blur_pyr_scharr_bit_exact(oracle, W, H):
    import torch
    from lk_optical_flow_method_b200 import ops
    rng = np.random.default_rng(W * 7 + H)
    frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8)]
    if W >= 64:
        frames.append(synth_frames(W, H, 1)[0])
    bgr = np.stack(frames)
    d = _dev(bgr)
    gray = ops.bgr2gray(d)
    ref_gray = np.stack([oracle.bgr2gray(f) for f in bgr])
    assert np.array_equal(gray.cpu().numpy(), ref_gray)
    blur = ops.gaussian_blur3(gray)
    ref_blur = np.stack([oracle.blur3(g) for g in ref_gray])
    assert np.array_equal(blur.cpu().numpy(), ref_blur)
    fused = ops.process_img(d)
    assert np.array_equal(fused.cpu().numpy(), ref_blur)
    pd = ops.pyr_down(blur)
    ref_pd = np.stack([oracle.pyrdown(g) for g in ref_blur])
    assert np.array_equal(pd.cpu().numpy(), ref_pd)
    if min(pd.shape[1:]) >= 2:
        pd2 = ops.pyr_down(pd)
        assert np.array_equal(pd2.cpu().numpy(), np.stack([oracle.pyrdown(g) for g in ref_pd]))
    sch = ops.scharr(blur)
    ref_s = np.stack([oracle.scharr(g) for g in ref_blur])
    assert np.array_equal(sch.cpu().numpy(), ref_s)
    torch.cuda.synchronize()


@pytest.mark.parametrize("W,H,levels", [(1920, 1080, 3), (960, 540, 3), (480, 16, 3), (1440, 36, 3), (3840, 2160, 3),
                                        (1920, 1080, 4), (860, 483, 3), (960, 542, 3), (67, 35, 2)])
def test_process_pyramid_bit_exact(oracle, W, H, levels):
    """lkf_process_pyramid_u8: process_img + the whole pyramid in one ABI call -- every level bit-exact against the oracle,
    including the mirrored borders of every level and the seams between thread strips."""
    import torch
    from lk_optical_flow_method_b200 import ops
    from lk_optical_flow_method_b200.device import pyramid_shapes
    rng = np.random.default_rng(W + 3 * H)
    bgr = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    if W >= 64 and W * H <= 1920 * 1080:
        bgr[1] = synth_frames(W, H, 1)[0]
    d = _dev(bgr)
    outs = [torch.full((2, h, w), 7, dtype=torch.uint8, device="cuda") for (h, w) in pyramid_shapes(H, W, levels)]
    ops.process_pyramid(d, outs)
    ref = [np.stack([oracle.blur3(oracle.bgr2gray(f)) for f in bgr])]
    for l in range(1, levels):
        ref.append(np.stack([oracle.pyrdown(g) for g in ref[-1]]))
    for l in range(levels):
        got = outs[l].cpu().numpy()
        assert got.shape == ref[l].shape
        bad = np.argwhere(got != ref[l])
        assert len(bad) == 0, (l, len(bad), bad[:5])


def test_unaligned_views(oracle):
    """Row pitch / base pointers that are not 4- or 16-byte aligned take the scalar paths."""
    from lk_optical_flow_method_b200 import ops
    rng = np.random.default_rng(5)
    big = rng.integers(0, 256, (2, 70, 131, 3), dtype=np.uint8)
    d = _dev(big)
    view = d[:, 3:64, 1:122]            # 61 x 121, offset base, pitch 131*3
    ref_gray = np.stack([oracle.bgr2gray(f) for f in big[:, 3:64, 1:122]])
    assert np.array_equal(ops.bgr2gray(view).cpu().numpy(), ref_gray)
    ref_blur = np.stack([oracle.blur3(g) for g in ref_gray])
    assert np.array_equal(ops.process_img(view).cpu().numpy(), ref_blur)
    g = _dev(np.pad(ref_gray, ((0, 0), (2, 1), (3, 2))))[:, 2:-1, 3:-2]
    assert np.array_equal(ops.gaussian_blur3(g).cpu().numpy(), ref_blur)
    assert np.array_equal(ops.pyr_down(g).cpu().numpy(), np.stack([oracle.pyrdown(x) for x in ref_gray]))


def test_against_cv2_when_available(oracle):
    cv2 = pytest.importorskip("cv2")
    from lk_optical_flow_method_b200 import ops
    f = synth_frames(860, 483, 1)[0]
    d = _dev(f[None])
    g = cv2.GaussianBlur(cv2.cvtColor(f, cv2.COLOR_BGR2GRAY), (3, 3), 0)
    out = ops.process_img(d)
    assert np.array_equal(out[0].cpu().numpy(), g)
    assert np.array_equal(ops.pyr_down(out)[0].cpu().numpy(), cv2.pyrDown(g))
    s = ops.scharr(out)[0].cpu().numpy()
    assert np.array_equal(s[..., 0], cv2.Scharr(g, cv2.CV_16S, 1, 0))
    assert np.array_equal(s[..., 1], cv2.Scharr(g, cv2.CV_16S, 0, 1))


def test_cv_resize_dropin_is_cv2_resize():
    """cv.resize (what imutils.resize calls, VP.py:433,531): INTER_AREA non-integer down-scales of uint8 frames run on the GPU
    and equal cv2.resize byte for byte; every other use is forwarded to cv2; patch(cv2) swaps it in."""
    cv2 = pytest.importorskip("cv2")
    import torch
    import lk_optical_flow_method_b200.cv as cv
    from lk_optical_flow_method_b200.device import Workspace
    from conftest import synth_frames
    lib = Workspace.current().lib
    for (W, H) in ((1280, 720), (1920, 1080)):
        f = synth_frames(W, H, 1)[0]
        n0 = lib.lkf_ctx_launch_count(None)
        got = cv.resize(f, (860, 483), interpolation=cv2.INTER_AREA)
        assert lib.lkf_ctx_launch_count(None) > n0                                   # the CUDA kernel ran
        assert got.dtype == np.uint8 and np.array_equal(got, cv2.resize(f, (860, 483), interpolation=cv2.INTER_AREA))
        g = np.ascontiguousarray(f[..., 1])
        assert np.array_equal(cv.resize(g, (430, 241), interpolation=cv2.INTER_AREA), cv2.resize(g, (430, 241), interpolation=cv2.INTER_AREA))
    d = cv.resize(torch.from_numpy(f).cuda(), (860, 483), interpolation=cv2.INTER_AREA)  # device in -> device out
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), cv2.resize(f, (860, 483), interpolation=cv2.INTER_AREA))
    n0 = lib.lkf_ctx_launch_count(None)
    assert np.array_equal(cv.resize(f, (960, 540), interpolation=cv2.INTER_AREA), cv2.resize(f, (960, 540), interpolation=cv2.INTER_AREA))  # integer factor
    assert np.array_equal(cv.resize(f, (860, 483)), cv2.resize(f, (860, 483)))          # INTER_LINEAR
    assert np.array_equal(cv.resize(f, None, fx=0.5, fy=0.5), cv2.resize(f, None, fx=0.5, fy=0.5))
    assert lib.lkf_ctx_launch_count(None) == n0                                       # all three went to cv2
    saved = cv.patch(cv2)
    try:
        assert cv2.resize is cv.resize
        assert np.array_equal(cv2.resize(f, (860, 483), interpolation=cv2.INTER_AREA), got)
    finally:
        cv.unpatch(cv2, saved)
    assert cv2.resize is not cv.resize
