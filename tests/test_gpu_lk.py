"""GPU parity, a7 (pyramidal LK tracker) through the C ABI, against the oracle on the same inputs.
Tolerances (BASELINE.json north_star): flow within 1e-2 px on status-agreeing points, status agreement
>= 99.5 %; err compared on status-1 points only (cv2 leaves it undefined elsewhere)."""
import numpy as np
import pytest

from conftest import processed_pair

pytestmark = pytest.mark.gpu

FLOW_TOL = 1e-2
STATUS_MIN = 0.995


def _points(oracle, prev, W, H, n, q, d, bs, seed=0):
    from lk_optical_flow_method_b200 import synth
    c = oracle.good_features_to_track(prev, n, q, d, blockSize=bs)
    pts = c.reshape(-1, 2)
    rng = np.random.default_rng(seed)
    flat = np.stack([rng.uniform(0, W, 16), rng.uniform(0, H, 16)], 1).astype(np.float32)
    frac = pts[: len(pts) // 4] + rng.uniform(-0.5, 0.5, (len(pts) // 4, 2)).astype(np.float32)
    return np.concatenate([pts, frac, flat, synth.adversarial_points(W, H, max(8, n // 50))]).astype(np.float32)


def _compare(got, ref):
    p1, st, err = got
    q1, sq, eq = ref
    agree = (st == sq).mean()
    both = ((st == 1) & (sq == 1)).ravel()
    dflow = np.abs(p1.reshape(-1, 2) - q1.reshape(-1, 2))[both].max() if both.any() else 0.0
    derr = np.abs(err - eq).ravel()[both].max() if both.any() else 0.0
    return agree, dflow, derr


CASES = [
    # W, H, n, q, d, bs, win, maxLevel, criteria
    (860, 483, 500, 0.01, 7, 7, (15, 15), 2, (3, 10, 0.03)),       # script lk_params (VP.py:98-100)
    (1920, 1080, 2000, 0.01, 7, 7, (21, 21), 2, (3, 10, 0.03)),    # C2
    (3840, 2160, 10000, 0.001, 3, 3, (31, 31), 3, (3, 10, 0.03)),  # C3: 4K, 10k features, 4-level pyramid, 31x31
    (641, 359, 300, 0.05, 3.5, 3, (31, 31), 4, (3, 30, 0.01)),     # cv2 defaults, level clamp
    (320, 240, 200, 0.01, 3, 3, (5, 5), 2, (3, 10, 0.03)),         # C4 window
    (320, 240, 200, 0.01, 3, 3, (9, 13), 3, (1, 7, 0.0)),          # non-square, COUNT only
    (320, 240, 200, 0.01, 3, 3, (11, 11), 1, (2, 0, 0.05)),        # EPS only -> 30 iterations max
]


@pytest.mark.parametrize("W,H,n,q,d,bs,win,ml,crit", CASES)
def test_pyrlk_matches_oracle(oracle, W, H, n, q, d, bs, win, ml, crit):
    import lk_optical_flow_method_b200.cv as cv
    prev, nxt = processed_pair(W, H)
    pts = _points(oracle, prev, W, H, n, q, d, bs).reshape(-1, 1, 2)
    got = cv.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=win, maxLevel=ml, criteria=crit)
    ref = oracle.pyrlk(prev, nxt, pts, win, ml, crit)
    assert got[0].shape == pts.shape and got[0].dtype == np.float32
    assert got[1].shape == (len(pts), 1) and got[1].dtype == np.uint8
    assert got[2].shape == (len(pts), 1) and got[2].dtype == np.float32
    agree, dflow, derr = _compare(got, ref)
    assert agree >= STATUS_MIN, agree
    assert dflow <= FLOW_TOL, dflow
    assert derr <= 1e-2, derr
    assert (ref[1] == 0).any() and (ref[1] == 1).any()      # both status paths exercised
    got[1][:] = 0                                           # outputs are writable like cv2's (VP.py:553)


def test_pyrlk_flags_and_shapes(oracle):
    import lk_optical_flow_method_b200.cv as cv
    W, H = 320, 240
    prev, nxt = processed_pair(W, H)
    pts = _points(oracle, prev, W, H, 100, 0.01, 5, 3)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))
    # (N,2) and (1,N,2) layouts are accepted and mirrored in the output
    for shape in [(-1, 2), (1, -1, 2)]:
        p = pts.reshape(shape)
        got = cv.calcOpticalFlowPyrLK(prev, nxt, p, None, **lk)
        assert got[0].shape == p.shape
        ref = oracle.pyrlk(prev, nxt, p, **lk)
        agree, dflow, _ = _compare(got, ref)
        assert agree >= STATUS_MIN and dflow <= FLOW_TOL
    # OPTFLOW_USE_INITIAL_FLOW
    init = (pts + np.float32(1.5)).reshape(-1, 1, 2)
    got = cv.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), init.copy(), flags=cv.OPTFLOW_USE_INITIAL_FLOW, **lk)
    ref = oracle.pyrlk(prev, nxt, pts.reshape(-1, 1, 2), flags=4, nextPts=init.copy(), **lk)
    agree, dflow, _ = _compare(got, ref)
    assert agree >= STATUS_MIN and dflow <= FLOW_TOL
    # OPTFLOW_LK_GET_MIN_EIGENVALS: err carries the min eigenvalue
    got = cv.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2), None, flags=cv.OPTFLOW_LK_GET_MIN_EIGENVALS, **lk)
    ref = oracle.pyrlk(prev, nxt, pts.reshape(-1, 1, 2), flags=8, **lk)
    agree, dflow, derr = _compare(got, ref)
    assert agree >= STATUS_MIN and dflow <= FLOW_TOL and derr <= 1e-4
    # empty input and argument errors behave like cv2
    assert cv.calcOpticalFlowPyrLK(prev, nxt, np.zeros((0, 1, 2), np.float32), None, **lk) == (None, None, None)
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2)[::5], None, **lk)        # non-contiguous
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(prev, nxt, pts.astype(np.float64), None, **lk)
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(2, 2))
    with pytest.raises(cv.error):
        cv.calcOpticalFlowPyrLK(prev, nxt[:-1], pts, None, **lk)


def test_pyrlk_against_cv2_when_available(oracle):
    cv2 = pytest.importorskip("cv2")
    import lk_optical_flow_method_b200.cv as cv
    W, H = 860, 483
    prev, nxt = processed_pair(W, H)
    pts = _points(oracle, prev, W, H, 500, 0.01, 7, 7).reshape(-1, 1, 2)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))
    got = cv.calcOpticalFlowPyrLK(prev, nxt, pts, None, **lk)
    ref = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, **lk)
    agree, dflow, derr = _compare(got, ref)
    assert agree >= STATUS_MIN and dflow <= FLOW_TOL and derr <= 1e-2


def test_pyrlk_batched_device_api(oracle):
    """Batched, device-resident entry: ragged point counts, several images per launch."""
    import torch
    from lk_optical_flow_method_b200 import ops
    W, H = 320, 240
    pairs = [processed_pair(W, H, seed=s, t0=2 + s) for s in range(3)]
    pts = [_points(oracle, p[0], W, H, 60 + 30 * i, 0.01, 5, 3, seed=i) for i, p in enumerate(pairs)]
    N = max(len(p) for p in pts)
    P = np.zeros((3, N, 2), np.float32)
    for i, p in enumerate(pts):
        P[i, :len(p)] = p
    prev = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    nxt = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    counts = torch.tensor([len(p) for p in pts], dtype=torch.int32, device="cuda")
    params = ops.lk_params_struct((15, 15), 2, (3, 10, 0.03))
    nl = ops.effective_levels(H, W, (15, 15), 2)
    out = ops.pyrlk(ops.build_pyramid(prev, nl), ops.build_pyramid(nxt, nl), torch.from_numpy(P).cuda(), params,
                    counts=counts, want_iters=True)
    n1, st, er, it = [o.cpu().numpy() for o in out]
    for i, p in enumerate(pts):
        ref = oracle.pyrlk(pairs[i][0], pairs[i][1], p.reshape(-1, 1, 2), (15, 15), 2, (3, 10, 0.03), return_iters=True)
        k = len(p)
        agree, dflow, _ = _compare((n1[i, :k], st[i, :k].reshape(-1, 1), er[i, :k].reshape(-1, 1)), ref[:3])
        assert agree >= STATUS_MIN and dflow <= FLOW_TOL
        assert np.abs(it[i, :k] - ref[3]).mean() < 0.05          # same iteration counts (roofline op count)


def _grid(W, H):
    ys, xs = np.mgrid[0:H, 0:W]
    return np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32).reshape(-1, 1, 2)


@pytest.mark.parametrize("win,crit", [((5, 5), (3, 10, 0.03)), ((5, 5), (1, 1, 0.0)), ((3, 3), (3, 10, 0.03)), ((7, 7), (3, 10, 0.03))])
def test_dense_flow_small_frame_matches_oracle(oracle, win, crit):
    """C4 (dense per-pixel LK, small window, 3-level pyramid) on a frame the oracle finishes in seconds: every pixel of
    the frame is a feature -- image borders, flat regions and out-of-range iterates included."""
    import lk_optical_flow_method_b200.cv as cv
    W, H = 200, 120
    prev, nxt = processed_pair(W, H)
    pts = _grid(W, H)
    got = cv.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=win, maxLevel=2, criteria=crit)
    ref = oracle.pyrlk(prev, nxt, pts, win, 2, crit)
    agree, dflow, derr = _compare(got, ref)
    assert agree >= STATUS_MIN and dflow <= FLOW_TOL and derr <= 1e-2, (agree, dflow, derr)
    assert (ref[1] == 0).any() and (ref[1] == 1).any()


def test_dense_flow_1080p_full_grid(oracle):
    """C4 at full size: 1920x1080 = 2 073 600 features, 5x5 window, 3 levels.  Features are independent, so (i) a strided
    subset tracked alone must reproduce the dense run bit for bit, and (ii) that subset is checked against the oracle."""
    import lk_optical_flow_method_b200.cv as cv
    W, H = 1920, 1080
    prev, nxt = processed_pair(W, H)
    lk = dict(winSize=(5, 5), maxLevel=2, criteria=(3, 10, 0.03))
    pts = _grid(W, H)
    dense = cv.calcOpticalFlowPyrLK(prev, nxt, pts, None, **lk)
    assert dense[0].shape == pts.shape and dense[1].shape == (W * H, 1)
    sel = np.arange(0, W * H, 61)
    sub = cv.calcOpticalFlowPyrLK(prev, nxt, np.ascontiguousarray(pts[sel]), None, **lk)
    for a, b in zip(dense, sub):
        assert np.array_equal(a[sel], b)
    ref = oracle.pyrlk(prev, nxt, np.ascontiguousarray(pts[sel]), (5, 5), 2, (3, 10, 0.03))
    agree, dflow, derr = _compare(sub, ref)
    assert agree >= STATUS_MIN and dflow <= FLOW_TOL and derr <= 1e-2, (agree, dflow, derr)
    assert 0.2 < dense[1].mean() < 1.0                   # a plausible share of pixels is trackable with a 5x5 window


def test_tma_form_of_the_warp_tracker_is_bit_identical():
    """k_lk_track_t has a second form that stages the `next` tile and the template neighbourhood with TMA boxes
    (cp.async.bulk.tensor + mbarrier; LKF_LK_TMA=1, off by default because it measured slower: profiles/r02_lk_tma.md).
    The switch is read once per process, so both forms run in child processes on the C2 stage (4 pairs x 2000 features,
    21x21, 3 levels) and must print the same digest of (nextPts, status, err)."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for flag in ("0", "1"):
        env = dict(os.environ, LKF_LK_TMA=flag)
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "exp_round.py"), "lk", "4"], env=env, capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append(re.search(r"digest (\w+)", out.stdout).group(1))
    assert digests[0] == digests[1] and digests[0] != "-"
