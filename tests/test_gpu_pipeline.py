"""GPU: the frame-pair pipeline that bench.py times (PairPipeline, two streams inside a step, CUDA-graph replay, host
feeder) must produce, pair by pair, exactly what the drop-in calls produce one at a time -- and those are checked
against the oracle.  Guards the benchmark against skipped or reordered work."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, B, N, EVERY = 640, 360, 7, 300, 3
LK = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 10, 0.03))
GFTT = dict(maxCorners=N, qualityLevel=0.01, minDistance=7, blockSize=7)


def _inputs():
    import torch
    from lk_optical_flow_method_b200 import synth
    tex = synth.make_texture(W, H, 3)
    frames = np.stack([synth.zoom_frame(tex, W, H, 2 + t) for t in range(B + 1)])
    rng = np.random.default_rng(0)
    pts = np.stack([np.stack([rng.uniform(-5, W + 5, N), rng.uniform(-5, H + 5, N)], 1) for _ in range(B)]).astype(np.float32)
    return frames, pts


def _reference(frames, pts):
    """Pair by pair through the cv2-signature boundary (numpy in / numpy out)."""
    import lk_optical_flow_method_b200.cv as cv
    gray = [cv.GaussianBlur(cv.cvtColor(f, cv.COLOR_BGR2GRAY), (3, 3), 0) for f in frames]
    out = []
    for i in range(B):
        p = pts[i].reshape(-1, 1, 2)
        if i % EVERY == 0:
            c = cv.goodFeaturesToTrack(gray[i], mask=None, **GFTT)
            p = c if c is not None else np.zeros((0, 1, 2), np.float32)
        if len(p) == 0:
            out.append((p, None, None, None))
            continue
        p1, st, err = cv.calcOpticalFlowPyrLK(gray[i], gray[i + 1], np.ascontiguousarray(p), None, **LK)
        out.append((p, p1, st, err))
    return gray, out


def _check(pipe_out, ref):
    nxt, st, err, cnt = [x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x) for x in pipe_out]
    for i, (p, p1, s, e) in enumerate(ref):
        k = len(p)
        assert int(cnt[i]) == k, (i, int(cnt[i]), k)
        if k == 0:
            continue
        assert np.array_equal(nxt[i, :k], p1.reshape(-1, 2)), i
        assert np.array_equal(st[i, :k], s.ravel()), i
        assert np.array_equal(err[i, :k], e.ravel()), i


def test_pair_pipeline_matches_single_calls(oracle):
    import torch
    from lk_optical_flow_method_b200.pipeline import HostFeeder, PairPipeline
    frames, pts = _inputs()
    gray, ref = _reference(frames, pts)
    # the drop-in calls themselves against the oracle on one refreshed and one ordinary pair
    for i in (0, 1):
        p, p1, st, err = ref[i]
        q1, sq, eq = oracle.pyrlk(gray[i], gray[i + 1], np.ascontiguousarray(p), **LK)
        assert (st == sq).mean() >= 0.995
        both = ((st == 1) & (sq == 1)).ravel()
        assert np.abs(p1 - q1).reshape(-1, 2)[both].max() <= 1e-2
    d_frames = torch.from_numpy(frames).cuda()
    d_pts = torch.from_numpy(pts).cuda()
    for overlap in (True, False):
        pipe = PairPipeline(B, H, W, N, LK, GFTT, EVERY)
        pipe.overlap = overlap
        pipe.prime(d_frames[0])
        pipe.set_points(d_pts)
        pipe.step(d_frames[1:])                                   # eager, two streams when overlap
        torch.cuda.synchronize()
        _check(pipe.results(), ref)
        for lv, g in zip([pipe.levels[0][i] for i in range(1, B + 1)], gray[1:]):
            assert np.array_equal(lv.cpu().numpy(), g)
    # CUDA-graph replay (what `value` times), twice: the step must be re-runnable on the same inputs
    pipe = PairPipeline(B, H, W, N, LK, GFTT, EVERY)
    pipe.prime(d_frames[0])
    pipe.set_points(d_pts)
    replay = pipe.capture(d_frames[1:])
    for _ in range(2):
        pipe.prime(d_frames[0])
        pipe.set_points(d_pts)
        pipe.counts.fill_(N)
        pipe.next_pts.zero_(); pipe.status.zero_()
        replay()
        torch.cuda.synchronize()
        _check(pipe.results(), ref)
    # host feeder (what `e2e` times): pinned host buffers in, pinned host results out
    pipe.prime(d_frames[0])
    feeder = HostFeeder(pipe)
    hf, hp = torch.from_numpy(frames[1:]).pin_memory(), torch.from_numpy(pts).pin_memory()
    outs = feeder.run([(hf, hp)])
    _check(outs[0], ref)
