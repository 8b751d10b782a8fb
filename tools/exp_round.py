"""Per-stage timings of the bench workload with a checksum of the outputs (bit-identity check across builds): the LK
stage, the two pyrDown levels, the INTER_AREA resize.  Any LKF_* environment variable is echoed in the tag, so a
temporary tuning switch (read once per process by the library) can be compared run against run.
Usage: exp_round.py lk|pyr|resize [B]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lk_optical_flow_method_b200 import ops, synth

what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
H, W = 1080, 1920
tex = torch.from_numpy(synth.make_texture(W, H, 0)).cuda()
frames = torch.stack([synth.zoom_frame(tex, W, H, t) for t in range(B + 1)])
gray = ops.process_img(frames)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def digest(*ts):
    h = hashlib.sha1()
    for t in ts:
        h.update(t.detach().cpu().numpy().tobytes())
    return h.hexdigest()[:12]


tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("LKF_"))
if what == "lk":
    from lk_optical_flow_method_b200.pipeline import PairPipeline
    LK = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 10, 0.03))
    pipe = PairPipeline(B, H, W, 2000, LK, None, 0)
    pipe.prime(frames[0])
    pipe.stage_process(frames[1:]); pipe.stage_pyramid()
    e = ops.min_eigen(gray[:B], 7)
    cor, cnt = ops.gftt_select(e, None, 2000, 0.01, 7, 2000)
    pipe.set_points(cor, cnt)
    us = timeit(pipe.stage_track)
    outs = [getattr(pipe, n) for n in ("next_pts", "status", "err") if hasattr(pipe, n)]
    print(f"[{tag}] lk stage {us:8.1f} us per {B} pairs  digest {digest(*outs) if outs else '-'}")
elif what == "resize":
    for (h, w) in ((1080, 1920), (720, 1280)):
        src = frames[:, :h, :w].contiguous()
        out = ops.resize_area(src, (860, 483))
        us = timeit(lambda: ops.resize_area(src, (860, 483), out=out))
        print(f"[{tag}] resize_area {w}x{h} -> 860x483: {us:7.1f} us per {B + 1} frames ({src.numel() / us / 1e6:6.2f} TB/s read)  digest {digest(out)}")
else:
    l1 = ops.pyr_down(gray); l2 = ops.pyr_down(l1)
    a = timeit(lambda: ops.pyr_down(gray, out=l1)); b = timeit(lambda: ops.pyr_down(l1, out=l2))
    print(f"[{tag}] pyrdown L0->L1 {a:7.1f} us  L1->L2 {b:7.1f} us per {B + 1} frames  digest {digest(l1, l2)}")
