"""Run one operator a few times (for `ncu -k <kernel>` captures).  Usage: prof_one.py <op> [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lk_optical_flow_method_b200 import ops, synth

op = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
H, W = 1080, 1920
tex = torch.from_numpy(synth.make_texture(W, H, 0)).cuda()
frames = torch.stack([synth.zoom_frame(tex, W, H, t) for t in range(B)])
gray = ops.process_img(frames)
if op == "min_eigen":
    for _ in range(3):
        e = ops.min_eigen(gray, 7)
elif op == "gftt":
    e = ops.min_eigen(gray, 7)
    for _ in range(3):
        ops.gftt_select(e, None, 2000, 0.01, 7, 2000)
elif op == "process_img":
    for _ in range(3):
        ops.process_img(frames, out=gray)
elif op == "lk":
    from lk_optical_flow_method_b200.pipeline import PairPipeline
    LK = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 10, 0.03))
    pipe = PairPipeline(B - 1, H, W, 2000, LK, None, 0)
    pipe.prime(frames[0])
    pipe.stage_process(frames[1:]); pipe.stage_pyramid()
    e = ops.min_eigen(gray[:B - 1], 7)
    cor, cnt = ops.gftt_select(e, None, 2000, 0.01, 7, 2000)
    pipe.set_points(cor, cnt)
    for _ in range(3):
        pipe.stage_track()
elif op == "resize":
    for _ in range(3):
        r = ops.resize_area(frames, (860, 483))
elif op == "dense":
    LK = dict(winSize=(5, 5), maxLevel=2, criteria=(3, 10, 0.03))
    nl = ops.effective_levels(H, W, LK["winSize"], LK["maxLevel"])
    pyr = ops.build_pyramid(gray[:2], nl)
    ys, xs = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float()[None].contiguous()
    params = ops.lk_params_struct(**LK)
    for _ in range(3):
        ops.pyrlk([p[:1] for p in pyr], [p[1:2] for p in pyr], pts, params)
elif op == "pyrdown":
    for _ in range(3):
        l1 = ops.pyr_down(gray)
torch.cuda.synchronize()
print("done", op)
