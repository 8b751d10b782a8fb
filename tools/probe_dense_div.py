"""C4 probe: iteration statistics of the dense 5x5 grid (mean per feature, mean of the per-warp maximum for the
thread-per-feature mapping = 32 consecutive x), status rate, time.  Usage: probe_dense_div.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lk_optical_flow_method_b200 import ops, synth

H, W = 1080, 1920
tex = torch.from_numpy(synth.make_texture(W, H, 0)).cuda()
frames = torch.stack([synth.zoom_frame(tex, W, H, t) for t in range(2)])
gray = ops.process_img(frames)
LK = dict(winSize=(5, 5), maxLevel=2, criteria=(3, 10, 0.03))
nl = ops.effective_levels(H, W, LK["winSize"], LK["maxLevel"])
pyr = ops.build_pyramid(gray, nl)
ys, xs = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
pts = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1).float()[None].contiguous()
params = ops.lk_params_struct(**LK)
nxt, st, err, it = ops.pyrlk([p[:1] for p in pyr], [p[1:2] for p in pyr], pts, params, want_iters=True)
it = it.reshape(H, W).float()
print("mean iterations/feature", it.mean().item(), "status rate", st.float().mean().item())
wm = it.reshape(H, W // 32, 32).max(-1).values
print("mean of warp(32x1)-max", wm.mean().item(), "-> lane utilisation of the iteration loops", (it.mean() / wm.mean()).item())
bm = it.reshape(H // 4, 4, W // 8, 8).amax((1, 3))
print("mean of block(8x4)-max", bm.mean().item())
hist = torch.bincount(it.reshape(-1).long(), minlength=31)
print("iteration histogram", hist.tolist())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.pyrlk([p[:1] for p in pyr], [p[1:2] for p in pyr], pts, params)
e1.record(); torch.cuda.synchronize()
print("dense pair ms", e0.elapsed_time(e1) / 5)
