"""Micro-benchmark of the stencil kernels: back-to-back launches between two CUDA events (amortises launch gaps)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lk_optical_flow_method_b200 import ops, synth

B, H, W = 32, 1080, 1920
tex = torch.from_numpy(synth.make_texture(W, H, 0)).cuda()
frames = torch.stack([synth.zoom_frame(tex, W, H, t) for t in range(B)])
peak = 6533.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
P = H * W
gray = ops.process_img(frames)
l1 = ops.pyr_down(gray); l2 = ops.pyr_down(l1)
eig = ops.min_eigen(gray[:4], 7)


def timeit(fn, nbytes, name, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s  {nbytes/ms/1e6/peak:6.3f} of HBM peak")


timeit(lambda: ops.process_img(frames, out=gray), B * 4 * P, "process_img (a1+a2)")
timeit(lambda: ops.bgr2gray(frames, out=gray), B * 4 * P, "bgr2gray (a1)")
g2 = torch.empty_like(gray)
timeit(lambda: ops.gaussian_blur3(gray, out=g2), B * 2 * P, "blur3 (a2)")
timeit(lambda: ops.pyr_down(gray, out=l1), B * 1.25 * P, "pyrdown L0->L1")
timeit(lambda: ops.pyr_down(l1, out=l2), B * 1.25 * P / 4, "pyrdown L1->L2")
timeit(lambda: ops.min_eigen(gray[:4], 7, out=eig), 4 * 5 * P, "min_eigen bs7 (4 img)")
eig1 = ops.min_eigen(gray[:1], 7)
timeit(lambda: ops.min_eigen(gray[:1], 7, out=eig1), 1 * 5 * P, "min_eigen bs7 (1 img)")
eig32 = ops.min_eigen(gray, 7)
timeit(lambda: ops.min_eigen(gray, 7, out=eig32), B * 5 * P, "min_eigen bs7 (32 img)")
eig3 = ops.min_eigen(gray, 3)
timeit(lambda: ops.min_eigen(gray, 3, out=eig3), B * 5 * P, "min_eigen bs3 (32 img)")
del eig32, eig3
sch = ops.scharr(gray[:8])
timeit(lambda: ops.scharr(gray[:8], out=sch), 8 * 5 * P, "scharr (8 img)")
a = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
timeit(lambda: b.copy_(a), 2 * a.numel(), "torch copy 256 MiB (ref)")
