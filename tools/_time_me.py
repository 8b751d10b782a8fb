import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lk_optical_flow_method_b200 import ops, synth
H, W, B = 1080, 1920, 8
tex = torch.from_numpy(synth.make_texture(W, H, 0)).cuda()
gray = ops.process_img(torch.stack([synth.zoom_frame(tex, W, H, t) for t in range(B)]))
e = ops.min_eigen(gray, 7)
for _ in range(3): ops.min_eigen(gray, 7, out=e)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(20): ops.min_eigen(gray, 7, out=e)
b.record(); torch.cuda.synchronize()
print("min_eigen<7> 8 frames: %.1f us  checksum %s" % (a.elapsed_time(b) / 20 * 1e3, float(e.double().sum())))
