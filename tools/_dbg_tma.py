import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lk_optical_flow_method_b200 import ops, synth
H, W = 1080, 1920
tex = torch.from_numpy(synth.make_texture(W, H, 0)).cuda()
frames = torch.stack([synth.zoom_frame(tex, W, H, t) for t in range(2)])
gray = ops.process_img(frames)
pyr = ops.build_pyramid(gray, 3)
pts = torch.tensor([[[500.3, 400.7], [900.0, 300.0], [1200.5, 700.25]]], device="cuda")
params = ops.lk_params_struct((21, 21), 2, (3, 10, 0.03))
out = ops.pyrlk([p[:1] for p in pyr], [p[1:] for p in pyr], pts, params)
torch.cuda.synchronize()
print("ok", out[0].cpu().numpy().tolist(), out[1].cpu().numpy().tolist())
