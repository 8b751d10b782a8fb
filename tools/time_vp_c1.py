"""C1 as BASELINE.json names it (GRMN0307 texture, 1280x720, TP_NUM = 500 -> up to C(500,2) flow-line pairs per frame): run the
first N frames of the clip through the clip runner (one clip).  Under `ncu --metrics gpu__time_duration.sum -k regex:k_vp_frame`
this gives the per-frame time of k_vp_frame at L ~ 400-500 lines.  Usage: time_vp_c1.py [frames] [TP_NUM]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2, numpy as np, torch
from lk_optical_flow_method_b200 import synth, vp as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
meta = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_grmn300_tp500.json")))
over = dict(meta["overrides"])
if len(sys.argv) > 2:
    over["TP_NUM"] = int(sys.argv[2])
W, H = meta["W"], meta["H"]
crop = os.path.join(ROOT, "baseline", "_ref", "GRMN0307_crop.png")
tex = cv2.resize(cv2.imread(crop), (2 * W, 2 * H), interpolation=cv2.INTER_CUBIC) if os.path.exists(crop) else synth.make_texture(W, H, 0)
frames = [synth.zoom_frame(tex, W, H, t) for t in range(n)]
rows, runner = V.run_clips((torch.from_numpy(f[None]).cuda() for f in frames), variant="VP", expected_frames=n + 4, **over)
torch.cuda.synchronize()
print("frames", n, "TP_NUM", over.get("TP_NUM"), "rows", len(rows[0]))
