python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import cv2, numpy as np
from lk_optical_flow_method_b200 import synth
tex = synth.make_texture(1920, 1080, 0)
f = synth.zoom_frame(tex, 1920, 1080, 5)
ok, buf = cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])
open(".exp/frame.jpg", "wb").write(buf.tobytes()); print("jpeg bytes", len(buf), "of", f.nbytes)
import time
t=time.perf_counter()
for _ in range(20): cv2.imdecode(buf, cv2.IMREAD_COLOR)
print("cv2.imdecode ms", (time.perf_counter()-t)/20*1e3)
PY
./.exp/nvjpeg_probe .exp/frame.jpg 64
